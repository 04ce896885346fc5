// placeholder until the tcgen05 kernel lands (replaced below in this round)
#include "common.cuh"
namespace fn {
int siren_points_fast(const FnLayout&, const unsigned char*, const float*, const float*, const float*, int, long long,
                      int, int, float*, cudaStream_t) {
    return fail(FENERF_E_UNSUPPORTED, "tcgen05 point-network kernel not built");
}
}  // namespace fn
