// Shared host/device helpers of libfenerf_b200.  Internal.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/fenerf_b200.h"
#include "layout.h"

namespace fn {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define FN_CUDA_OK(expr)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return fn::fail(FENERF_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                 \
    } while (0)

#define FN_LAUNCH_OK(name)                                                                        \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess)                                                                   \
            return fn::fail(FENERF_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
        fn::count_launch();                                                                      \
    } while (0)

#define FN_REQUIRE(cond, ...)                                  \
    do {                                                      \
        if (!(cond)) return fn::fail(FENERF_E_ARG, __VA_ARGS__); \
    } while (0)

constexpr int kMaxDevices = 64;

inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// SM count of the CURRENT device (cached per device: one process may render on several GPUs)
inline int num_sms() {
    static std::atomic<int> cached[kMaxDevices];
    const int dev = current_device();
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n <= 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): remember the largest value set on each
// device (`state` is one array per kernel instantiation) and raise it when a launch needs more.
template <typename F>
inline cudaError_t ensure_dynamic_smem(F kernel, std::atomic<int>* state /*[kMaxDevices]*/, int bytes) {
    const int dev = current_device();
    if (state[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) state[dev].store(bytes, std::memory_order_release);
    return e;
}

// ---- entry points of the individual translation units (called by abi.cu) ----
int pack_field(const fenerf_field_desc* f, const FnLayout& L, const fenerf_field_params* p, void* packed,
               cudaStream_t st);
int siren_points_exact(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                       const float* film, int batch, long long ppb, int dir_group, int lock_dirs,
                       const int32_t* only_idx, int n_only, float* out, cudaStream_t st, int sigma_only = 0);
int field_fingerprint(const FnLayout& L, const fenerf_field_params* p, unsigned long long* out, cudaStream_t st);
void set_fast_trace(long long* buf);
long long* get_fast_trace();
int siren_points_fast3(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                       const float* film, int batch, long long ppb, int dir_group, int lock_dirs, float* out,
                       long long* trace, int sigma_only, cudaStream_t st, float* sigma_out = nullptr);
int guard_refine(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                 const float* film, int batch, long long rays_per_batch, int num_steps, int lock_dirs, float tau,
                 const float* noise_far, long long noise_stride, float noise_std,
                 float* raw, int32_t* scratch_idx, int32_t* stats, cudaStream_t st);
int camera_poses(int n, int mode, float h_std, float v_std, float h_mean, float v_mean, const float* draw_theta,
                 const float* draw_phi, float* c2w, float* pitch, float* yaw, cudaStream_t st);
int ray_setup(const fenerf_render_desc* rd, const float* x_lin, const float* y_lin, const float* z_lin,
              const float* cam2world, const float* rng_perturb, float* points, float* z_vals, float* dirs,
              float* origins, cudaStream_t st);
int resample(const fenerf_render_desc* rd, int C, const float* raw, const float* z, const float* dirs,
             const float* origins, const float* noise, const float* u, float* z_fine, float* pts_fine,
             long long* inds, cudaStream_t st, int sort_fine = 0, const float* sigma_compact = nullptr);
int composite_sorted(const fenerf_render_desc* rd, int C, const float* raw_c, const float* z_c, const float* raw_f,
                     const float* z_f, const float* noise, float* pixels, float* depth, float* wsum, float* weights,
                     cudaStream_t st);
int composite(const fenerf_render_desc* rd, int C, const float* raw_c, const float* z_c, const float* raw_f,
              const float* z_f, const float* noise, float* pixels, float* depth, float* wsum, float* weights,
              int32_t* sort_idx, cudaStream_t st);

// gemm5.cu
int gemm_nt(const void* A, const void* B, long long M, float* c32, void* c16, void* a_out, void* gate_out, const float* bias,
            const float* film, long long film_stride, long long ppb, cudaStream_t st, const void* gate_mul = nullptr,
            const void* A2 = nullptr, const void* B2 = nullptr);
int gemm_tn(const void* X, const void* Y, int batch, long long ppb, int slices, float* partial, cudaStream_t st,
            float* colsum = nullptr);
// mapping.cu
int mapping_film(const float* const* w, const float* const* b, const float* z, int B, int z_dim, int n_layers, int layer0,
                 int n_film_total, const float* avg_f, const float* avg_p, float psi, float* h_scratch, float* film,
                 cudaStream_t st);
// frames.cu
int mask2color(const float* masks, int B, int K, long long HW, float* out, cudaStream_t st);
int frames_to_u8(const float* frames, int B, int C, int c0, int nc, long long HW, unsigned char* out, cudaStream_t st);
// backward.cu
int composite_backward(const fenerf_render_desc* rd, int C, const float* raw_c, const float* z_c, const float* raw_f,
                       const float* z_f, const float* noise, const float* d_pixels, float* d_raw_c, float* d_raw_f,
                       cudaStream_t st);
int film_forward_stash(const float* z, const float* bias, const float* film_layer, long long film_batch_stride, long long P,
                       long long ppb, const float* xin, int kx, const float* wx, void* a_out, void* gate_out, int f32,
                       cudaStream_t st);
int gate_backward(void* dA, const void* gate, long long P, long long ppb, float* colsum, int f32, cudaStream_t st);
int head_grads(const float* d_raw, const float* raw, long long P, int C, int L, const float* scale, void* dH, void* dRGB,
               int f32, cudaStream_t st);
int extras_gather(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs, long long P,
                  long long ppb, int dir_group, int lock_dirs, float* out, cudaStream_t st);
int grid_scatter_add(const FnLayout& L, const float* points, const void* d_feat, int ld, long long P, float* grad_cl,
                     int f32, cudaStream_t st);
int grid_unpack_grad(const FnLayout& L, const float* grad_cl, float* out, const float* inv_scale, cudaStream_t st);

}  // namespace fn
