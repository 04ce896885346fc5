// Microbenchmark 4: the FiLM epilogue's inner loop in isolation -- tcgen05.ld.32x32b.x16 (double-buffered) -> FFMA -> sin.approx ->
// cvt.rn.f16x2 -> 16-byte swizzled shared stores -- with 1, 2, 4 warps per SM sub-partition and each ingredient switched off in turn.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/epi_bench.cu -o tools/epi_bench
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld16(uint32_t t, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(t) : "memory");
}
__device__ __forceinline__ void fake16(uint32_t (&r)[16], uint32_t seed) {       // opaque register values (no hoisting), no TMEM traffic
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("mov.b32 %0, %1;" : "=r"(r[i]) : "r"(seed + i));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// FLAGS: 1 = tcgen05.ld, 2 = sin, 4 = shared stores
template <int FLAGS>
__global__ void bench(int iters, long long* out, float f_h, float p_h) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tslot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    const int q = warp & 3, jw = warp >> 2;
    const int fl = q * 32 + lane;
    const uint32_t t_lane = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(jw & 3) * 128u;
    const uint32_t kk = (uint32_t)(fl & 63);
    unsigned char* rowp = smem + (uint32_t)(jw & 3) * 32768u + (uint32_t)(fl >> 6) * 16384u + (kk >> 3) * 2048u + (kk & 7u) * 128u;
    const uint32_t sw = kk & 7u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[2][16];
        if (FLAGS & 1) ld16(t_lane, r[0]); else fake16(r[0], (uint32_t)it);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            if (FLAGS & 1) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g < 7) ld16(t_lane + (g + 1) * 16, r[(g + 1) & 1]);
            } else if (g < 7) fake16(r[(g + 1) & 1], (uint32_t)(it + g));
#pragma unroll
            for (int j8 = 0; j8 < 2; ++j8) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float u = fmaf(f_h, __uint_as_float(r[g & 1][j8 * 8 + i]), p_h);
                    v[i] = (FLAGS & 2) ? __sinf(u) : u;
                }
                uint4 pk;
                pk.x = pack2(v[0], v[1]); pk.y = pack2(v[2], v[3]); pk.z = pack2(v[4], v[5]); pk.w = pack2(v[6], v[7]);
                const uint32_t pt8 = (uint32_t)(g * 2 + j8);
                if ((FLAGS & 4) || pk.x == 0x12345678u)
                    *reinterpret_cast<uint4*>(rowp + (pt8 >> 3) * 1024u + (((pt8 & 7u) ^ sw) << 4)) = pk;
            }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory"); }
}
template <int FLAGS>
int run(const char* name, long long* d_out) {
    CK(cudaFuncSetAttribute(bench<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    for (int wpq = 1; wpq <= 4; wpq *= 2) {
        const int iters = 1000;
        bench<FLAGS><<<1, 128 * wpq, 128 * 1024>>>(iters, d_out, 31.f, 0.5f);
        CK(cudaDeviceSynchronize());
        long long h[16];
        CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
        double per = (double)h[0] / iters / 128.0;
        printf("%-28s warps/sub-partition %d : %6.2f cycles per element per warp -> %5.2f per element per sub-partition\n", name, wpq, per, per / wpq);
    }
    return 0;
}
int main() {
    long long* d_out;
    CK(cudaMalloc(&d_out, 16 * 8));
    if (run<7>("ld + sin + sts (production)", d_out)) return 1;
    if (run<6>("     sin + sts", d_out)) return 1;
    if (run<3>("ld + sin", d_out)) return 1;
    if (run<5>("ld       + sts", d_out)) return 1;
    if (run<2>("     sin", d_out)) return 1;
    if (run<1>("ld", d_out)) return 1;
    return 0;
}
