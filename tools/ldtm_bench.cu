// Microbenchmark: tcgen05.ld (32x32b, x16 / x32 / x64) round-trip time and throughput per SM sub-partition, with 1, 2 or 4
// reader warps per TMEM lane quadrant, optionally while one thread streams tcgen05.mma (M = 128, N = 128, K = 16) into TMEM and
// optionally with 16 MUFU.SIN per 16 columns between the loads (the FiLM epilogue's shape).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ldtm_bench.cu -o tools/ldtm_bench && tools/ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int W> struct Ld;
template <> struct Ld<16> { static __device__ __forceinline__ void go(uint32_t t, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(t) : "memory"); } };
template <> struct Ld<32> { static __device__ __forceinline__ void go(uint32_t t, uint32_t* r) { Ld<16>::go(t, r); Ld<16>::go(t + 16, r + 16); } };
template <> struct Ld<64> { static __device__ __forceinline__ void go(uint32_t t, uint32_t* r) { Ld<32>::go(t, r); Ld<32>::go(t + 32, r + 32); } };

// mode: 0 = ld; wait; (consume)   1 = + 1 sin per element, double-buffered like the production epilogue
template <int W>
__global__ void bench(int iters, int readers_per_quadrant, int with_mma, int with_sin, long long* out, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tslot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) stop = 0;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tslot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tslot;
    const int n_readers = 4 * readers_per_quadrant;
    if (warp < n_readers) {
        const int q = warp & 3, j = warp >> 2;
        const uint32_t base = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 128) % 512u;
        uint32_t r[2][W];
        float acc = 0.f;
        __syncwarp();
        const long long t0 = clock64();
        Ld<W>::go(base, r[0]);
        for (int it = 0; it < iters; ++it) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            Ld<W>::go(base + (uint32_t)(((it + 1) * W) % 128), r[(it + 1) & 1]);
#pragma unroll
            for (int i = 0; i < W; ++i) {
                float v = __uint_as_float(r[it & 1][i]);
                acc += with_sin ? __sinf(fmaf(v, 31.f, 0.5f)) : v;
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
        if (warp == 0 && lane == 0) stop = 1;
    } else if (warp == n_readers && with_mma && lane == 0) {
        const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = s32(smem), b0 = s32(smem) + 65536;
        int s = 0;
        long long n = 0;
        const long long t0 = clock64();
        while (!stop) {
            for (int k = 0; k < 4; ++k) mma(tm + 256u + (uint32_t)(s & 1) * 128u, desc(a0 + (s & 3) * 16384 + k * 32), desc(b0 + (s & 3) * 16384 + k * 32), idesc, 1);
            ++s; n += 4;
        }
        out[blockIdx.x * 32 + 30] = n; out[blockIdx.x * 32 + 31] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory"); }
}

template <int W>
int run(int rpq, int with_mma, int with_sin, long long* d_out, float* d_sink) {
    const int iters = 2000;
    const int threads = (4 * rpq + 1) * 32;
    CK(cudaFuncSetAttribute(bench<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    CK(cudaMemset(d_out, 0, 32 * 8));
    bench<W><<<1, threads, 128 * 1024>>>(iters, rpq, with_mma, with_sin, d_out, d_sink);
    CK(cudaDeviceSynchronize());
    long long h[32];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    double per = (double)h[0] / iters;
    printf("x%-3d readers/quadrant %d  mma %d  sin %d : %7.1f cycles per ld round (%5.1f B/clk per quadrant, %5.2f cycles per column)",
           W, rpq, with_mma, with_sin, per, rpq * W * 128.0 / per, per / W);
    if (with_mma) printf("   mma: %.1f cycles each", (double)h[31] / (double)h[30]);
    printf("\n");
    return 0;
}

int main() {
    long long* d_out; float* d_sink;
    CK(cudaMalloc(&d_out, 32 * 8)); CK(cudaMalloc(&d_sink, 1024 * 4));
    for (int sin = 0; sin < 2; ++sin)
        for (int mm = 0; mm < 2; ++mm)
            for (int rpq = 1; rpq <= 4; rpq *= 2) {
                if (run<16>(rpq, mm, sin, d_out, d_sink)) return 1;
                if (run<32>(rpq, mm, sin, d_out, d_sink)) return 1;
                if (rpq <= 2 && run<64>(rpq, mm, sin, d_out, d_sink)) return 1;
            }
    return 0;
}
