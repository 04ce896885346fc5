// Microbenchmark: how fast can one SM pull L2-resident data into shared memory?
//   mode 0: cp.async.bulk (1-D), `nissue` issuing warps (one elected lane each), each with its own
//           ring of `depth` slots of `sz` bytes
//   mode 1: cp.async (LDGSTS 16 B) by all threads
//   mode 2: plain LDG.128 -> STS.128 by all threads
// Prints bytes/clk/SM for each configuration. L2-resident source (2 MB, re-read).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(par) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void bench(const unsigned char* __restrict__ src, size_t src_bytes, int mode, int sz, int depth, int nissue, int iters,
                      long long* out_cycles, unsigned long long* out_bytes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(s32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long t0 = clock64();
    unsigned long long moved = 0;
    if (mode == 0) {
        if (warp < nissue && lane == 0) {
            unsigned char* ring = smem + (size_t)warp * depth * sz;
            uint32_t bar0 = s32(&bars[warp * 8]);
            size_t off = ((size_t)blockIdx.x * 131 + warp * 17) * 4096 % src_bytes;
            // prime
            for (int i = 0; i < depth; ++i) { mbar_expect(bar0 + 8 * i, sz); bulk(s32(ring + (size_t)i * sz), src + off, sz, bar0 + 8 * i); off = (off + sz) % (src_bytes - sz); off &= ~(size_t)15; }
            for (int it = 0; it < iters; ++it) {
                int slot = it % depth; uint32_t ph = (it / depth) & 1;
                mbar_wait(bar0 + 8 * slot, ph);
                moved += sz;
                mbar_expect(bar0 + 8 * slot, sz); bulk(s32(ring + (size_t)slot * sz), src + off, sz, bar0 + 8 * slot);
                off = (off + sz) % (src_bytes - sz); off &= ~(size_t)15;
            }
            for (int i = 0; i < depth; ++i) { int it = iters + i; mbar_wait(bar0 + 8 * (it % depth), (it / depth) & 1); }
        }
    } else if (mode == 1) {
        size_t off = ((size_t)blockIdx.x * 131) * 4096 % src_bytes;
        for (int it = 0; it < iters; ++it) {
            for (int i = threadIdx.x * 16; i < sz; i += blockDim.x * 16) {
                uint32_t d = s32(smem + (it % depth) * (size_t)sz + i);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + off + i));
            }
            asm volatile("cp.async.commit_group;" ::);
            if (it >= depth - 1) asm volatile("cp.async.wait_group %0;" ::"n"(3));
            off = (off + sz) % (src_bytes - sz); off &= ~(size_t)15;
            if (threadIdx.x == 0) moved += sz;
        }
        asm volatile("cp.async.wait_group 0;" ::);
    } else {
        size_t off = ((size_t)blockIdx.x * 131) * 4096 % src_bytes;
        for (int it = 0; it < iters; ++it) {
            for (int i = threadIdx.x * 16; i < sz; i += blockDim.x * 16) {
                uint4 v = *reinterpret_cast<const uint4*>(src + off + i);
                *reinterpret_cast<uint4*>(smem + (it % depth) * (size_t)sz + i) = v;
            }
            off = (off + sz) % (src_bytes - sz); off &= ~(size_t)15;
            if (threadIdx.x == 0) moved += sz;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (lane == 0 && (warp < nissue || mode != 0)) atomicAdd(&out_bytes[blockIdx.x], moved);
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
}

int main() {
    size_t src_bytes = 2u << 20;
    unsigned char* src; CK(cudaMalloc(&src, src_bytes)); CK(cudaMemset(src, 1, src_bytes));
    long long* cyc; unsigned long long* byt; CK(cudaMalloc(&cyc, 1024 * 8)); CK(cudaMalloc(&byt, 1024 * 8));
    CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int sms = 148;
    struct Cfg { int mode, sz, depth, nissue, ctas_per_sm; };
    Cfg cfgs[] = {
        {0, 16384, 1, 1, 1}, {0, 16384, 2, 1, 1}, {0, 16384, 4, 1, 1}, {0, 16384, 8, 1, 1},
        {0, 4096, 4, 1, 1}, {0, 4096, 16, 1, 1}, {0, 2048, 32, 1, 1}, {0, 32768, 4, 1, 1},
        {0, 16384, 2, 2, 1}, {0, 16384, 2, 4, 1}, {0, 8192, 4, 4, 1}, {0, 4096, 4, 8, 1},
        {0, 16384, 2, 1, 2}, {0, 16384, 4, 1, 2},
        {1, 16384, 4, 1, 1}, {2, 16384, 4, 1, 1},
    };
    for (auto c : cfgs) {
        int grid = sms * c.ctas_per_sm;
        size_t smem = c.mode == 0 ? (size_t)c.nissue * c.depth * c.sz : (size_t)c.depth * c.sz;
        if (c.ctas_per_sm == 2 && smem > 100 * 1024) continue;
        int iters = 400;
        CK(cudaMemset(byt, 0, 1024 * 8));
        bench<<<grid, 256, smem>>>(src, src_bytes, c.mode, c.sz, c.depth, c.nissue, 50, cyc, byt);   // warm
        CK(cudaDeviceSynchronize());
        CK(cudaMemset(byt, 0, 1024 * 8));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        bench<<<grid, 256, smem>>>(src, src_bytes, c.mode, c.sz, c.depth, c.nissue, iters, cyc, byt);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        static long long hc[1024]; static unsigned long long hb[1024];
        CK(cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hb, byt, grid * 8, cudaMemcpyDeviceToHost));
        double tot = 0, cycs = 0; for (int i = 0; i < grid; ++i) { tot += hb[i]; cycs += hc[i]; }
        double per_cta = tot / cycs;   // bytes per clk per CTA (avg)
        printf("mode %d sz %6d depth %2d issuers %d ctas/SM %d : %6.1f B/clk/CTA  %6.1f B/clk/SM  chip %.2f TB/s (%.3f ms)\n", c.mode, c.sz, c.depth,
               c.nissue, c.ctas_per_sm, per_cta, per_cta * c.ctas_per_sm, tot / (ms * 1e-3) / 1e12, ms);
    }
    return 0;
}
