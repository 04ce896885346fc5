// Microbenchmark: tcgen05.mma kind::f16 SS-mode issue/execute rate from resident shared memory,
// M = 128, N in {64, 128, 256}, K = 16 per instruction, optional concurrent smem store traffic.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t par) {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(par) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) __trap();
    }
}
// smem: A region 64 KB (4 chunks), B region 64 KB. Each "stage" = 4 K-steps on chunk (i % 4).
__global__ void bench(int n, int stages, int store_warps, int epi_mode, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tslot)) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tm = tslot;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0 && lane == 0) {
        uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        uint32_t a0 = s32(smem), b0 = s32(smem) + 65536;
        long long t0 = clock64();
        for (int s = 0; s < stages; ++s) {
            uint32_t a = a0 + (s & 3) * 16384, b = b0 + (s & 3) * 16384;
            for (int k = 0; k < 4; ++k) mma(tm + ((s & 1) ? (uint32_t)n : 0u) % 256u, desc(a + k * 32), desc(b + k * 32), idesc, k > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        long long t1 = clock64();
        mbar_wait(s32(&bar), 0);
        long long t2 = clock64();
        stop = 1;
        out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0;
    } else if (warp >= 1 && warp <= store_warps) {
        unsigned char* scratch = smem + 131072 + warp * 2048;
        int i = 0;
        if (epi_mode == 0) {
            // concurrent 2-byte shared stores only
            while (!stop) {
#pragma unroll
                for (int j = 0; j < 32; ++j) *reinterpret_cast<volatile unsigned short*>(scratch + ((i + j) & 15) * 128 + lane * 2) = (unsigned short)j;
                i += 32;
            }
        } else {
            // FiLM-epilogue look-alike: tcgen05.ld 32 columns (bit 0), sin per element (bit 1), 2-byte stores (bit 2)
            uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 32;   // columns the MMAs do not touch
            float acc = 0.f;
            while (!stop) {
                uint32_t r[32];
                if (epi_mode & 1) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                        : "r"(taddr) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint((float)(i + j));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]);
                    if (epi_mode & 2) v = __sinf(fmaf(v, 1.3f, 0.7f));
                    if (epi_mode & 4) *reinterpret_cast<volatile unsigned short*>(scratch + ((i + j) & 15) * 128 + lane * 2) = (unsigned short)__float_as_uint(v);
                    else acc += v;
                }
                i += 32;
            }
            if (acc == 12345.f) out[1000] = 1;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}
int main() {
    long long* out; CK(cudaMalloc(&out, 4096));
    CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    struct Cfg { int n, w, mode; };
    Cfg cfgs[] = {{64,0,0},{128,0,0},{256,0,0},{128,8,0},{128,8,1},{128,8,2},{128,8,4},{128,8,3},{128,8,7},{128,4,7},{256,8,7}};
    for (auto c : cfgs) {
        int n = c.n, w = c.w;
        int stages = 64;
        bench<<<148, 320, 160 * 1024>>>(n, stages, w, c.mode, out); CK(cudaDeviceSynchronize());
        bench<<<148, 320, 160 * 1024>>>(n, stages, w, c.mode, out); CK(cudaDeviceSynchronize());
        long long h[296]; CK(cudaMemcpy(h, out, 296 * 8, cudaMemcpyDeviceToHost));
        double issue = 0, total = 0; for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
        issue /= 148; total /= 148;
        double per_mma = total / (stages * 4);
        printf("N=%3d warps=%d epi_mode=%d (1=tmem ld,2=sin,4=sts) : issue %.0f cyc, complete %.0f cyc for %d MMAs -> %.1f cyc/MMA (ideal %d), %.0f%% of peak\n", n, w, c.mode, issue, total,
               stages * 4, per_mma, n / 2, 100.0 * (n / 2) / per_mma);
    }
    return 0;
}
