// Microbenchmark 3: sin.approx.f32 (FMUL.RZ + MUFU.SIN) and cvt.rn.f16x2.f32 (F2FP) issue cost per warp instruction with
// 1, 2, 4 warps per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mufu_bench.cu -o tools/mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
template <int MODE>
__global__ void bench(int iters, long long* out, float* sink, float seed) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + 0.01f * (float)(i + lane);
    uint32_t pk = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) x[i] = __sinf(x[i]);                                  // FMUL.RZ + MUFU.SIN
            if (MODE == 1) x[i] = __sinf(fmaf(x[i], 31.f, 0.5f));                // the epilogue's FFMA + FMUL.RZ + MUFU.SIN
            if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));   // MUFU.EX2 alone
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                __half2 h = __floats2half2_rn(x[i], x[i + 1]);
                pk ^= *reinterpret_cast<uint32_t*>(&h);
                x[i] += 1.0f; x[i + 1] += 1.0f;
            }
        }
        if (MODE == 4) {       // the whole epilogue element: FFMA, sin, pack
            uint32_t p[8];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                __half2 h = __floats2half2_rn(__sinf(fmaf(x[i], 31.f, 0.5f)), __sinf(fmaf(x[i + 1], 31.f, 0.5f)));
                p[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { pk ^= p[i]; x[2 * i] += 0.37f; x[2 * i + 1] += 0.11f; }
        }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    sink[threadIdx.x] = s + __uint_as_float(pk);
}
template <int MODE>
int run(const char* name, double instr_per_iter, long long* d_out, float* d_sink) {
    for (int wpq = 1; wpq <= 4; wpq *= 2) {
        const int iters = 4000;
        bench<MODE><<<1, 128 * wpq>>>(iters, d_out, d_sink, 0.3f);
        CK(cudaDeviceSynchronize());
        long long h[16];
        CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
        double per = (double)h[0] / iters;
        printf("%-34s warps/sub-partition %d : %7.1f cycles per 16 elements per warp -> %5.2f cycles per element per sub-partition (%4.1f lanes/clk/SM)\n",
               name, wpq, per, per / 16.0 / wpq, 4.0 * 32.0 * 16.0 * wpq / per);
    }
    return 0;
}
int main() {
    long long* d_out; float* d_sink;
    CK(cudaMalloc(&d_out, 16 * 8)); CK(cudaMalloc(&d_sink, 1024 * 4));
    if (run<0>("sin.approx (FMUL.RZ+MUFU.SIN)", 16, d_out, d_sink)) return 1;
    if (run<1>("FFMA + sin.approx", 16, d_out, d_sink)) return 1;
    if (run<2>("ex2.approx (MUFU.EX2)", 16, d_out, d_sink)) return 1;
    if (run<3>("cvt.rn.f16x2.f32 (8 per 16 elems)", 8, d_out, d_sink)) return 1;
    if (run<4>("FFMA + sin + pack (epilogue element)", 16, d_out, d_sink)) return 1;
    return 0;
}
