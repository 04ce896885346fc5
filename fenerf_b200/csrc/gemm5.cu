// tcgen05 GEMMs of the backward: the three 256-wide products per FiLM layer.
//
//   gemm_nt_kernel     C[M, 256] = A[M, 256] . B[256, 256]^T        A, B fp16 row-major (K contiguous)
//                        recompute   z = a W^T   -> fp32, or with the FiLM epilogue fused: a' = sin(f (z + b) + p) and
//                                                   the gate f cos(.) written as fp16, z never leaves the SM
//                        backward    dA' = dZ W  -> fp16 (B = W^T, transposed once per call by the host)
//   gemm_tn_kernel     C_b[256, 256] = sum over the points of image b of  X[p, :]^T  Y[p, :]   (split-K over CTAs)
//                        dW_b = dZ^T a: both operands are MN-major views of the row-major (P, 256) streams
//
// One persistent CTA per SM, 192 threads: warps 0..3 epilogue (one TMEM lane quadrant each), warp 4 loader
// (16-byte cp.async into the 128B-swizzled UMMA layouts -- the operands are plain row-major tensors, no tensor maps),
// warp 5 MMA issuer (elect.sync).  gemm_nt keeps the whole 256 x 256 fp16 B matrix resident in shared memory (128 KB)
// (+ 32 KB for an optional fifth, narrow k-chunk) and streams A through a 3-deep ring of 64-wide k-chunks ([128 rows][64 k],
// 16 KB each); the accumulator is
// double-buffered in TMEM (2 x 256 columns), so the loads and MMAs of tile i+1 overlap the epilogue of tile i.  It is
// HBM-bound: 16.8 MFLOP per tile against >= 128 KB moved, ~130 FLOP/B vs the machine's 260.  gemm_tn streams 64-point
// stages (2 x 32 KB) through a 3-deep ring and holds the 256 x 256 fp32 result in all 512 TMEM columns.  The loader never
// waits for its own copies: cp.async.mbarrier.arrive.noinc hands each lane's arrival to the stage's mbarrier.
#include "common.cuh"
#include "tc5.cuh"

namespace fn {

namespace {

using namespace tc5;

constexpr int kThreads = 192;
constexpr int kLoadWarp = 4, kMmaWarp = 5;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// the mbarrier receives this thread's arrival once all of its earlier cp.async have landed (no wait in the loader: the
// ring stays full); the consumer runs fence.proxy.async after its wait, before handing the buffer to the tensor core
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_shared_zero16(uint32_t dst) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
}
// MN-major 128B-swizzled operand laid out [k/8][mn/64][k%8][64 mn]: LBO (between 64-element MN atoms) and SBO (between
// 8-row K groups) in bytes
__device__ __forceinline__ uint64_t desc_hi_mn(uint32_t lbo, uint32_t sbo) {
    return ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// ------------------------------------------------------------------------------------------------------------------
struct NtArgs {
    const __half* A;      // (M, 256)
    const __half* B;      // (256, 256): C = A B^T
    float* C32;           // (M, 256) fp32 out, or
    __half* C16;          // (M, 256) fp16 out, or (FiLM epilogue) both of:
    __half* a_out;        // (M, 256) sin(f (c + bias) + p)
    __half* gate_out;     // (M, 256) f cos(f (c + bias) + p)
    const __half* gate_mul;  // optional (M, 256): the fp16 output is multiplied by it (dZ' = (dZ W) * gate of the layer below)
    const __half* A2;     // optional fifth k-chunk: narrow inputs (M, 64) fp16 (zero padded) ...
    const __half* B2;     // ... against (256, 64) fp16: C += A2 B2^T  (the first colour layer's [dir, grid features])
    const float* bias;    // (256)
    const float* film;    // image 0's [2][256] block of the layer
    long long film_stride, ppb;
    long long M;
};

constexpr uint32_t NT_SB = 0;                 // B: 5 k-chunks of [256 rows][64 k] = 5 x 32 KB (the fifth only with narrow inputs)
constexpr uint32_t NT_SA = 163840;            // A ring: NT_RING k-chunks of [128 rows][64 k], 16 KB each
constexpr int NT_RING = 3;
constexpr uint32_t NT_BAR = NT_SA + NT_RING * 16384;    // barriers + tmem slot
constexpr uint32_t NT_FILM = NT_BAR + 128;    // [3][256] floats: f, p, bias of the tile's first image
constexpr uint32_t NT_SMEM = NT_FILM + 3 * 256 * 4;

__global__ void __launch_bounds__(kThreads, 1) gemm_nt_kernel(const __grid_constant__ NtArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    // warp index by lane-0 broadcast: descriptor / barrier arithmetic of the issuer then stays in uniform registers (see
    // siren_fast3.cu); every wait below is therefore the vote-terminated form, which leaves the warp converged
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t bar_b = sbase + NT_BAR;
    const uint32_t bar_accfull = bar_b + 8 /* [2] */, bar_accempty = bar_b + 24 /* [2] */;
    const uint32_t bar_afull = bar_b + 40 /* [NT_RING] */, bar_aempty = bar_afull + 8 * NT_RING /* [NT_RING] */;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + NT_BAR + 120);
    float* s_film = reinterpret_cast<float*>(smem + NT_FILM);
    if (threadIdx.x == 0) {
        mbar_init(bar_b, 32);
        for (int i = 0; i < NT_RING; ++i) { mbar_init(bar_afull + 8 * i, 32); mbar_init(bar_aempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar_accfull + 8 * i, 1); mbar_init(bar_accempty + 8 * i, 4); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long n_tiles = (a.M + 127) / 128;
    const int n_chunks = a.A2 ? 5 : 4;

    if (warp == kLoadWarp) {
        // ---- B once: (row, kc, piece) -> kc * 32 KB + sw128(row, piece * 8)
        for (int i = lane; i < 256 * 32; i += 32) {
            const int row = i >> 5, kc = (i >> 3) & 3, j = i & 7;
            cp_async16(sbase + NT_SB + kc * 32768 + fn_sw128_offset(row, j * 8), a.B + row * 256 + kc * 64 + j * 8);
        }
        if (a.A2)
            for (int i = lane; i < 256 * 8; i += 32) {
                const int row = i >> 3, j = i & 7;
                cp_async16(sbase + NT_SB + 4 * 32768 + fn_sw128_offset(row, j * 8), a.B2 + row * 64 + j * 8);
            }
        cp_async_arrive_noinc(bar_b);
        uint32_t ch = 0;                                   // running k-chunk number: ring slot ch % NT_RING
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const long long m0 = t * 128;
            for (int kc = 0; kc < n_chunks; ++kc, ++ch) {
                const uint32_t slot = ch % NT_RING, use = ch / NT_RING;
                mbar_wait_warp_spin(bar_aempty + 8 * slot, (use & 1) ^ 1);
                const uint32_t base = sbase + NT_SA + slot * 16384;
                const __half* src = kc < 4 ? a.A + kc * 64 : a.A2;
                const long long ld = kc < 4 ? 256 : 64;
                for (int i = lane; i < 128 * 8; i += 32) {           // (row, 16-byte piece) of this 64-wide k-chunk
                    const int row = i >> 3, j = i & 7;
                    const uint32_t dst = base + fn_sw128_offset(row, j * 8);
                    if (m0 + row < a.M) cp_async16(dst, src + (m0 + row) * ld + j * 8);
                    else st_shared_zero16(dst);
                }
                cp_async_arrive_noinc(bar_afull + 8 * slot);
            }
        }
    } else if (warp == kMmaWarp) {
        mbar_wait_warp_spin(bar_b, 0);
        fence_async_smem();
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_f16(256, 0, 0);
        const uint32_t a_lo = (sbase + NT_SA) >> 4, b_lo = (sbase + NT_SB) >> 4;
        uint32_t it = 0, ch = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait_warp_spin(bar_accempty + 8 * buf, (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + buf * 256u;
            for (int kc = 0; kc < n_chunks; ++kc, ++ch) {
                const uint32_t slot = ch % NT_RING, cuse = ch / NT_RING;
                mbar_wait_warp_spin(bar_afull + 8 * slot, cuse & 1);
                fence_async_smem();          // the chunk was written by cp.async (generic proxy): order it before the MMA's reads
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    tc_mma_f16_elect(d, kDescHi | (uint64_t)(a_lo + slot * (16384 >> 4) + 2 * k),
                                     kDescHi | (uint64_t)(b_lo + kc * (32768 >> 4) + 2 * k), idesc, (kc | k) ? 1u : 0u);
                tc_commit_elect(bar_aempty + 8 * slot);
            }
            tc_commit_elect(bar_accfull + 8 * buf);
        }
    } else {
        // ---- epilogue: thread = row of the tile (TMEM lane), 256 columns in 8 groups of 32
        const int q = warp & 3, row = q * 32 + lane;
        const bool film_mode = a.a_out != nullptr;
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            const long long m = t * 128 + row;
            const bool valid = m < a.M;
            long long img0 = 0, img = 0;
            if (film_mode) {
                img0 = (t * 128) / a.ppb;
                img = valid ? m / a.ppb : img0;
                // the four epilogue warps stage the FiLM rows of the tile's first image (named barrier 1, 128 threads)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int i = threadIdx.x; i < 256; i += 128) {
                    s_film[i] = a.film[img0 * a.film_stride + i];
                    s_film[256 + i] = a.film[img0 * a.film_stride + 256 + i];
                    s_film[512 + i] = a.bias[i];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait_warp_spin(bar_accfull + 8 * buf, use & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u;
#pragma unroll 1
            for (int g = 0; g < 8; ++g) {
                uint32_t r[32];
                tc_ld32(taddr + g * 32, r);
                tc_wait_ld();
                if (!valid) continue;
                if (film_mode) {
                    const float* fl = (img == img0) ? nullptr : a.film + img * a.film_stride;
                    __align__(16) __half av[32], gv[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int c = g * 32 + j;
                        const float fr = fl ? __ldg(fl + c) : s_film[c], ph = fl ? __ldg(fl + 256 + c) : s_film[256 + c];
                        const float u = fmaf(fr, __uint_as_float(r[j]) + s_film[512 + c], ph);
                        float sn, cs;
                        __sincosf(u, &sn, &cs);          // MUFU: ~|u| * 2^-24 absolute, far inside the fp16 streams' rounding
                        av[j] = __float2half_rn(sn);
                        gv[j] = __float2half_rn(fr * cs);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        reinterpret_cast<uint4*>(a.a_out + m * 256 + g * 32)[j] = reinterpret_cast<const uint4*>(av)[j];
                        reinterpret_cast<uint4*>(a.gate_out + m * 256 + g * 32)[j] = reinterpret_cast<const uint4*>(gv)[j];
                    }
                } else if (a.C32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        reinterpret_cast<float4*>(a.C32 + m * 256 + g * 32)[j] =
                            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                        __uint_as_float(r[4 * j + 3]));
                } else {
                    if (a.gate_mul) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 gq = *reinterpret_cast<const uint4*>(a.gate_mul + m * 256 + g * 32 + j * 8);
                            const __half2* g2 = reinterpret_cast<const __half2*>(&gq);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 gf = __half22float2(g2[i]);
                                r[8 * j + 2 * i] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i]) * gf.x);
                                r[8 * j + 2 * i + 1] = __float_as_uint(__uint_as_float(r[8 * j + 2 * i + 1]) * gf.y);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 pk;
                        pk.x = pack_half2(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
                        pk.y = pack_half2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
                        pk.z = pack_half2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
                        pk.w = pack_half2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                        reinterpret_cast<uint4*>(a.C16 + m * 256 + g * 32)[j] = pk;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct TnArgs {
    const __half* X;      // (B * ppb, 256): rows of image b are [b * ppb, (b + 1) * ppb)
    const __half* Y;      // (B * ppb, 256)
    float* partial;       // (B, slices, 256, 256): X_b^T Y_b summed over this CTA's stages
    float* colsum;        // optional (B, slices, 256): column sums of X over the same stages (the bias / phase gradients)
    long long ppb;
    int slices;
};

constexpr int TN_STAGES = 3;
constexpr uint32_t TN_STAGE_BYTES = 65536;          // X half-tile 32 KB + Y half-tile 32 KB (64 points each)
constexpr uint32_t TN_BAR = TN_STAGES * TN_STAGE_BYTES;
constexpr uint32_t TN_SMEM = TN_BAR + 128;

__global__ void __launch_bounds__(kThreads, 1) gemm_tn_kernel(const __grid_constant__ TnArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    // warp index by lane-0 broadcast: descriptor / barrier arithmetic of the issuer then stays in uniform registers (see
    // siren_fast3.cu); every wait below is therefore the vote-terminated form, which leaves the warp converged
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + TN_BAR /* [3] */, bar_empty = bar_full + 24 /* [3] */, bar_done = bar_full + 48;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TN_BAR + 64);
    if (threadIdx.x == 0) {
        // a stage is released by the MMA's commit and, when column sums are wanted, by the four epilogue warps reading it
        for (int i = 0; i < TN_STAGES; ++i) { mbar_init(bar_full + 8 * i, 32); mbar_init(bar_empty + 8 * i, a.colsum ? 5 : 1); }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int slice = blockIdx.x, b = blockIdx.y;
    const long long n_stage_total = (a.ppb + 63) / 64;                       // 64-point stages of this image
    const long long n_mine = (n_stage_total - slice + a.slices - 1) / a.slices;   // stages slice, slice + slices, ...
    const long long row0 = (long long)b * a.ppb;

    if (warp == kLoadWarp) {
        for (long long i = 0; i < n_mine; ++i) {
            const uint32_t st = (uint32_t)(i % TN_STAGES), use = (uint32_t)(i / TN_STAGES);
            mbar_wait_warp_spin(bar_empty + 8 * st, (use & 1) ^ 1);
            const long long p0 = (slice + i * a.slices) * 64;
            const uint32_t sx = sbase + st * TN_STAGE_BYTES, sy = sx + 32768;
            // row k of the stage = point p0 + k: four 128-byte segments (64 features each) -> [k/8][seg][k%8][64]
            for (int idx = lane; idx < 64 * 32; idx += 32) {
                const int k = idx >> 5, seg = (idx >> 3) & 3, j = idx & 7;
                const uint32_t off = (uint32_t)(k >> 3) * 4096u + (uint32_t)seg * 1024u + (uint32_t)(k & 7) * 128u + (uint32_t)((j ^ (k & 7)) << 4);
                if (p0 + k < a.ppb) {
                    const long long g = (row0 + p0 + k) * 256 + seg * 64 + j * 8;
                    cp_async16(sx + off, a.X + g);
                    cp_async16(sy + off, a.Y + g);
                } else {
                    st_shared_zero16(sx + off);
                    st_shared_zero16(sy + off);
                }
            }
            cp_async_arrive_noinc(bar_full + 8 * st);
        }
    } else if (warp == kMmaWarp) {
        constexpr uint32_t idesc = umma_idesc_f16(256, 1, 1);
        const uint64_t hi = desc_hi_mn(1024, 4096);
        for (long long i = 0; i < n_mine; ++i) {
            const uint32_t st = (uint32_t)(i % TN_STAGES), use = (uint32_t)(i / TN_STAGES);
            mbar_wait_warp_spin(bar_full + 8 * st, use & 1);
            fence_async_smem();
            tc_fence_after();
            const uint32_t x_lo = (sbase + st * TN_STAGE_BYTES) >> 4, y_lo = x_lo + (32768 >> 4);
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    tc_mma_f16_elect(tmem_base + h * 256u, hi | (uint64_t)(x_lo + h * (2048 >> 4) + ks * (8192 >> 4)),
                                     hi | (uint64_t)(y_lo + ks * (8192 >> 4)), idesc, (i > 0 || ks > 0) ? 1u : 0u);
            tc_commit_elect(bar_empty + 8 * st);
        }
        tc_commit_elect(bar_done);
    } else {
        const int q = warp & 3, row = q * 32 + lane;
        if (a.colsum) {
            // while the tensor core works: column sums of X straight from the staged tiles.  Thread t owns features 2t,
            // 2t+1; element (k, f) of a stage sits at (k/8)*4096 + (f/64)*1024 + (k%8)*128 + (((f%64)/8) ^ (k%8))*16 + (f%8)*2
            const int t = threadIdx.x, f = 2 * t;
            float s0 = 0.f, s1 = 0.f;
            for (long long i = 0; i < n_mine; ++i) {
                const uint32_t st = (uint32_t)(i % TN_STAGES), use = (uint32_t)(i / TN_STAGES);
                mbar_wait_warp_spin(bar_full + 8 * st, use & 1);
                const unsigned char* sx = smem + st * TN_STAGE_BYTES + (f >> 6) * 1024 + (f & 7) * 2;
#pragma unroll 8
                for (int k = 0; k < 64; ++k) {
                    const __half2 v = *reinterpret_cast<const __half2*>(sx + (k >> 3) * 4096 + (k & 7) * 128 + ((((f & 63) >> 3) ^ (k & 7)) << 4));
                    const float2 vf = __half22float2(v);
                    s0 += vf.x; s1 += vf.y;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8 * st);
            }
            float* cs = a.colsum + ((size_t)b * a.slices + slice) * 256;
            cs[f] = s0; cs[f + 1] = s1;
        }
        mbar_wait_warp_spin(bar_done, 0);
        tc_fence_after();
        float* out = a.partial + ((size_t)b * a.slices + slice) * 65536;
#pragma unroll 1
        for (int h = 0; h < 2; ++h)
#pragma unroll 1
            for (int g = 0; g < 8; ++g) {
                uint32_t r[32];
                if (n_mine > 0) {
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256u + g * 32, r);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<float4*>(out + (size_t)(h * 128 + row) * 256 + g * 32)[j] =
                        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                    __uint_as_float(r[4 * j + 3]));
            }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace

int gemm_nt(const void* A, const void* B, long long M, float* c32, void* c16, void* a_out, void* gate_out, const float* bias,
            const float* film, long long film_stride, long long ppb, cudaStream_t st, const void* gate_mul, const void* A2,
            const void* B2) {
    static_assert(NT_SMEM <= 232448, "gemm_nt shared memory");
    NtArgs a;
    a.A = (const __half*)A; a.B = (const __half*)B; a.C32 = c32; a.C16 = (__half*)c16; a.a_out = (__half*)a_out;
    a.gate_out = (__half*)gate_out; a.gate_mul = (const __half*)gate_mul; a.A2 = (const __half*)A2; a.B2 = (const __half*)B2;
    a.bias = bias; a.film = film; a.film_stride = film_stride; a.ppb = ppb > 0 ? ppb : 1; a.M = M;
    if (M <= 0) return 0;
    static std::atomic<int> set[kMaxDevices];
    FN_CUDA_OK(ensure_dynamic_smem(gemm_nt_kernel, set, (int)NT_SMEM));
    const long long tiles = (M + 127) / 128;
    const int blocks = (int)(tiles < (long long)num_sms() ? tiles : (long long)num_sms());
    gemm_nt_kernel<<<blocks, kThreads, NT_SMEM, st>>>(a);
    FN_LAUNCH_OK("gemm_nt_kernel");
    return 0;
}

int gemm_tn(const void* X, const void* Y, int batch, long long ppb, int slices, float* partial, cudaStream_t st, float* colsum) {
    static_assert(TN_SMEM <= 232448, "gemm_tn shared memory");
    TnArgs a;
    a.X = (const __half*)X; a.Y = (const __half*)Y; a.partial = partial; a.colsum = colsum; a.ppb = ppb; a.slices = slices;
    static std::atomic<int> set[kMaxDevices];
    FN_CUDA_OK(ensure_dynamic_smem(gemm_tn_kernel, set, (int)TN_SMEM));
    gemm_tn_kernel<<<dim3(slices, batch), kThreads, TN_SMEM, st>>>(a);
    FN_LAUNCH_OK("gemm_tn_kernel");
    return 0;
}

}  // namespace fn
