// Point network, EXACT mode: the whole FiLM-SIREN evaluated per 64-point tile in fp32 on the CUDA
// cores, activations never leaving shared memory.
//
// Replaces <SIREN>.forward_with_frequencies_phase_shifts (siren/siren.py:164-178 for TALLSIREN,
// :1509-1530 for the texture-embedding field) -- in the reference one addmm + three elementwise
// passes per FiLM layer, each materialising a (P, 256) fp32 tensor in memory.
//
// Role on B200: this is the fp32-faithful engine.  It serves FENERF_PRECISION_EXACT, and in the
// default GUARD mode it re-evaluates the handful of far samples whose sigma sits next to the
// reference's relu(sigma) * 1e10 step (volumetric_rendering.py:24,32); the bulk of the points go
// through the tcgen05 kernel in siren_fast.cu.  FP32-pipe bound: 2 * 256 * 256 FLOP per layer per
// point on 128 FMA lanes / SM.
//
// Tile structure (256 threads, 2 CTAs / SM):
//   A    [304][64] f32  activations, k-major (row k = feature k of the 64 points); rows 256.. hold
//                       the first colour layer's extra inputs dir(3), grid_feat(G), zero pad
//   Wbuf [2][16][256]   double-buffered 16-row slabs of the k-major weights via cp.async
//   each thread owns a 4-point x 16-column micro tile; 5 LDS.128 feed 64 FFMA per k
//   (the 16-point gather variant maps lanes point-major so a warp's weight reads broadcast)
#include "common.cuh"

namespace fn {

namespace {

constexpr int KA = 304;          // 256 + max padded extras (3 + 32 -> 48)
constexpr int KC = 16;           // weight rows per pipeline slab
constexpr int NTHREADS = 256;

struct ExactArgs {
    FnLayout L;
    const unsigned char* packed;
    const float* points;
    const float* dirs;
    const float* film;
    const int32_t* only_idx;
    const int32_t* n_only_dev;   // gather mode: entry count lives on the device (GUARD refinement)
    float* out;
    long long ppb;       // points per batch element
    long long n_items;   // tiles cover: batch * tiles_per_batch (dense) or ceil(n_only / 64) (gather)
    long long tiles_per_batch;
    int n_only;
    int dir_group;
    int lock_dirs;
    int sigma_only;      // stop after the density head: out[..., C-1] only (GUARD refinement, density grids)
    int32_t* guard_stats;  // GUARD self-check (fenerf_b200.h: fenerf_guard_stats) or NULL
};

// P = points per thread (4: dense 64-point tiles; 1: 16-point tiles for the sparse GUARD gather, where
// a tile's latency matters more than FMA efficiency)
template <int P>
struct Smem {
    static constexpr int TM = 16 * P;
    // weight-slab pipeline depth: 2 for the dense tiles (two CTAs per SM must fit), 6 for the small
    // gather tiles whose 256 FFMA per slab cannot hide an L2 round trip behind a single prefetch
    static constexpr int NS = P == 4 ? 2 : (P == 2 ? 4 : 6);
    float A[KA][TM];
    float W[NS][KC][FN_H];
    float pos[3][TM];
    long long flat[TM];   // b * ppb + p of each tile slot, -1 if the slot is empty
    int bidx[TM];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void load_slab(float (*dst)[FN_H], const float* src, int tid) {
    // 16 rows x 256 floats = 1024 float4, 4 per thread, fully coalesced
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int q = tid + i * NTHREADS;
        cp_async16(&dst[0][0] + q * 4, src + q * 4);
    }
}

// acc[m][j*4+i] += sum_k A[k][tm*P+m] * Wt[k][tn*4 + j*64 + i]   for k in [0, K)
template <int P>
__device__ __forceinline__ void gemm_tile(Smem<P>& s, const float* __restrict__ wt, int K, float (&acc)[P][16], int tid) {
    const int tn = P == 4 ? (tid & 15) : (tid >> 4), tm = P == 4 ? (tid >> 4) : (tid & 15);
    const int nslab = K / KC;
    constexpr int NS = Smem<P>::NS;
#pragma unroll
    for (int st = 0; st < NS - 1; ++st) {
        if (st < nslab) load_slab(s.W[st], wt + (size_t)st * KC * FN_H, tid);
        cp_async_commit();
    }
    for (int c = 0; c < nslab; ++c) {
        cp_async_wait<NS - 2>();       // slab c has landed (one group per slab, NS-1 in flight)
        __syncthreads();               // ... for every thread, and slab c-1's buffer is free again
        if (c + NS - 1 < nslab) load_slab(s.W[(c + NS - 1) % NS], wt + (size_t)(c + NS - 1) * KC * FN_H, tid);
        cp_async_commit();
        const float(*W)[FN_H] = s.W[c % NS];
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float av[P];
            if constexpr (P == 4) {
                const float4 a4 = *reinterpret_cast<const float4*>(&s.A[c * KC + kk][tm * 4]);
                av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
            } else {
#pragma unroll
                for (int m = 0; m < P; ++m) av[m] = s.A[c * KC + kk][tm * P + m];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(&W[kk][tn * 4 + j * 64]);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int m = 0; m < P; ++m)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][j * 4 + i] = fmaf(av[m], wv[i], acc[m][j * 4 + i]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();                   // all reads of A and W done before the caller overwrites A
}

// A[col][pt] = sin(freq * acc + phase), per-point FiLM rows (points of a tile may belong to
// different batch elements in gather mode)
template <int P>
__device__ __forceinline__ void film_store(Smem<P>& s, const float* __restrict__ film, int n_film, int layer,
                                           float (&acc)[P][16], int tid) {
    const int tn = P == 4 ? (tid & 15) : (tid >> 4), tm = P == 4 ? (tid >> 4) : (tid & 15);
    const float* fl[P];
#pragma unroll
    for (int m = 0; m < P; ++m) fl[m] = film + ((size_t)s.bidx[tm * P + m] * n_film + layer) * 2 * FN_H;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int col = tn * 4 + j * 64 + i;
#pragma unroll
            for (int m = 0; m < P; ++m) {
                float fr = __ldg(fl[m] + col), ph = __ldg(fl[m] + FN_H + col);
                // torch: sin(freq * x + phase_shift), mul and add rounded separately (siren.py:123)
                s.A[col][tm * P + m] = sinf(__fadd_rn(__fmul_rn(fr, acc[m][j * 4 + i]), ph));
            }
        }
}

template <int P>
__device__ __forceinline__ void init_bias(const float* __restrict__ bias, float (&acc)[P][16], int tid) {
    const int tn = P == 4 ? (tid & 15) : (tid >> 4);
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float b = __ldg(bias + tn * 4 + j * 64 + i);
#pragma unroll
            for (int m = 0; m < P; ++m) acc[m][j * 4 + i] = b;
        }
}

// trilinear lookup, align_corners=True, zero padding; x -> W (innermost), y -> H, z -> D
// (siren/siren.py:314-330; corner order and weight products as ATen's grid_sampler_3d)
__device__ __forceinline__ float grid_feature(const float* __restrict__ grid, int R, int G, float x, float y, float z, int ch) {
    const float half = (float)(R - 1);
    float ix = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.f), 2.f), half);
    float iy = __fmul_rn(__fdiv_rn(__fadd_rn(y, 1.f), 2.f), half);
    float iz = __fmul_rn(__fdiv_rn(__fadd_rn(z, 1.f), 2.f), half);
    float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    float x1f = x0f + 1.f, y1f = y0f + 1.f, z1f = z0f + 1.f;
    float wx0 = __fsub_rn(x1f, ix), wx1 = __fsub_rn(ix, x0f);
    float wy0 = __fsub_rn(y1f, iy), wy1 = __fsub_rn(iy, y0f);
    float wz0 = __fsub_rn(z1f, iz), wz1 = __fsub_rn(iz, z0f);
    // guard the float->int conversion against wild coordinates
    auto clampi = [](float f) { return (int)fminf(fmaxf(f, -2.f), 1.0e6f); };
    int x0 = clampi(x0f), y0 = clampi(y0f), z0 = clampi(z0f);
    int x1 = x0 + 1, y1 = y0 + 1, z1 = z0 + 1;
    float out = 0.f;
    auto tap = [&](int zz, int yy, int xx, float w) {
        if ((unsigned)xx < (unsigned)R && (unsigned)yy < (unsigned)R && (unsigned)zz < (unsigned)R)
            out = __fadd_rn(out, __fmul_rn(__ldg(grid + (((size_t)zz * R + yy) * R + xx) * G + ch), w));
    };
    tap(z0, y0, x0, __fmul_rn(__fmul_rn(wx0, wy0), wz0));
    tap(z0, y0, x1, __fmul_rn(__fmul_rn(wx1, wy0), wz0));
    tap(z0, y1, x0, __fmul_rn(__fmul_rn(wx0, wy1), wz0));
    tap(z0, y1, x1, __fmul_rn(__fmul_rn(wx1, wy1), wz0));
    tap(z1, y0, x0, __fmul_rn(__fmul_rn(wx0, wy0), wz1));
    tap(z1, y0, x1, __fmul_rn(__fmul_rn(wx1, wy0), wz1));
    tap(z1, y1, x0, __fmul_rn(__fmul_rn(wx0, wy1), wz1));
    tap(z1, y1, x1, __fmul_rn(__fmul_rn(wx1, wy1), wz1));
    return out;
}

template <bool kGather, int P>
__device__ __forceinline__ void siren_exact_body(const ExactArgs& a, unsigned char* smem_raw) {
    constexpr int TM = 16 * P;
    Smem<P>& s = *reinterpret_cast<Smem<P>*>(smem_raw);
    const int tid = threadIdx.x;
    const FnLayout& L = a.L;
    const unsigned char* pk = a.packed;
    const int C = L.out_dim;

    long long n_items = a.n_items;
    int n_only = a.n_only;
    if (kGather && a.n_only_dev) {
        n_only = *a.n_only_dev;
        n_items = ((long long)n_only + TM - 1) / TM;
    }
    for (long long tile = blockIdx.x; tile < n_items; tile += gridDim.x) {
        // ---- tile bookkeeping, inputs ----
        if (tid < TM) {
            long long flat = -1;
            if (kGather) {
                long long q = tile * TM + tid;
                if (q < n_only) flat = a.only_idx[q];
            } else {
                long long b = tile / a.tiles_per_batch;
                long long p = (tile % a.tiles_per_batch) * TM + tid;
                if (p < a.ppb) flat = b * a.ppb + p;
            }
            s.flat[tid] = flat;
            long long fsafe = flat < 0 ? 0 : flat;
            int b = (int)(fsafe / a.ppb);
            s.bidx[tid] = b;
            float px = 0.f, py = 0.f, pz = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (flat >= 0) {
                px = __fmul_rn(a.points[flat * 3 + 0], L.input_scale);
                py = __fmul_rn(a.points[flat * 3 + 1], L.input_scale);
                pz = __fmul_rn(a.points[flat * 3 + 2], L.input_scale);
                if (a.lock_dirs) { d2 = -1.f; }
                else {
                    long long p = fsafe % a.ppb;
                    long long di = (long long)b * (a.ppb / a.dir_group) + p / a.dir_group;
                    d0 = a.dirs[di * 3 + 0]; d1 = a.dirs[di * 3 + 1]; d2 = a.dirs[di * 3 + 2];
                }
            }
            s.pos[0][tid] = px; s.pos[1][tid] = py; s.pos[2][tid] = pz;
            s.A[FN_H + 0][tid] = d0; s.A[FN_H + 1][tid] = d1; s.A[FN_H + 2][tid] = d2;
        }
        for (int i = tid; i < (KA - FN_H - 3) * TM; i += NTHREADS) s.A[FN_H + 3 + i / TM][i % TM] = 0.f;
        __syncthreads();
        if (L.grid_channels > 0 && !a.sigma_only) {
            const float* grid = reinterpret_cast<const float*>(pk + L.grid);
            const int G = L.grid_channels;
            for (int it = tid; it < TM * G; it += NTHREADS) {
                int pt = it / G, ch = it % G;
                s.A[FN_H + 3 + ch][pt] = grid_feature(grid, L.grid_res, G, s.pos[0][pt], s.pos[1][pt], s.pos[2][pt], ch);
            }
        }
        float acc[P][16];
        const int tn = P == 4 ? (tid & 15) : (tid >> 4), tm = P == 4 ? (tid >> 4) : (tid & 15);
        // ---- first layer: 3 -> 256 ----
        {
            const float* wt = reinterpret_cast<const float*>(pk + L.first_w);
            init_bias(reinterpret_cast<const float*>(pk + L.first_b), acc, tid);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float av[P];
#pragma unroll
                for (int m = 0; m < P; ++m) av[m] = s.pos[k][tm * P + m];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float w = __ldg(wt + k * FN_H + tn * 4 + j * 64 + i);
#pragma unroll
                        for (int m = 0; m < P; ++m) acc[m][j * 4 + i] = fmaf(av[m], w, acc[m][j * 4 + i]);
                    }
            }
            __syncthreads();   // grid features / pos reads done before A rows < 256 are written
            film_store(s, a.film, L.n_film, 0, acc, tid);
            __syncthreads();
        }
        // ---- hidden layers ----
        for (int l = 0; l < L.n_hidden; ++l) {
            if (l == L.trunk_hidden) {
                // heads on the trunk output: sigma, then the pre-multiplied label map
                const float* sw = reinterpret_cast<const float*>(pk + L.sigma_w);
                const float* lw = reinterpret_cast<const float*>(pk + L.label_w);
                for (int it = tid; it < TM * (1 + (a.sigma_only ? 0 : L.label_dim)); it += NTHREADS) {
                    int pt = it % TM, o = it / TM;
                    const float* w = o == 0 ? sw : lw + (size_t)(o - 1) * FN_H;
                    float r = o == 0 ? __ldg(sw + FN_H) : __ldg(lw + FENERF_MAX_LABEL * FN_H + (o - 1));
                    for (int k = 0; k < FN_H; ++k) r = fmaf(s.A[k][pt], __ldg(w + k), r);
                    long long flat = s.flat[pt];
                    if (flat >= 0) {
                        if (kGather && a.guard_stats && o == 0) {
                            // how far the tcgen05 density was from this fp32 one, and whether its sign was wrong: the
                            // margin of the GUARD threshold, measured on the very weights / points being rendered
                            const float old = a.out[flat * C + C - 1];
                            atomicMax(a.guard_stats + 1, __float_as_int(fabsf(r - old)));
                            if ((old > 0.f) != (r > 0.f)) atomicAdd(a.guard_stats + 2, 1);
                        }
                        a.out[flat * C + (o == 0 ? C - 1 : o - 1)] = r;
                    }
                }
                if (kGather && a.guard_stats && tile == 0 && tid == 0) a.guard_stats[0] = n_only;
                __syncthreads();
                if (a.sigma_only) break;          // the colour branch does not feed the density
            }
            const int K = FN_H + (l == L.trunk_hidden ? L.kx_pad : 0);
            init_bias(reinterpret_cast<const float*>(pk + L.hid_b[l]), acc, tid);
            gemm_tile(s, reinterpret_cast<const float*>(pk + L.hid_w32[l]), K, acc, tid);
            film_store(s, a.film, L.n_film, l + 1, acc, tid);
            __syncthreads();
        }
        // ---- rgb head: sigmoid(Linear(256 -> 3)) ----
        if (!a.sigma_only) {
            const float* rw = reinterpret_cast<const float*>(pk + L.rgb_w);
            for (int it = tid; it < TM * 3; it += NTHREADS) {
                int pt = it % TM, o = it / TM;
                float r = __ldg(rw + 3 * FN_H + o);
                for (int k = 0; k < FN_H; ++k) r = fmaf(s.A[k][pt], __ldg(rw + o * FN_H + k), r);
                r = __fdiv_rn(1.f, __fadd_rn(1.f, expf(-r)));
                long long flat = s.flat[pt];
                if (flat >= 0) a.out[flat * C + L.label_dim + o] = r;
            }
        }
        __syncthreads();
    }
}

template <bool kGather, int P>
__global__ void __launch_bounds__(NTHREADS, 2) siren_exact_kernel(ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    siren_exact_body<kGather, P>(a, smem_raw);
}

// GUARD refinement: the list length is only known on the device.  Short lists (small renders; fields whose densities
// rarely come near zero) take 16-point tiles -- lowest latency per tile, one wave as long as the list fits
// 16 x gridDim points -- longer ones 32-point tiles, which keep it to one wave up to twice that and do twice the FFMA
// per weight slab fetched from L2.
__global__ void __launch_bounds__(NTHREADS, 1) siren_exact_guard_kernel(ExactArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n_only = *a.n_only_dev;
    if (n_only <= 16 * (int)gridDim.x) siren_exact_body<true, 1>(a, smem_raw);
    else siren_exact_body<true, 2>(a, smem_raw);
}

}  // namespace

int siren_points_exact(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                       const float* film, int batch, long long ppb, int dir_group, int lock_dirs,
                       const int32_t* only_idx, int n_only, float* out, cudaStream_t st, int sigma_only) {
    static_assert(sizeof(Smem<4>) <= 113 * 1024, "two CTAs per SM must fit");
    constexpr int TM = 64;
    ExactArgs a;
    a.L = L; a.packed = packed; a.points = points; a.dirs = dirs; a.film = film; a.only_idx = only_idx; a.out = out;
    a.n_only_dev = nullptr; a.guard_stats = nullptr;
    a.ppb = ppb; a.n_only = n_only; a.dir_group = dir_group < 1 ? 1 : dir_group; a.lock_dirs = lock_dirs;
    a.sigma_only = sigma_only ? 1 : 0;
    a.tiles_per_batch = (ppb + TM - 1) / TM;
    const bool gather = only_idx != nullptr;
    a.n_items = gather ? ((long long)n_only + TM - 1) / TM : (long long)batch * a.tiles_per_batch;
    if (a.n_items <= 0) return 0;
    FN_REQUIRE(L.kx_pad <= KA - FN_H, "extra colour inputs (%d) exceed the tile", L.kx_pad);
    FN_REQUIRE(ppb % a.dir_group == 0, "points_per_batch %lld not a multiple of dir_group %d", ppb, a.dir_group);
    size_t smem = sizeof(Smem<4>);
    int blocks = (int)(a.n_items < (long long)num_sms() * 2 ? a.n_items : (long long)num_sms() * 2);
    if (gather) {
        FN_CUDA_OK(cudaFuncSetAttribute(siren_exact_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        siren_exact_kernel<true, 4><<<blocks, NTHREADS, smem, st>>>(a);
    } else {
        FN_CUDA_OK(cudaFuncSetAttribute(siren_exact_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        siren_exact_kernel<false, 4><<<blocks, NTHREADS, smem, st>>>(a);
    }
    FN_LAUNCH_OK("siren_exact_kernel");
    return 0;
}

// ---- GUARD refinement -------------------------------------------------------------------------
// The reference's last compositing interval has delta = 1e10 (volumetric_rendering.py:24), so the
// far sample's alpha is a step function of sign(sigma): a 3e-4 fp16 error in sigma flips a pixel by
// O(1) when |sigma| is that small.  After the fast pass over the coarse samples, rays whose far
// sigma lies within tau of zero get that one sample re-evaluated by the exact kernel (the far
// coarse sample is always the last one after the merge: every fine depth lies below the last
// coarse mid-point).
namespace {
__global__ void guard_scan_kernel(const float* __restrict__ raw, long long n_rays, int S, int C, float tau,
                                  const float* __restrict__ noise_far, long long noise_stride, float noise_std,
                                  int32_t* __restrict__ count, int32_t* __restrict__ list, int32_t* __restrict__ stats,
                                  long long probe_stride) {
    if (stats && blockIdx.x == 0 && threadIdx.x == 0) stats[3] = __float_as_int(tau);
    for (long long ray = (long long)blockIdx.x * blockDim.x + threadIdx.x; ray < n_rays;
         ray += (long long)gridDim.x * blockDim.x) {
        long long pt = ray * S + (S - 1);
        float sig = raw[pt * C + (C - 1)];
        // the step of the final compositing sits at sigma + noise * noise_std = 0 (volumetric_rendering.py:27-32);
        // the far sample is the last one after the merge, so its noise is draw #6 at [ray, n_samples - 1]
        float pre = noise_far ? __fadd_rn(sig, __fmul_rn(noise_far[ray * noise_stride], noise_std)) : sig;
        // ~128 rays per launch are PROBES: re-evaluated whatever their density, so that the self-check statistics
        // (fenerf_guard_stats) see the fp16 error even when it is larger than tau (then few samples fall below tau and
        // the flagged ones alone would say nothing)
        if (fabsf(pre) < tau || !isfinite(sig) || (ray % probe_stride) == 0) {
            int slot = atomicAdd(count, 1);
            list[slot] = (int32_t)pt;
        }
    }
}
}  // namespace

int guard_refine(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                 const float* film, int batch, long long rays_per_batch, int num_steps, int lock_dirs, float tau,
                 const float* noise_far, long long noise_stride, float noise_std,
                 float* raw, int32_t* scratch_idx, int32_t* stats, cudaStream_t st) {
    long long n_rays = rays_per_batch * batch;
    FN_REQUIRE(n_rays * num_steps < 2147483647LL, "too many points for the 32-bit guard list");
    FN_CUDA_OK(cudaMemsetAsync(scratch_idx, 0, sizeof(int32_t), st));
    int threads = 256;
    long long want = (n_rays + threads - 1) / threads;
    int blocks = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
    if (stats) FN_CUDA_OK(cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), st));
    guard_scan_kernel<<<blocks, threads, 0, st>>>(raw, n_rays, num_steps, L.out_dim, tau, noise_far, noise_stride, noise_std,
                                                  scratch_idx, scratch_idx + 1, stats, n_rays / 128 > 0 ? n_rays / 128 : 1);
    FN_LAUNCH_OK("guard_scan_kernel");
    ExactArgs a;
    a.L = L; a.packed = packed; a.points = points; a.dirs = dirs; a.film = film; a.out = raw;
    a.only_idx = scratch_idx + 1; a.n_only_dev = scratch_idx; a.n_only = 0; a.n_items = 0;
    a.ppb = rays_per_batch * num_steps; a.tiles_per_batch = 1; a.dir_group = num_steps; a.lock_dirs = lock_dirs;
    a.sigma_only = 1;      // only the sign of the far sample's density matters; its colour stays the tcgen05 one
    a.guard_stats = stats;
    // (stats[0..2] were zeroed and stats[3] = tau written by guard_scan_kernel's launch above: no host memory is
    // touched here, so the whole refinement can sit inside a captured CUDA graph)
    // one CTA per SM (118 / 140 KB of shared memory); tile size chosen on the device from the list length
    size_t smem = sizeof(Smem<2>) > sizeof(Smem<1>) ? sizeof(Smem<2>) : sizeof(Smem<1>);
    FN_CUDA_OK(cudaFuncSetAttribute(siren_exact_guard_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    siren_exact_guard_kernel<<<num_sms(), NTHREADS, smem, st>>>(a);
    FN_LAUNCH_OK("siren_exact_kernel(guard)");
    return 0;
}

}  // namespace fn
