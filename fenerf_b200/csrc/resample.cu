// fenerf_resample: coarse compositing weights -> inverse-CDF resampling -> fine sample points.
//
// Replaces, per ray, fancy_integration(coarse)[2] (generators/volumetric_rendering.py:18-38), the
// resample prep (generators/generators.py:63-74) and sample_pdf (volumetric_rendering.py:259-300).
// The reference runs ~25 elementwise/scan/gather passes over (B*N, S) tensors plus a
// searchsorted; here one warp owns a ray, everything stays in shared memory / registers and the
// only HBM traffic is sigma + z + u in, z_fine + points_fine out.
// HBM-bound: algorithmic bytes per ray = S * (4 sigma + 4 z + 4 u [+4 noise]) in,
//                                         S * (4 z_fine + 12 point [+8 inds]) out.
//
// Rounding follows the reference op by op: the transmittance is the same left-to-right product
// as torch.cumprod, the CDF the same left-to-right sum as torch.cumsum, `inds` is
// searchsorted(cdf, u, right=False).  (torch.sum's vectorised order is host-ISA dependent and is
// not reproduced: a sequential sum is used for the pdf normaliser.)
#include "common.cuh"

namespace fn {

namespace {

constexpr int kMaxS = 64;
constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float softplus_torch(float x) {
    // F.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
resample_kernel(long long n_rays, long long rays_per_batch, int S, int C, int clamp_mode, float noise_std,
                const float* __restrict__ raw, const float* __restrict__ z_vals, const float* __restrict__ dirs,
                const float* __restrict__ origins, const float* __restrict__ noise, const float* __restrict__ u,
                float* __restrict__ z_fine, float* __restrict__ pts_fine, long long* __restrict__ inds, int sort_fine) {
    __shared__ float s_z[kWarpsPerBlock][kMaxS];
    __shared__ float s_t[kWarpsPerBlock][kMaxS];    // 1 - alpha + 1e-10
    __shared__ float s_a[kWarpsPerBlock][kMaxS];    // alpha
    __shared__ float s_w[kWarpsPerBlock][kMaxS];    // interior weights + 2e-5, then pdf
    __shared__ float s_cdf[kWarpsPerBlock][kMaxS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* z = s_z[warp];
    float* t = s_t[warp];
    float* a = s_a[warp];
    float* w = s_w[warp];
    float* cdf = s_cdf[warp];

    for (long long ray = (long long)blockIdx.x * kWarpsPerBlock + warp; ray < n_rays;
         ray += (long long)gridDim.x * kWarpsPerBlock) {
        const long long base = ray * S;
        for (int s = lane; s < S; s += 32) z[s] = z_vals[base + s];
        __syncwarp();
        // the far sample (s = S-1) never enters the interior weights, so it is not read
        for (int s = lane; s < S - 1; s += 32) {
            float sig = raw[(base + s) * C + (C - 1)];
            if (noise) sig = __fadd_rn(sig, __fmul_rn(noise[base + s], noise_std));
            float delta = (s < S - 1) ? __fsub_rn(z[s + 1], z[s]) : 1e10f;
            float act = clamp_mode == FENERF_CLAMP_RELU ? fmaxf(sig, 0.f) : softplus_torch(sig);
            float alpha = __fsub_rn(1.f, expf(__fmul_rn(-delta, act)));
            a[s] = alpha;
            t[s] = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
        }
        __syncwarp();
        // weights = alpha * cumprod([1, t...])[:-1]; keep only the interior ones, +1e-5 twice
        // (generators.py:63 and sample_pdf's eps, volumetric_rendering.py:273)
        for (int j = lane; j < S - 2; j += 32) {
            int s = j + 1;
            float T = 1.f;
            for (int q = 0; q < s; ++q) T = __fmul_rn(T, t[q]);
            float wt = __fmul_rn(a[s], T);
            w[j] = __fadd_rn(__fadd_rn(wt, 1e-5f), 1e-5f);
        }
        __syncwarp();
        float total = 0.f;
        for (int j = 0; j < S - 2; ++j) total = __fadd_rn(total, w[j]);   // every lane, same order
        __syncwarp();
        for (int j = lane; j < S - 2; j += 32) w[j] = __fdiv_rn(w[j], total);   // pdf, one division per bin
        __syncwarp();
        // cdf[0] = 0, cdf[j+1] = cdf[j] + pdf[j]   (S-1 entries)
        for (int i = lane; i < S - 1; i += 32) {
            float c = 0.f;
            for (int j = 0; j < i; ++j) c = __fadd_rn(c, w[j]);
            cdf[i] = c;
        }
        __syncwarp();
        const int b = (int)((unsigned)ray / (unsigned)rays_per_batch);     // n_rays < 2^31 (checked by the host)
        const float o0 = origins[b * 3 + 0], o1 = origins[b * 3 + 1], o2 = origins[b * 3 + 2];
        const float d0 = dirs[ray * 3 + 0], d1 = dirs[ray * 3 + 1], d2 = dirs[ray * 3 + 2];
        const int n_cdf = S - 1;
        for (int k = lane; k < S; k += 32) {
            float uu = u[base + k];
            // searchsorted, right=False: first i with cdf[i] >= uu, n_cdf if none
            int lo = 0, hi = n_cdf;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cdf[mid] < uu) lo = mid + 1; else hi = mid;
            }
            int below = lo - 1 < 0 ? 0 : lo - 1;
            int above = lo > S - 2 ? S - 2 : lo;
            float cb = cdf[below], ca = cdf[above];
            float bb = __fmul_rn(0.5f, __fadd_rn(z[below], z[below + 1]));
            float ba = __fmul_rn(0.5f, __fadd_rn(z[above], z[above + 1]));
            float denom = __fsub_rn(ca, cb);
            if (denom < 1e-5f) denom = 1.f;
            float zf = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uu, cb), denom), __fsub_rn(ba, bb)));
            if (inds) inds[base + k] = lo;
            if (sort_fine) { t[k] = zf; continue; }      // t[] is free by now: holds the unsorted fine depths
            z_fine[base + k] = zf;
            pts_fine[(base + k) * 3 + 0] = __fadd_rn(o0, __fmul_rn(d0, zf));
            pts_fine[(base + k) * 3 + 1] = __fadd_rn(o1, __fmul_rn(d1, zf));
            pts_fine[(base + k) * 3 + 2] = __fadd_rn(o2, __fmul_rn(d2, zf));
        }
        __syncwarp();
        if (sort_fine) {
            // the render path wants the fine samples in depth order (the merge with the coarse list is then a
            // two-pointer walk, composite.cu); the order of a ray's fine samples is otherwise immaterial: the
            // reference sorts them itself right after (generators.py:85-89).  Stable rank sort, ties keep draw order.
            for (int k = lane; k < S; k += 32) {
                const float zk = t[k];
                int r = 0;
                for (int j = 0; j < S; ++j) {
                    const float zj = t[j];
                    r += (zj < zk) || (zj == zk && j < k);
                }
                z_fine[base + r] = zk;
                pts_fine[(base + r) * 3 + 0] = __fadd_rn(o0, __fmul_rn(d0, zk));
                pts_fine[(base + r) * 3 + 1] = __fadd_rn(o1, __fmul_rn(d1, zk));
                pts_fine[(base + r) * 3 + 2] = __fadd_rn(o2, __fmul_rn(d2, zk));
            }
            __syncwarp();
        }
    }
}

}  // namespace

int resample(const fenerf_render_desc* rd, int C, const float* raw, const float* z, const float* dirs,
             const float* origins, const float* noise, const float* u, float* z_fine, float* pts_fine,
             long long* inds, cudaStream_t st, int sort_fine) {
    FN_REQUIRE(rd->num_steps >= 3 && rd->num_steps <= kMaxS, "num_steps %d outside [3, %d] for resampling",
               rd->num_steps, kMaxS);
    long long rpb = (long long)rd->img_h * rd->img_w;
    long long n_rays = rpb * rd->batch;
    FN_REQUIRE(n_rays < (1ll << 31), "too many rays for one launch: %lld", n_rays);
    long long want = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
    int blocks = (int)(want < (long long)num_sms() * 8 ? want : (long long)num_sms() * 8);
    if (blocks < 1) blocks = 1;
    resample_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(n_rays, rpb, rd->num_steps, C, rd->clamp_mode, rd->noise_std,
                                                            raw, z, dirs, origins, noise, u, z_fine, pts_fine, inds, sort_fine);
    FN_LAUNCH_OK("resample_kernel");
    return 0;
}

}  // namespace fn
