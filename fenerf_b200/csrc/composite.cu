// fenerf_composite: merge coarse + fine samples by depth, alpha-composite every channel, apply the
// background / fill options and write the image in NCHW, already mapped to [-1, 1].
//
// Replaces, per ray, cat + torch.sort + 2x gather (generators/generators.py:85-89), the final
// fancy_integration (generators/volumetric_rendering.py:18-106) and the softmax / reshape /
// permute / *2-1 epilogue (generators.py:97-104).  The reference materialises the gathered
// (B,N,2S,C) tensor (277 MB per 4 faces for the 22-channel field) and ~20 more elementwise passes;
// here one warp owns a ray: a stable rank sort of the 2S depths in shared memory, the same
// left-to-right transmittance product as torch.cumprod, lanes = channels for the weighted sums.
// HBM-bound: algorithmic bytes per ray = S' * (4 C + 4 [+4 noise]) in, 4 (C_img + 2) out.
#include "common.cuh"

namespace fn {

namespace {

constexpr int kMaxSamples = 128;   // 2 * 64
constexpr int kRaysPerBlock = 8;   // one warp per ray

__device__ __forceinline__ float softplus_torch(float x) { return x > 20.f ? x : log1pf(expf(x)); }

struct CompositeArgs {
    long long n_rays, rays_per_batch;
    int S, n_samples, C, C_img;
    int clamp_mode, last_back, white_back, black_back, fill_mode, softmax_label;
    float noise_std, fill_color;
    const float *raw_c, *z_c, *raw_f, *z_f, *noise;
    float *pixels, *depth, *wsum, *weights;
    int32_t* sort_idx;
};

__global__ void __launch_bounds__(kRaysPerBlock * 32) composite_kernel(CompositeArgs A) {
    __shared__ float s_z[kRaysPerBlock][kMaxSamples];      // unsorted, then sorted depths
    __shared__ float s_zs[kRaysPerBlock][kMaxSamples];
    __shared__ int s_ord[kRaysPerBlock][kMaxSamples];      // sorted position -> original sample
    __shared__ float s_a[kRaysPerBlock][kMaxSamples];
    __shared__ float s_t[kRaysPerBlock][kMaxSamples];
    __shared__ float s_w[kRaysPerBlock][kMaxSamples];
    __shared__ float s_out[FENERF_MAX_LABEL + 8][kRaysPerBlock + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = A.n_samples, S = A.S, C = A.C;
    const bool hier = (n != S);
    const long long n_groups = (A.n_rays + kRaysPerBlock - 1) / kRaysPerBlock;

    for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const long long ray = grp * kRaysPerBlock + warp;
        const bool valid = ray < A.n_rays;
        if (valid) {
            float* z = s_z[warp];
            float* zs = s_zs[warp];
            int* ord = s_ord[warp];
            float* a = s_a[warp];
            float* t = s_t[warp];
            float* w = s_w[warp];
            const long long base = ray * S;
            // concatenation order of the reference: [fine, coarse]
            for (int i = lane; i < n; i += 32) z[i] = hier ? (i < S ? A.z_f[base + i] : A.z_c[base + i - S]) : A.z_c[base + i];
            __syncwarp();
            // stable rank sort (ties keep concatenation order; torch.sort is unstable there, ties
            // have measure zero)
            for (int i = lane; i < n; i += 32) {
                float zi = z[i];
                int r = 0;
                for (int j = 0; j < n; ++j) {
                    float zj = z[j];
                    r += (zj < zi) || (zj == zi && j < i);
                }
                zs[r] = zi;
                ord[r] = i;
            }
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                int o = ord[j];
                const float* src = hier ? (o < S ? A.raw_f + (base + o) * C : A.raw_c + (base + o - S) * C)
                                        : A.raw_c + (base + o) * C;
                float sig = src[C - 1];
                if (A.noise) sig = __fadd_rn(sig, __fmul_rn(A.noise[ray * n + j], A.noise_std));
                float delta = (j < n - 1) ? __fsub_rn(zs[j + 1], zs[j]) : 1e10f;
                float act = A.clamp_mode == FENERF_CLAMP_RELU ? fmaxf(sig, 0.f) : softplus_torch(sig);
                float alpha = __fsub_rn(1.f, expf(__fmul_rn(-delta, act)));
                a[j] = alpha;
                t[j] = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            }
            __syncwarp();
            for (int j = lane; j < n; j += 32) {
                float T = 1.f;
                for (int q = 0; q < j; ++q) T = __fmul_rn(T, t[q]);
                w[j] = __fmul_rn(a[j], T);
            }
            __syncwarp();
            float wsum = 0.f;
            for (int j = 0; j < n; ++j) wsum = __fadd_rn(wsum, w[j]);
            if (A.last_back) {
                __syncwarp();
                if (lane == 0) w[n - 1] = __fadd_rn(w[n - 1], __fsub_rn(1.f, wsum));
                __syncwarp();
            }
            // lanes = channels: c < C-1 colour/label channels, lane C-1 accumulates depth
            float acc = 0.f;
            if (lane < C) {
                for (int j = 0; j < n; ++j) {
                    int o = ord[j];
                    const float* src = hier ? (o < S ? A.raw_f + (base + o) * C : A.raw_c + (base + o - S) * C)
                                            : A.raw_c + (base + o) * C;
                    float v = lane < C - 1 ? src[lane] : zs[j];
                    acc = __fadd_rn(acc, __fmul_rn(w[j], v));
                }
            }
            if (lane == C - 1 && A.depth) A.depth[ray] = acc;
            if (lane == 0 && A.wsum) A.wsum[ray] = wsum;
            if (A.weights) for (int j = lane; j < n; j += 32) A.weights[ray * n + j] = w[j];
            if (A.sort_idx) for (int j = lane; j < n; j += 32) A.sort_idx[ray * n + j] = ord[j];

            float v = acc;   // meaningful on lanes < C-1
            if (A.white_back) v = __fsub_rn(__fadd_rn(v, 1.f), wsum);
            if (A.black_back) v = __fadd_rn(v, __fmul_rn(__fsub_rn(1.f, wsum), -1.f));
            // fill modes (volumetric_rendering.py:53-102); out channel index oc for this lane
            int oc = lane;
            const bool pad = (A.fill_mode == FENERF_FILL_SEG_PADDING_BACKGROUND ||
                              A.fill_mode == FENERF_FILL_EVAL_SEG_PADDING_BACKGROUND);
            const bool empty = wsum < 0.9f;
            if (pad) {
                // background channel 0 is zero, colour/label channels shift up by one
                oc = lane + 1;
                if (empty && A.fill_color >= 0.f) v = A.fill_color;
            } else if (A.fill_mode == FENERF_FILL_DEBUG || A.fill_mode == FENERF_FILL_WEIGHT_DEBUG) {
                if (empty) v = (lane == 0) ? 1.f : 0.f;
            } else if (A.fill_mode == FENERF_FILL_EVAL_WHITE_BACK) {
                if (empty) v = 1.f;
            }
            if (lane < C - 1) s_out[oc][warp] = v;
            if (pad && lane == 0) s_out[0][warp] = (empty && A.fill_color >= 0.f) ? 1.f : 0.f;
            __syncwarp();
            if (A.softmax_label) {
                // softmax over the channels before the last three (generators.py:97-100)
                const int n_seg = A.C_img - 3;
                float x = lane < n_seg ? s_out[lane][warp] : -INFINITY;
                float m = x;
                for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                float e = lane < n_seg ? expf(__fsub_rn(x, m)) : 0.f;
                float sum = e;
                for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (lane < n_seg) s_out[lane][warp] = __fdiv_rn(e, sum);
            }
        }
        __syncthreads();
        // coalesced NCHW store: 8 consecutive rays per channel row, *2-1
        for (int i = threadIdx.x; i < A.C_img * kRaysPerBlock; i += blockDim.x) {
            int c = i / kRaysPerBlock, r = i % kRaysPerBlock;
            long long rr = grp * kRaysPerBlock + r;
            if (rr < A.n_rays) {
                long long b = rr / A.rays_per_batch, p = rr % A.rays_per_batch;
                A.pixels[(b * A.C_img + c) * A.rays_per_batch + p] = __fsub_rn(__fmul_rn(s_out[c][r], 2.f), 1.f);
            }
        }
        __syncthreads();
    }
}

}  // namespace

int composite(const fenerf_render_desc* rd, int C, const float* raw_c, const float* z_c, const float* raw_f,
              const float* z_f, const float* noise, float* pixels, float* depth, float* wsum, float* weights,
              int32_t* sort_idx, cudaStream_t st) {
    CompositeArgs A;
    A.rays_per_batch = (long long)rd->img_h * rd->img_w;
    A.n_rays = A.rays_per_batch * rd->batch;
    A.S = rd->num_steps;
    A.n_samples = rd->hierarchical ? 2 * rd->num_steps : rd->num_steps;
    FN_REQUIRE(A.n_samples <= kMaxSamples && A.S >= 2, "num_steps %d unsupported (max %d per pass)", rd->num_steps,
               kMaxSamples / 2);
    FN_REQUIRE(C >= 2 && C <= 32, "out_dim %d unsupported", C);
    A.C = C;
    bool pad = rd->fill_mode == FENERF_FILL_SEG_PADDING_BACKGROUND ||
               rd->fill_mode == FENERF_FILL_EVAL_SEG_PADDING_BACKGROUND;
    A.C_img = C - 1 + (pad ? 1 : 0);
    A.clamp_mode = rd->clamp_mode;
    A.last_back = rd->last_back; A.white_back = rd->white_back; A.black_back = rd->black_back;
    A.fill_mode = rd->fill_mode; A.softmax_label = rd->softmax_label;
    A.noise_std = rd->noise_std; A.fill_color = rd->fill_color;
    A.raw_c = raw_c; A.z_c = z_c; A.raw_f = raw_f; A.z_f = z_f; A.noise = noise;
    A.pixels = pixels; A.depth = depth; A.wsum = wsum; A.weights = weights; A.sort_idx = sort_idx;
    if (rd->hierarchical) FN_REQUIRE(raw_f && z_f, "hierarchical render needs raw_fine and z_fine");
    long long groups = (A.n_rays + kRaysPerBlock - 1) / kRaysPerBlock;
    int blocks = (int)(groups < (long long)num_sms() * 8 ? groups : (long long)num_sms() * 8);
    if (blocks < 1) blocks = 1;
    composite_kernel<<<blocks, kRaysPerBlock * 32, 0, st>>>(A);
    FN_LAUNCH_OK("composite_kernel");
    return 0;
}

}  // namespace fn
