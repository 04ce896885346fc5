// fenerf_ray_setup: camera rays, stratified perturbation and the camera-to-world transform.
//
// Replaces get_initial_rays_trig + perturb_points + the three bmm of transform_sampled_points
// (generators/volumetric_rendering.py:109-168).  The reference materialises B identical copies of
// the camera-space points, a homogeneous pad and three batched matmuls (five passes over a
// (B,N,S,3..4) tensor); here one pass writes the world-space sample points and depths directly.
// HBM-bound: algorithmic bytes per (ray, sample) = 4 (rng) + 12 (point) + 4 (z) = 20 B.
//
// The arithmetic keeps the reference's operation order (unfused mul/add where torch rounds twice)
// so that positions agree to the last ulp or two; ray index p = row * W + col exactly
// (volumetric_rendering.py:115-118).
#include "common.cuh"

namespace fn {

namespace {

__global__ void __launch_bounds__(256)
ray_setup_kernel(int B, int H, int W, int S, float tan_half, const float* __restrict__ x_lin,
                 const float* __restrict__ y_lin, const float* __restrict__ z_lin, const float* __restrict__ c2w,
                 const float* __restrict__ rng, float* __restrict__ points, float* __restrict__ z_vals,
                 float* __restrict__ dirs, float* __restrict__ origins) {
    const long long N = (long long)H * W;
    const long long total = (long long)B * N * S;
    const float zc = __fdiv_rn(-1.0f, tan_half);
    const float dist = __fsub_rn(z_lin[1], z_lin[0]);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int s = (int)(idx % S);
        long long ray = idx / S;
        int b = (int)(ray / N);
        int p = (int)(ray % N);
        int row = p / W, col = p % W;
        float x = x_lin[col], y = y_lin[row];
        float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(zc, zc)));
        float d0 = __fdiv_rn(x, nrm), d1 = __fdiv_rn(y, nrm), d2 = __fdiv_rn(zc, nrm);
        float zl = z_lin[s];
        float off = __fmul_rn(__fsub_rn(rng[idx], 0.5f), dist);
        float zv = __fadd_rn(zl, off);
        float p0 = __fadd_rn(__fmul_rn(d0, zl), __fmul_rn(off, d0));
        float p1 = __fadd_rn(__fmul_rn(d1, zl), __fmul_rn(off, d1));
        float p2 = __fadd_rn(__fmul_rn(d2, zl), __fmul_rn(off, d2));
        const float* M = c2w + (size_t)b * 16;
        float out[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float acc = __fmul_rn(M[i * 4 + 0], p0);
            acc = fmaf(M[i * 4 + 1], p1, acc);
            acc = fmaf(M[i * 4 + 2], p2, acc);
            acc = __fadd_rn(acc, M[i * 4 + 3]);
            out[i] = acc;
        }
        points[idx * 3 + 0] = out[0];
        points[idx * 3 + 1] = out[1];
        points[idx * 3 + 2] = out[2];
        z_vals[idx] = zv;
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float acc = __fmul_rn(M[i * 4 + 0], d0);
                acc = fmaf(M[i * 4 + 1], d1, acc);
                acc = fmaf(M[i * 4 + 2], d2, acc);
                dirs[ray * 3 + i] = acc;
            }
            if (p == 0) {
                origins[b * 3 + 0] = M[3];
                origins[b * 3 + 1] = M[7];
                origins[b * 3 + 2] = M[11];
            }
        }
    }
}

}  // namespace

int ray_setup(const fenerf_render_desc* rd, const float* x_lin, const float* y_lin, const float* z_lin,
              const float* cam2world, const float* rng_perturb, float* points, float* z_vals, float* dirs,
              float* origins, cudaStream_t st) {
    long long total = (long long)rd->batch * rd->img_h * rd->img_w * rd->num_steps;
    int threads = 256;
    long long want = (total + threads - 1) / threads;
    int blocks = (int)(want < (long long)num_sms() * 16 ? want : (long long)num_sms() * 16);
    if (blocks < 1) blocks = 1;
    ray_setup_kernel<<<blocks, threads, 0, st>>>(rd->batch, rd->img_h, rd->img_w, rd->num_steps, rd->tan_half_fov, x_lin,
                                                 y_lin, z_lin, cam2world, rng_perturb, points, z_vals, dirs, origins);
    FN_LAUNCH_OK("ray_setup_kernel");
    return 0;
}

}  // namespace fn
