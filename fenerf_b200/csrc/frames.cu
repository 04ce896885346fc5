// Frame consumers: what the render CLIs and the FID dump do with every frame.
//
//   mask2color_kernel     argmax over the label channels + the 19-entry colour table
//                         (train_double_latent_semantic.py:36-55, 66-72: a Python loop of 19 masked assignments on the CPU)
//   frames_to_u8_kernel   [-1, 1] float NCHW -> uint8 NHWC, torchvision.utils.save_image(normalize=True, range=(-1, 1))
//                         semantics (fid_evaluation.py:146-151): ((x + 1) / 2 clamped to [0, 1]) * 255 + 0.5, truncated
// Both are pure streaming kernels (HBM-bound, one pass): at ~1100 faces/s the reference's CPU loops here would be the
// bottleneck of every render script.
#include "common.cuh"

namespace fn {

namespace {

__constant__ unsigned char kColorMap[19][3] = {
    {0, 0, 0}, {204, 0, 0}, {76, 153, 0}, {204, 204, 0}, {51, 51, 255}, {204, 0, 204}, {0, 255, 255}, {255, 204, 204},
    {102, 51, 0}, {255, 0, 0}, {102, 204, 0}, {255, 255, 0}, {0, 0, 153}, {0, 0, 204}, {255, 51, 153}, {0, 204, 204},
    {0, 51, 0}, {255, 153, 51}, {0, 204, 0}};

__global__ void __launch_bounds__(256) mask2color_kernel(const float* __restrict__ masks, int B, int K, long long HW,
                                                         float* __restrict__ out) {
    const long long total = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / HW, p = i % HW;
        const float* m = masks + b * K * HW + p;
        float best = m[0];
        int arg = 0;
        for (int k = 1; k < K; ++k) {
            const float v = m[(long long)k * HW];
            if (v > best) { best = v; arg = k; }          // first maximum wins, as torch.argmax
        }
        float r = 0.f, g = 0.f, bl = 0.f;
        if (arg < 19) { r = kColorMap[arg][0]; g = kColorMap[arg][1]; bl = kColorMap[arg][2]; }
        float* o = out + b * 3 * HW + p;
        o[0] = r; o[HW] = g; o[2 * HW] = bl;
    }
}

__global__ void __launch_bounds__(256) frames_to_u8_kernel(const float* __restrict__ frames, int B, int C, int c0, int nc, long long HW,
                                                           unsigned char* __restrict__ out) {
    const long long total = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / HW, p = i % HW;
        for (int c = 0; c < nc; ++c) {
            float v = frames[(b * C + c0 + c) * HW + p];
            v = fminf(fmaxf((v + 1.f) * 0.5f, 0.f), 1.f);
            out[(b * HW + p) * nc + c] = (unsigned char)fminf(fmaxf(v * 255.f + 0.5f, 0.f), 255.f);
        }
    }
}

int blocks_for(long long n) {
    long long want = (n + 255) / 256, cap = (long long)num_sms() * 16;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

int mask2color(const float* masks, int B, int K, long long HW, float* out, cudaStream_t st) {
    mask2color_kernel<<<blocks_for((long long)B * HW), 256, 0, st>>>(masks, B, K, HW, out);
    FN_LAUNCH_OK("mask2color_kernel");
    return 0;
}

int frames_to_u8(const float* frames, int B, int C, int c0, int nc, long long HW, unsigned char* out, cudaStream_t st) {
    frames_to_u8_kernel<<<blocks_for((long long)B * HW), 256, 0, st>>>(frames, B, C, c0, nc, HW, out);
    FN_LAUNCH_OK("frames_to_u8_kernel");
    return 0;
}

}  // namespace fn
