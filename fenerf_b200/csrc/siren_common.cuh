// Device helpers shared by the two point-network kernels.  Internal.
#pragma once
#include "common.cuh"

namespace fn {

// Trilinear lookup of all 32 channels of the channels-last feature grid at one position:
// align_corners=True, zero padding; x indexes the innermost grid axis (W), y -> H, z -> D
// (sample_from_3dgrid, siren/siren.py:314-330; corner order and weight products follow ATen's
// grid_sampler_3d so the fp32 result matches the reference to an ulp or two).
__device__ __forceinline__ void grid_features32(const float* __restrict__ grid, int R, float x, float y, float z,
                                                float (&out)[32]) {
    const float half = (float)(R - 1);
    float ix = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.f), 2.f), half);
    float iy = __fmul_rn(__fdiv_rn(__fadd_rn(y, 1.f), 2.f), half);
    float iz = __fmul_rn(__fdiv_rn(__fadd_rn(z, 1.f), 2.f), half);
    float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    float wx1 = __fsub_rn(ix, x0f), wx0 = __fsub_rn(x0f + 1.f, ix);
    float wy1 = __fsub_rn(iy, y0f), wy0 = __fsub_rn(y0f + 1.f, iy);
    float wz1 = __fsub_rn(iz, z0f), wz0 = __fsub_rn(z0f + 1.f, iz);
    auto clampi = [](float f) { return (int)fminf(fmaxf(f, -2.f), 1.0e6f); };
    const int x0 = clampi(x0f), y0 = clampi(y0f), z0 = clampi(z0f);
#pragma unroll
    for (int c = 0; c < 32; ++c) out[c] = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        const int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;   // tnw, tne, tsw, tse, bnw, ...
        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        if ((unsigned)xx < (unsigned)R && (unsigned)yy < (unsigned)R && (unsigned)zz < (unsigned)R) {
            const float w = __fmul_rn(__fmul_rn(dx ? wx1 : wx0, dy ? wy1 : wy0), dz ? wz1 : wz0);
            const float4* src = reinterpret_cast<const float4*>(grid + (((size_t)zz * R + yy) * R + xx) * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 g = __ldg(src + c4);
                out[c4 * 4 + 0] = __fadd_rn(out[c4 * 4 + 0], __fmul_rn(g.x, w));
                out[c4 * 4 + 1] = __fadd_rn(out[c4 * 4 + 1], __fmul_rn(g.y, w));
                out[c4 * 4 + 2] = __fadd_rn(out[c4 * 4 + 2], __fmul_rn(g.z, w));
                out[c4 * 4 + 3] = __fadd_rn(out[c4 * 4 + 3], __fmul_rn(g.w, w));
            }
        }
    }
}

// The same lookup from the fp16 channels-last copy of the grid (one voxel = 64 B): corner weights and the
// accumulation stay fp32, only the stored features are rounded (they are rounded to fp16 MMA operands right after).
__device__ __forceinline__ void grid_features32_h(const __half* __restrict__ grid, int R, float x, float y, float z,
                                                  float (&out)[32]) {
    const float half = (float)(R - 1);
    float ix = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.f), 2.f), half);
    float iy = __fmul_rn(__fdiv_rn(__fadd_rn(y, 1.f), 2.f), half);
    float iz = __fmul_rn(__fdiv_rn(__fadd_rn(z, 1.f), 2.f), half);
    float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    float wx1 = __fsub_rn(ix, x0f), wx0 = __fsub_rn(x0f + 1.f, ix);
    float wy1 = __fsub_rn(iy, y0f), wy0 = __fsub_rn(y0f + 1.f, iy);
    float wz1 = __fsub_rn(iz, z0f), wz0 = __fsub_rn(z0f + 1.f, iz);
    auto clampi = [](float f) { return (int)fminf(fmaxf(f, -2.f), 1.0e6f); };
    const int x0 = clampi(x0f), y0 = clampi(y0f), z0 = clampi(z0f);
#pragma unroll
    for (int c = 0; c < 32; ++c) out[c] = 0.f;
#pragma unroll
    for (int corner = 0; corner < 8; ++corner) {
        const int dx = corner & 1, dy = (corner >> 1) & 1, dz = corner >> 2;
        const int xx = x0 + dx, yy = y0 + dy, zz = z0 + dz;
        if ((unsigned)xx < (unsigned)R && (unsigned)yy < (unsigned)R && (unsigned)zz < (unsigned)R) {
            const float w = __fmul_rn(__fmul_rn(dx ? wx1 : wx0, dy ? wy1 : wy0), dz ? wz1 : wz0);
            const uint4* src = reinterpret_cast<const uint4*>(grid + (((size_t)zz * R + yy) * R + xx) * 32);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                const uint4 g = __ldg(src + c8);
                const __half2* h2 = reinterpret_cast<const __half2*>(&g);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f2 = __half22float2(h2[i]);
                    out[c8 * 8 + 2 * i] = fmaf(f2.x, w, out[c8 * 8 + 2 * i]);
                    out[c8 * 8 + 2 * i + 1] = fmaf(f2.y, w, out[c8 * 8 + 2 * i + 1]);
                }
            }
        }
    }
}

}  // namespace fn
