/*
 * fenerf_b200 -- C-ABI of the B200-native volumetric face renderer.
 *
 * This is the drop-in boundary for the one hot path of MrTornado24/FENeRF: the per-image
 * volumetric render inside generators.*Generator3d.forward / staged_forward.  The reference has
 * no FFI on this path (it is ~330 lines of torch ops behind a Python class API), so the entry
 * points below are what a binding for that path binds: each one replaces a span of the
 * reference's Python, cited per function (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes; every pointer is a DEVICE pointer unless marked host;
 *   - the caller owns every buffer; the library never allocates, holds no global state besides a
 *     thread-local error string, and is re-entrant per stream;
 *   - all launches go to the given cudaStream_t (passed as void*; NULL = legacy default stream);
 *   - return 0 on success, a negative FENERF_E_* code otherwise; fenerf_last_error() explains;
 *   - tensors are contiguous fp32 unless noted; B batch, N = img_h*img_w rays, S = num_steps,
 *     C = out_dim channels per point ordered [labels.., r, g, b, sigma].
 *
 * INTEGRATION.md shows the ctypes stub that binds this header from the reference side.
 */
#ifndef FENERF_B200_H
#define FENERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FENERF_ABI_VERSION 2   /* 2: FENERF_MAX_COLOR 4 -> 8 (fenerf_field_params grew), backward entry points */

/* error codes */
#define FENERF_OK            0
#define FENERF_E_ARG        -1   /* bad argument (null pointer, unsupported size, misaligned) */
#define FENERF_E_UNSUPPORTED -2  /* valid in the reference, not built here (message says what) */
#define FENERF_E_CUDA       -3   /* a CUDA runtime call failed */
#define FENERF_E_WORKSPACE  -4   /* workspace too small */
#define FENERF_E_CLAMP_MODE -5   /* reference raises "Need to choose clamp mode"
                                    (generators/volumetric_rendering.py:33-34) */

/* ---- point network ("field") --------------------------------------------------------------
 * Describes a FiLM-SIREN point network of the reference's siren/siren.py family:
 *   x = pos * input_scale                                 (UniformBoxWarp, siren.py:181-187)
 *   x = FiLM_0(3->256)(x); x = FiLM_i(256->256)(x), i < trunk_layers      (siren.py:113-123)
 *   sigma  = Linear(256->1)(x)
 *   labels = Linear chain 256->256->256->label_dim (no activation; pre-multiplied)   [optional]
 *   c = FiLM(cat[dir(3), grid_feat(G), x(256)] -> 256); c = FiLM(256->256)(c) ...  color_layers
 *   rgb = sigmoid(Linear(256->3)(c))
 *   out = [labels, rgb, sigma]
 * TALLSIREN (siren.py:126-178):  trunk 8, color 1, label 0, grid 0, scale 1, out 4.
 * TextureEmbeddingPiGAN256SEMANTICDISENTANGLE_DIM_96 (siren.py:1451-1546):
 *                                trunk 8, color 3, label 18, grid 32 x 96^3, scale 2/0.24, out 22.
 * The hidden width is fixed at 256.                                                            */
#define FENERF_MAX_TRUNK 8
#define FENERF_MAX_COLOR 8
#define FENERF_MAX_LABEL 32
#define FENERF_HIDDEN 256

typedef struct fenerf_field_desc {
    int32_t trunk_layers;   /* 2..8 */
    int32_t color_layers;   /* 1..8 */
    int32_t label_dim;      /* 0..32 */
    int32_t grid_channels;  /* 0 or 32 */
    int32_t grid_res;       /* cubic grid side (D = H = W), 0 if no grid */
    int32_t out_dim;        /* label_dim + 4 */
    float   input_scale;
    int32_t reserved;
} fenerf_field_desc;

/* Raw parameters exactly as torch stores them: nn.Linear.weight is [out][in] row-major fp32,
 * bias [out]; grid is the reference's channel-major (1, G, D, H, W) tensor (siren.py:1546). */
typedef struct fenerf_field_params {
    const float* trunk_w[FENERF_MAX_TRUNK];  /* [256][3] then [256][256] */
    const float* trunk_b[FENERF_MAX_TRUNK];
    const float* sigma_w;                    /* [1][256] */
    const float* sigma_b;                    /* [1] */
    const float* color_w[FENERF_MAX_COLOR];  /* first [256][3+G+256] (cat order dir, feat, x), rest [256][256] */
    const float* color_b[FENERF_MAX_COLOR];
    const float* rgb_w;                      /* [3][256] */
    const float* rgb_b;                      /* [3] */
    const float* label_w[3];                 /* [256][256], [256][256], [label_dim][256]; NULL if label_dim == 0;
                                                label_w[1] / label_b[1] NULL for a two-layer chain (siren.py:1189-1191) */
    const float* label_b[3];
    const float* grid;                       /* (1, G, R, R, R) or NULL */
} fenerf_field_params;

/* Bytes of the packed, kernel-layout copy of a field's parameters. */
size_t fenerf_packed_bytes(const fenerf_field_desc* field);

/* Re-lays the raw parameters out for the kernels (k-major fp32 for the exact path, UMMA
 * 128B-swizzled fp16 images for the tcgen05 path, pre-multiplied label chain, channels-last
 * grid).  Call again whenever the parameters change (optimizer step, EMA copy_to).
 * `packed` must be 1024-byte aligned. */
int fenerf_pack_field(const fenerf_field_desc* field, const fenerf_field_params* params,
                      void* packed, size_t packed_bytes, void* stream);

/* 128-bit fingerprint of the raw parameters (two position-weighted sums modulo 2^64 over their bit
 * patterns), written to `out[0..1]` on the device.  The host mirror compares it with the fingerprint taken
 * at pack time to decide whether the packed copy is stale: torch_ema's copy_to / restore
 * (render_multiview_images_double_semantic.py:63, train_double_latent_semantic.py:464-522) write through
 * `param.data.copy_`, which torch's version counters do not see. */
int fenerf_field_fingerprint(const fenerf_field_desc* field, const fenerf_field_params* params,
                             uint64_t* out /* device, 2 x u64 */, void* stream);

/* precision modes of the point network */
#define FENERF_PRECISION_EXACT 0  /* fp32 FFMA + precise sinf everywhere (CUDA cores)            */
#define FENERF_PRECISION_FAST  1  /* fp16 operands / fp32 accumulate on tcgen05, sin.approx      */
#define FENERF_PRECISION_GUARD 2  /* FAST, then EXACT re-evaluation of the far sample of every ray
                                     whose |sigma| < guard_tau (the reference's delta=1e10 step,
                                     volumetric_rendering.py:24,32) -- the default                */

/* Evaluates the field at arbitrary points.
 * Replaces <SIREN>.forward_with_frequencies_phase_shifts (siren/siren.py:164-178, 1509-1530);
 * also the entry extract_*shapes.py needs (extract_double_semantic_shapes.py:59,80).
 *   points  (B, P, 3)      positions, not yet box-warped
 *   dirs    (B, P/dir_group, 3)  one direction per `dir_group` consecutive points (1 = per point)
 *   film    (B, trunk+color, 2, 256)  [15*freq+30, phase] per FiLM layer
 *   out     (B, P, C)
 *   only_idx  optional int32 list (n_only entries) of flat point indices b*P+p to evaluate;
 *             others are left untouched in `out` (used by the GUARD refinement)               */
/* OR-ed into `precision`: only the density channel out[..., C-1] is required (a 256^3 density grid for
 * marching cubes, extract_double_semantic_shapes.py:59-62 keeps `[:, :, -1:]`); the colour / label
 * branches are skipped on the tcgen05 path and the other channels of `out` are then left unspecified. */
#define FENERF_POINTS_SIGMA_ONLY 0x100

int fenerf_siren_points(const fenerf_field_desc* field, const void* packed,
                        const float* points, const float* dirs, const float* film,
                        int32_t batch, int64_t points_per_batch, int32_t dir_group,
                        int32_t precision, const int32_t* only_idx, int32_t n_only,
                        float* out, void* stream);

/* ---- render ------------------------------------------------------------------------------ */
#define FENERF_CLAMP_RELU 0
#define FENERF_CLAMP_SOFTPLUS 1

/* fill_mode of fancy_integration (generators/volumetric_rendering.py:53-102) */
#define FENERF_FILL_NONE 0
#define FENERF_FILL_DEBUG 1
#define FENERF_FILL_WEIGHT 2
#define FENERF_FILL_WEIGHT_DEBUG 3
#define FENERF_FILL_SEG_PADDING_BACKGROUND 4
#define FENERF_FILL_EVAL_SEG_PADDING_BACKGROUND 5
#define FENERF_FILL_EVAL_WHITE_BACK 6

/* fill_color as the value written to the non-background channels: black 0, white 1, grey 0.5,
 * light_grey 0.81 (volumetric_rendering.py:74-81); negative = "no such colour" (pixels untouched) */

typedef struct fenerf_render_desc {
    int32_t batch;
    int32_t img_h, img_w;       /* reference always renders square; kept separate for clarity */
    int32_t num_steps;          /* S, coarse samples per ray (2..64) */
    int32_t hierarchical;       /* 1: resample S fine points per ray (generators.py:58-89) */
    int32_t clamp_mode;         /* FENERF_CLAMP_* ; anything else -> FENERF_E_CLAMP_MODE */
    int32_t last_back, white_back, black_back;
    int32_t fill_mode;          /* FENERF_FILL_* (staged_forward only) */
    float   fill_color;
    int32_t softmax_label;      /* softmax over the label channels of the composited pixel */
    int32_t lock_view_dependence; /* every direction := (0, 0, -1) (generators.py:50-52) */
    int32_t precision;          /* FENERF_PRECISION_* */
    float   noise_std;          /* nerf_noise */
    float   tan_half_fov;       /* (float) tan(2*pi*fov/360 / 2), volumetric_rendering.py:119 */
    float   guard_tau;          /* GUARD threshold on |sigma_far| (default 1.5e-3 if <= 0: 5x the fp16 path's
                                   measured 3e-4 max sigma error, profiles/r01_field_fast_vs_exact*.log) */
} fenerf_render_desc;

/* Camera pose sampling after the random draws, and the look-at camera-to-world matrix.
 * Replaces the arithmetic of sample_camera_positions + create_cam2world_matrix
 * (generators/volumetric_rendering.py:170-248); the caller makes the draws (theta first) so the RNG
 * stream is the reference's.
 *   mode FIXED: theta = h_mean, phi = v_mean (draws may be NULL)
 *   mode UNIFORM: (draw - 0.5) * 2 * stddev + mean            draws torch.rand (n,1)
 *   mode GAUSSIAN ('normal' / 'gaussian'): draw * stddev + mean   draws torch.randn (n,1)
 *   mode TRUNCATED_GAUSSIAN: draws torch.randn (n,1,4); the first of the four inside (-2, 2) (:170-177)
 *   mode SPHERICAL_UNIFORM: theta as UNIFORM; v = (draw - 0.5) * 2 * v_stddev + v_mean with v_stddev, v_mean
 *        ALREADY divided by pi by the caller (:214), clamped to [1e-5, 1 - 1e-5]; phi = arccos(1 - 2 v)
 *   'hybrid' (:198-204) is UNIFORM with doubled stddevs or GAUSSIAN, chosen by the caller's coin flip
 * out: cam2world (n,16) row-major; pitch (n) = clamped phi; yaw (n) = theta                        */
#define FENERF_CAMERA_FIXED 0
#define FENERF_CAMERA_UNIFORM 1
#define FENERF_CAMERA_GAUSSIAN 2
#define FENERF_CAMERA_TRUNCATED_GAUSSIAN 3
#define FENERF_CAMERA_SPHERICAL_UNIFORM 4
int fenerf_camera_poses(int32_t n, int32_t mode, float h_stddev, float v_stddev, float h_mean, float v_mean,
                        const float* draw_theta, const float* draw_phi, float* cam2world, float* pitch, float* yaw,
                        void* stream);

/* Camera rays, stratified perturbation and camera-to-world transform.
 * Replaces get_initial_rays_trig + perturb_points + the three bmm of transform_sampled_points
 * (generators/volumetric_rendering.py:109-168).
 *   x_lin (W) = linspace(-1,1,W); y_lin (H) = linspace(1,-1,H); z_lin (S) = linspace(near,far,S)
 *   cam2world (B,16) row-major 4x4 (create_cam2world_matrix, :230-248)
 *   rng_perturb (B,N,S) uniform [0,1) draw #1 (torch.rand, :135)
 * out: points (B,N,S,3) world space; z_vals (B,N,S); dirs (B,N,3) world; origins (B,3)        */
int fenerf_ray_setup(const fenerf_render_desc* rd, const float* x_lin, const float* y_lin,
                     const float* z_lin, const float* cam2world, const float* rng_perturb,
                     float* points, float* z_vals, float* dirs, float* origins, void* stream);

/* Coarse weights + inverse-CDF resampling + fine points.
 * Replaces fancy_integration(coarse) -> weights, the resample prep and sample_pdf
 * (generators/generators.py:59-74; volumetric_rendering.py:18-38, 259-300).
 *   raw_coarse (B,N,S,C) (sigma = last channel); z_vals (B,N,S)
 *   rng_noise (B,N,S) normal draw #4 or NULL (treated as 0; required if noise_std != 0)
 *   rng_u (B*N,S) uniform draw #5
 * out: z_fine (B,N,S); points_fine (B,N,S,3); inds (B*N,S) int64 searchsorted result or NULL */
int fenerf_resample(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse,
                    const float* z_vals, const float* dirs, const float* origins,
                    const float* rng_noise, const float* rng_u,
                    float* z_fine, float* points_fine, int64_t* inds, void* stream);

/* Merge-sort of coarse + fine samples, alpha compositing, fill modes, NCHW epilogue.
 * Replaces cat/sort/gather (generators/generators.py:85-89), the final fancy_integration
 * (volumetric_rendering.py:18-106) and the softmax / permute / *2-1 epilogue (:97-104).
 *   raw_fine / z_fine may be NULL when !hierarchical
 *   rng_noise (B,N,S') normal draw #6 (S' = 2S if hierarchical) or NULL
 * out: pixels (B, C_img, H, W) already *2-1, C_img = C-1 (+1 for the seg_padding fill modes)
 *      depth (B,N) or NULL;  weights_sum (B,N) or NULL;  weights (B,N,S') or NULL
 *      sort_idx (B,N,S') int32 merge order or NULL (debug / parity)                           */
int fenerf_composite(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse,
                     const float* z_coarse, const float* raw_fine, const float* z_fine,
                     const float* rng_noise, float* pixels, float* depth, float* weights_sum,
                     float* weights, int32_t* sort_idx, void* stream);

/* Scratch bytes fenerf_render_forward needs for this (render, field) pair. */
size_t fenerf_workspace_bytes(const fenerf_render_desc* rd, const fenerf_field_desc* field);

/* GUARD self-check.  The refinement pass knows, for every far sample it re-evaluates in fp32, what the tcgen05
 * density was: the largest |difference| and the number of wrong signs are the measured margin of `guard_tau` on
 * the weights and points actually rendered (the default tau was calibrated on the reference's random
 * initialisation; trained weights can have larger activations).  A refinement is only trustworthy while
 * max_abs_delta stays well below tau -- the host mirror raises tau and re-renders when it exceeds tau / 3.
 * Reads 16 bytes at the start of the workspace of the LAST fenerf_render_forward (precision GUARD) on `stream`
 * and synchronises that stream. */
typedef struct fenerf_guard_report {
    int32_t refined;        /* far samples re-evaluated (|sigma + noise| < tau) */
    float   max_abs_delta;  /* max |sigma_fp32 - sigma_tcgen05| over them */
    int32_t sign_flips;     /* of which the tcgen05 sign was wrong (these are what the GUARD fixes) */
    float   tau;            /* the threshold that was used */
} fenerf_guard_report;
int fenerf_guard_stats(const void* workspace, fenerf_guard_report* out, void* stream);

/* Where fenerf_render_forward leaves its intermediates inside the caller's workspace (byte offsets): the sample
 * points and depths of both passes, ray directions / origins and the raw field outputs -- what the backward
 * needs, and what a debugger wants to look at.  Valid after a fenerf_render_forward with the same (rd, field);
 * the fine entries only when rd->hierarchical. */
typedef struct fenerf_workspace_offsets {
    size_t points_coarse;  /* (B,N,S,3) */
    size_t z_coarse;       /* (B,N,S)   */
    size_t dirs;           /* (B,N,3)   */
    size_t origins;        /* (B,3)     */
    size_t raw_coarse;     /* (B,N,S,C) after the GUARD refinement */
    size_t z_fine;         /* (B,N,S)   */
    size_t points_fine;    /* (B,N,S,3) */
    size_t raw_fine;       /* (B,N,S,C) */
    size_t total;
} fenerf_workspace_offsets;
int fenerf_workspace_layout(const fenerf_render_desc* rd, const fenerf_field_desc* field, fenerf_workspace_offsets* out);

/* The whole per-batch render: ray_setup -> field(coarse) -> resample -> field(fine) -> composite.
 * Replaces the body of ImplicitGenerator3d.forward / DoubleImplicitGenerator3d.forward after the
 * mapping network (generators/generators.py:41-104, 465-527) and the chunked loops of
 * staged_forward* (:154-233, 569-646).
 *   rng_perturb (B,N,S) #1, rng_noise_c (B,N,S) #4 or NULL, rng_u (B*N,S) #5,
 *   rng_noise_f (B,N,2S) #6 or NULL  -- the caller draws them in the reference's order
 * out as fenerf_composite; inds_dbg as fenerf_resample.                                        */
int fenerf_render_forward(const fenerf_render_desc* rd, const fenerf_field_desc* field,
                          const void* packed, const float* film,
                          const float* x_lin, const float* y_lin, const float* z_lin,
                          const float* cam2world,
                          const float* rng_perturb, const float* rng_noise_c,
                          const float* rng_u, const float* rng_noise_f,
                          float* pixels, float* depth, float* weights_sum, float* weights,
                          int64_t* inds_dbg, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- mapping network -> FiLM table -------------------------------------------------------------------
 * CustomMappingNetwork (siren/siren.py:82-102: Linear(z, 256) + LeakyReLU(0.2), 3 x [Linear(256, 256) + LeakyReLU],
 * Linear(256, n_layers * 512)), the halves split into frequencies / phase shifts (:100-101), the `15 f + 30`
 * affine (siren.py:165) and, when the averages are given, the psi truncation of staged_forward
 * (generators.py:143-149): v = avg + psi (v - avg).  Writes layers [first_layer, first_layer + n_layers) of
 * film (B, n_film_total, 2, 256); a double-latent field calls it once per mapping network.  no_grad only.
 * h_scratch: min(B, 32) * 256 floats. */
typedef struct fenerf_mapping_params {
    const float* weight[5];   /* nn.Linear weights [out][in] fp32, in network order */
    const float* bias[5];
    int32_t z_dim;            /* multiple of 4, <= 512 */
    int32_t hidden_dim;       /* 256 */
} fenerf_mapping_params;
int fenerf_mapping_film(const fenerf_mapping_params* net, const float* z, int32_t batch, int32_t n_layers, int32_t first_layer,
                        int32_t n_film_total, const float* avg_frequencies, const float* avg_phase_shifts, float psi,
                        float* h_scratch, float* film, void* stream);

/* ---- frame consumers ---------------------------------------------------------------------------------
 * mask2color (train_double_latent_semantic.py:36-55, 66-72): masks (B, K, H, W) -> argmax over K -> the reference's
 * 19-entry colour table -> out (B, 3, H, W) float in 0..255 (classes >= 19 stay black, as in the reference). */
int fenerf_mask2color(const float* masks, int32_t batch, int32_t n_labels, int64_t pixels_per_image, float* out, void* stream);
/* frames (B, C, H, W) in [-1, 1] -> out (B, H, W, n_channels) uint8 of channels [first_channel, first_channel + n_channels):
 * torchvision.utils.save_image(img, normalize=True, range=(-1, 1)) rounding (fid_evaluation.py:146-151) -- the JPEG
 * encoder's input, a quarter of the bytes to move to the host. */
int fenerf_frames_to_u8(const float* frames, int32_t batch, int32_t channels, int32_t first_channel, int32_t n_channels,
                        int64_t pixels_per_image, uint8_t* out, void* stream);

/* ---- backward of the render ---------------------------------------------------------------------
 * What the reference differentiates (train_double_latent_semantic.py:405-446 G step,
 * inverse_render_double_semantic.py:385-407 inversion through forward_with_frequencies,
 * generators/generators.py:735-798): the final fancy_integration over the merged samples and both
 * point-network passes; ray set-up and resampling are no_grad there (generators.py:41, 59).
 * The host mirror (fenerf_b200/backward.py, a torch.autograd.Function) chains these entry points with the
 * plain 256-wide library GEMMs between them.                                                          */

/* element type of the backward's activation / gradient streams: fp16 (tensor-core GEMMs between the kernels, the
 * default) or fp32 (the parity mode that goes with FENERF_PRECISION_EXACT) */
#define FENERF_DTYPE_F16 0
#define FENERF_DTYPE_F32 1

/* The 256-wide products of a FiLM layer's backward on tcgen05 (csrc/gemm5.cu); fp16 row-major operands, fp32 accumulate.
 *   fenerf_gemm_nt_f16   C (M, 256) = A (M, 256) . B (256, 256)^T  -> c_f32 or c_f16 (exactly one non-NULL)
 *                        (dA' = dZ W: pass B = W^T); optional gate_mul (M, 256) fp16 multiplies the fp16 output in the
 *                        epilogue: dZ of the layer below = (dZ W) * its gate, without a pass of its own
 *   fenerf_gemm_nt_film  the recompute of a layer with its epilogue fused: z = A W^T never leaves the SM,
 *                        a_out = sin(f (z + bias) + p), gate_out = f cos(f (z + bias) + p), both (M, 256) fp16;
 *                        film_layer / film_batch_stride / points_per_batch as in fenerf_film_forward_stash; optional
 *                        narrow_in (M, 64) fp16 / narrow_w (256, 64) fp16, zero padded: a fifth k-chunk, z += narrow_in
 *                        narrow_w^T (the first colour layer's [dir, grid features] inputs, siren.py:1519-1522)
 *   fenerf_gemm_tn_f16   partial (batch, slices, 256, 256) fp32: for image b, slice s the sum over its 64-point stages
 *                        s, s + slices, ... of X[p, :]^T Y[p, :]  (dW_b = dZ^T a = the sum over the slices); optional
 *                        colsum (batch, slices, 256): column sums of X over the same stages (= d bias), computed by the
 *                        epilogue warps from the staged tiles while the tensor core runs                              */
int fenerf_gemm_nt_f16(const void* A, const void* B, int64_t M, float* c_f32, void* c_f16, const void* gate_mul, void* stream);
int fenerf_gemm_nt_film(const void* A, const void* W, int64_t M, const float* bias, const float* film_layer,
                        int64_t film_batch_stride, int64_t points_per_batch, const void* narrow_in, const void* narrow_w,
                        void* a_out, void* gate_out, void* stream);
int fenerf_gemm_tn_f16(const void* X, const void* Y, int32_t batch, int64_t points_per_batch, int32_t slices, float* partial,
                       float* colsum, void* stream);

/* d pixels (B, C-1, H, W) -> d raw outputs.  Backward of the merge + fancy_integration + softmax / *2-1
 * epilogue (generators.py:85-104, volumetric_rendering.py:18-50); same arguments as fenerf_composite.
 * d_raw_fine / raw_fine / z_fine NULL when !hierarchical.  fill modes are staged_forward-only (no_grad). */
int fenerf_composite_backward(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse,
                              const float* z_coarse, const float* raw_fine, const float* z_fine,
                              const float* rng_noise, const float* d_pixels,
                              float* d_raw_coarse, float* d_raw_fine, void* stream);

/* One FiLM layer's forward values for the backward (siren/siren.py:113-123): from the GEMM output
 * z (n_points, 256) fp32 (NULL for the first layer) plus optional narrow inputs narrow_in (n_points, w) fp32
 * against narrow_w (256, w) fp32 (positions; [dir, grid features] of the first colour layer):
 *   a = sin(f (z + bias) + p)  -> a_out (n_points, 256);   gate = f cos(f (z + bias) + p) -> gate_out   (both of `dtype`)
 * film_layer points at image 0's [2][256] block of the layer, film_batch_stride floats between images. */
int fenerf_film_forward_stash(const float* z, const float* bias, const float* film_layer, int64_t film_batch_stride,
                              int64_t n_points, int64_t points_per_batch, const float* narrow_in, int32_t narrow_width,
                              const float* narrow_w, void* a_out, void* gate_out, int32_t dtype, void* stream);

/* dZ = dA * gate in place ((n_points, 256) of `dtype`); colsum (B, 256) fp32 += per-image column sums of dZ
 * (= d bias; d phase = colsum / f; d freq follows from the per-image dW, see csrc/backward.cu). */
int fenerf_gate_backward(void* dA, const void* gate, int64_t n_points, int64_t points_per_batch, float* colsum,
                         int32_t dtype, void* stream);

/* d raw (n_points, C) -> scaled head gradients of `dtype`: d_heads (n_points, 32) = [d labels.., d sigma, 0..],
 * d_rgb (n_points, 8) = [d rgb * rgb (1 - rgb), 0..]; `scale` is a DEVICE scalar (power of two). */
int fenerf_head_grads(const float* d_raw, const float* raw, int64_t n_points, int32_t out_dim, int32_t label_dim,
                      const float* scale, void* d_heads, void* d_rgb, int32_t dtype, void* stream);

/* out (n_points, 3 + G) fp32 = [ray direction, trilinear grid features]: the narrow inputs of the first colour
 * layer (siren.py:1519-1522), from the packed channels-last grid. */
int fenerf_extras_gather(const fenerf_field_desc* field, const void* packed, const float* points, const float* dirs,
                         int64_t n_points, int64_t points_per_batch, int32_t dir_group, int32_t lock_dirs, float* out,
                         void* stream);

/* Backward of sample_from_3dgrid (siren.py:314-330): d features (n_points, ld >= 32) of `dtype` scattered with the
 * trilinear weights into grad_channels_last [R][R][R][32] fp32 (vector atomics); then the transpose back to
 * torch's (1, G, R, R, R) layout, multiplied by the DEVICE scalar *inv_scale. */
int fenerf_grid_scatter_add(const fenerf_field_desc* field, const float* points, const void* d_feat, int32_t ld,
                            int64_t n_points, float* grad_channels_last, int32_t dtype, void* stream);
int fenerf_grid_unpack_grad(const fenerf_field_desc* field, const float* grad_channels_last, float* out,
                            const float* inv_scale, void* stream);

/* Per-thread message for the last non-zero return. */
const char* fenerf_last_error(void);

/* ABI version and a counter of kernels launched by this library since load (bench evidence). */
int32_t fenerf_abi_version(void);
int64_t fenerf_launch_count(void);

/* Diagnostics: CUDA-event timing of the six stages of fenerf_render_forward (ray set-up, coarse field, GUARD
 * refinement, resampling, fine field, compositing).  enable = 1 with ms_out NULL switches it on, enable = 0 off;
 * enable = 1 with ms_out reads the six durations (ms) of the LAST call (synchronises on it).  Process-wide, not
 * thread-safe, inactive during CUDA-graph capture. */
int fenerf_debug_stage_times(int32_t enable, float* ms_out /* host, 6 floats */);

/* Diagnostics: install a device buffer of 4 * 4096 int64 that CTA 0 of the tcgen05 point-network
 * kernel fills with (tag, clock64) pairs for its first two tile pairs, one 4096-entry lane per warp
 * role (producer, MMA issuer, epilogue X, epilogue Y); NULL (the default) disables.  Process-wide,
 * not thread-safe. */
void fenerf_debug_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* FENERF_B200_H */
