// extern "C" surface of libfenerf_b200 (include/fenerf_b200.h): argument validation, precision
// dispatch and the five-launch render pipeline.  No torch types, no allocation, no global state
// beyond the thread-local error string and the launch counter.
#include "common.cuh"
#include <stdlib.h>

namespace fn {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace fn

using namespace fn;

namespace {

struct Workspace {
    size_t stats, points_c, z_c, dirs, origins, raw_c, z_f, points_f, raw_f, guard, sigma_c, total;
};

Workspace plan_workspace(const fenerf_render_desc* rd, int C) {
    Workspace w;
    size_t n_rays = (size_t)rd->batch * rd->img_h * rd->img_w;
    size_t pc = n_rays * rd->num_steps;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = fn_align_up(off + bytes, 256); return o; };
    w.stats = take(64);               // always at offset 0: fenerf_guard_stats
    w.points_c = take(pc * 3 * 4);
    w.z_c = take(pc * 4);
    w.dirs = take(n_rays * 3 * 4);
    w.origins = take((size_t)rd->batch * 3 * 4);
    w.raw_c = take(pc * C * 4);
    w.z_f = take(rd->hierarchical ? pc * 4 : 4);
    w.points_f = take(rd->hierarchical ? pc * 3 * 4 : 4);
    w.raw_f = take(rd->hierarchical ? pc * C * 4 : 4);
    w.guard = take((n_rays + 1) * 4);
    w.sigma_c = take(pc * 4);
    w.total = off;
    return w;
}

int check_render_desc(const fenerf_render_desc* rd) {
    FN_REQUIRE(rd, "render desc is NULL");
    FN_REQUIRE(rd->batch >= 1 && rd->img_h >= 1 && rd->img_w >= 1, "bad batch/img size %d %dx%d", rd->batch, rd->img_h,
               rd->img_w);
    FN_REQUIRE(rd->num_steps >= 2 && rd->num_steps <= 64, "num_steps %d outside [2, 64]", rd->num_steps);
    if (rd->clamp_mode != FENERF_CLAMP_RELU && rd->clamp_mode != FENERF_CLAMP_SOFTPLUS)
        return fail(FENERF_E_CLAMP_MODE, "Need to choose clamp mode");
    FN_REQUIRE(rd->fill_mode >= FENERF_FILL_NONE && rd->fill_mode <= FENERF_FILL_EVAL_WHITE_BACK, "unknown fill_mode %d",
               rd->fill_mode);
    FN_REQUIRE(rd->precision >= FENERF_PRECISION_EXACT && rd->precision <= FENERF_PRECISION_GUARD, "unknown precision %d",
               rd->precision);
    return 0;
}

// ---- diagnostics: per-stage CUDA-event timing of fenerf_render_forward (warm, in-step numbers; ncu's are cold) ----
constexpr int kStages = 6;       // ray_setup, field(coarse), guard, resample, field(fine), composite
bool g_stage_timing = false;
cudaEvent_t g_stage_ev[kStages + 1];

void stage_mark(int i, cudaStream_t st) {
    if (!g_stage_timing) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    cudaEventRecord(g_stage_ev[i], st);
}

int run_field(const FnLayout& L, const void* packed, const float* points, const float* dirs, const float* film,
              int batch, long long ppb, int dir_group, int lock_dirs, int precision, float* out, cudaStream_t st,
              int sigma_only = 0, float* sigma_out = nullptr) {
    const unsigned char* pk = static_cast<const unsigned char*>(packed);
    if (precision == FENERF_PRECISION_EXACT)
        return siren_points_exact(L, pk, points, dirs, film, batch, ppb, dir_group, lock_dirs, nullptr, 0, out, st, sigma_only);
    // the tcgen05 kernel (siren_fast3.cu: third generation; its predecessors are described in DESIGN.md
    // section 5 and live in the history only)
    return siren_points_fast3(L, pk, points, dirs, film, batch, ppb, dir_group, lock_dirs, out, get_fast_trace(), sigma_only, st,
                              sigma_out);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* fenerf_last_error(void) { return g_err; }
int32_t fenerf_abi_version(void) { return FENERF_ABI_VERSION; }
void fenerf_debug_trace(void* device_buffer) { set_fast_trace(static_cast<long long*>(device_buffer)); }
int64_t fenerf_launch_count(void) { return (int64_t)g_launches.load(); }

size_t fenerf_packed_bytes(const fenerf_field_desc* field) {
    FnLayout L;
    if (fn_make_layout(field, &L) != 0) {
        fail(FENERF_E_ARG, "unsupported field description");
        return 0;
    }
    return L.total;
}

int fenerf_pack_field(const fenerf_field_desc* field, const fenerf_field_params* params, void* packed,
                      size_t packed_bytes, void* stream) {
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(params && packed, "params/packed is NULL");
    FN_REQUIRE(((uintptr_t)packed & 1023) == 0, "packed buffer must be 1024-byte aligned");
    if (packed_bytes < L.total) return fail(FENERF_E_WORKSPACE, "packed buffer too small: %zu < %zu", packed_bytes, L.total);
    return pack_field(field, L, params, packed, (cudaStream_t)stream);
}

int fenerf_field_fingerprint(const fenerf_field_desc* field, const fenerf_field_params* params, uint64_t* out,
                             void* stream) {
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(params && out, "params/out is NULL");
    FN_REQUIRE(((uintptr_t)out & 7) == 0, "out must be 8-byte aligned");
    return field_fingerprint(L, params, reinterpret_cast<unsigned long long*>(out), (cudaStream_t)stream);
}

int fenerf_siren_points(const fenerf_field_desc* field, const void* packed, const float* points, const float* dirs,
                        const float* film, int32_t batch, int64_t points_per_batch, int32_t dir_group,
                        int32_t precision, const int32_t* only_idx, int32_t n_only, float* out, void* stream) {
    FnLayout L;
    const int sigma_only = (precision & FENERF_POINTS_SIGMA_ONLY) ? 1 : 0;
    precision &= 0xff;
    FN_REQUIRE(precision >= FENERF_PRECISION_EXACT && precision <= FENERF_PRECISION_GUARD, "unknown precision %d", precision);
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(packed && points && film && out, "NULL argument");
    FN_REQUIRE(dirs, "dirs is NULL (pass any (B,P/dir_group,3) tensor; the colour branch consumes it)");
    FN_REQUIRE(batch >= 1 && points_per_batch >= 1 && dir_group >= 1, "bad sizes");
    FN_REQUIRE(points_per_batch % dir_group == 0, "points_per_batch must be a multiple of dir_group");
    cudaStream_t st = (cudaStream_t)stream;
    if (only_idx) {
        if (n_only <= 0) return 0;
        return siren_points_exact(L, (const unsigned char*)packed, points, dirs, film, batch, points_per_batch, dir_group,
                                  0, only_idx, n_only, out, st);
    }
    FN_REQUIRE(precision >= FENERF_PRECISION_EXACT && precision <= FENERF_PRECISION_GUARD, "unknown precision %d", precision);
    return run_field(L, packed, points, dirs, film, batch, points_per_batch, dir_group, 0, precision, out, st, sigma_only);
}

int fenerf_camera_poses(int32_t n, int32_t mode, float h_stddev, float v_stddev, float h_mean, float v_mean,
                        const float* draw_theta, const float* draw_phi, float* cam2world, float* pitch, float* yaw,
                        void* stream) {
    FN_REQUIRE(n >= 1 && cam2world && pitch && yaw, "bad argument");
    FN_REQUIRE(mode >= FENERF_CAMERA_FIXED && mode <= FENERF_CAMERA_SPHERICAL_UNIFORM, "unsupported camera mode %d", mode);
    FN_REQUIRE(mode == FENERF_CAMERA_FIXED || (draw_theta && draw_phi), "camera mode %d needs the two random draws", mode);
    return camera_poses(n, mode, h_stddev, v_stddev, h_mean, v_mean, draw_theta, draw_phi, cam2world, pitch, yaw,
                        (cudaStream_t)stream);
}

int fenerf_ray_setup(const fenerf_render_desc* rd, const float* x_lin, const float* y_lin, const float* z_lin,
                     const float* cam2world, const float* rng_perturb, float* points, float* z_vals, float* dirs,
                     float* origins, void* stream) {
    if (int e = check_render_desc(rd)) return e;
    FN_REQUIRE(x_lin && y_lin && z_lin && cam2world && rng_perturb && points && z_vals && dirs && origins, "NULL argument");
    return ray_setup(rd, x_lin, y_lin, z_lin, cam2world, rng_perturb, points, z_vals, dirs, origins, (cudaStream_t)stream);
}

int fenerf_resample(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse, const float* z_vals,
                    const float* dirs, const float* origins, const float* rng_noise, const float* rng_u, float* z_fine,
                    float* points_fine, int64_t* inds, void* stream) {
    if (int e = check_render_desc(rd)) return e;
    FN_REQUIRE(raw_coarse && z_vals && dirs && origins && rng_u && z_fine && points_fine, "NULL argument");
    FN_REQUIRE(out_dim >= 2 && out_dim <= 36, "out_dim %d unsupported", out_dim);
    FN_REQUIRE(rd->noise_std == 0.f || rng_noise, "noise_std != 0 needs rng_noise");
    return resample(rd, out_dim, raw_coarse, z_vals, dirs, origins, rd->noise_std != 0.f ? rng_noise : nullptr, rng_u,
                    z_fine, points_fine, (long long*)inds, (cudaStream_t)stream);
}

int fenerf_composite(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse, const float* z_coarse,
                     const float* raw_fine, const float* z_fine, const float* rng_noise, float* pixels, float* depth,
                     float* weights_sum, float* weights, int32_t* sort_idx, void* stream) {
    if (int e = check_render_desc(rd)) return e;
    FN_REQUIRE(raw_coarse && z_coarse && pixels, "NULL argument");
    FN_REQUIRE(rd->noise_std == 0.f || rng_noise, "noise_std != 0 needs rng_noise");
    return composite(rd, out_dim, raw_coarse, z_coarse, raw_fine, z_fine, rd->noise_std != 0.f ? rng_noise : nullptr,
                     pixels, depth, weights_sum, weights, sort_idx, (cudaStream_t)stream);
}

size_t fenerf_workspace_bytes(const fenerf_render_desc* rd, const fenerf_field_desc* field) {
    if (!rd || !field) return 0;
    return plan_workspace(rd, field->out_dim).total;
}

int fenerf_debug_stage_times(int32_t enable, float* ms_out) {
    if (enable && !g_stage_timing) {
        for (int i = 0; i <= kStages; ++i) FN_CUDA_OK(cudaEventCreate(&g_stage_ev[i]));
        g_stage_timing = true;
        return 0;
    }
    if (!enable && g_stage_timing) {
        g_stage_timing = false;
        for (int i = 0; i <= kStages; ++i) cudaEventDestroy(g_stage_ev[i]);
        return 0;
    }
    if (g_stage_timing && ms_out) {
        FN_CUDA_OK(cudaEventSynchronize(g_stage_ev[kStages]));
        for (int i = 0; i < kStages; ++i) FN_CUDA_OK(cudaEventElapsedTime(ms_out + i, g_stage_ev[i], g_stage_ev[i + 1]));
    }
    return 0;
}

int fenerf_guard_stats(const void* workspace, fenerf_guard_report* out, void* stream) {
    FN_REQUIRE(workspace && out, "NULL argument");
    int32_t raw[4];
    FN_CUDA_OK(cudaMemcpyAsync(raw, workspace, sizeof(raw), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    FN_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    out->refined = raw[0];
    out->max_abs_delta = *reinterpret_cast<const float*>(&raw[1]);
    out->sign_flips = raw[2];
    out->tau = *reinterpret_cast<const float*>(&raw[3]);
    return 0;
}

int fenerf_workspace_layout(const fenerf_render_desc* rd, const fenerf_field_desc* field, fenerf_workspace_offsets* out) {
    if (int e = check_render_desc(rd)) return e;
    FN_REQUIRE(field && out, "NULL argument");
    Workspace w = plan_workspace(rd, field->out_dim);
    out->points_coarse = w.points_c; out->z_coarse = w.z_c; out->dirs = w.dirs; out->origins = w.origins;
    out->raw_coarse = w.raw_c; out->z_fine = w.z_f; out->points_fine = w.points_f; out->raw_fine = w.raw_f;
    out->total = w.total;
    return 0;
}

int fenerf_render_forward(const fenerf_render_desc* rd, const fenerf_field_desc* field, const void* packed,
                          const float* film, const float* x_lin, const float* y_lin, const float* z_lin,
                          const float* cam2world, const float* rng_perturb, const float* rng_noise_c,
                          const float* rng_u, const float* rng_noise_f, float* pixels, float* depth,
                          float* weights_sum, float* weights, int64_t* inds_dbg, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (int e = check_render_desc(rd)) return e;
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(packed && film && x_lin && y_lin && z_lin && cam2world && rng_perturb && pixels && workspace, "NULL argument");
    FN_REQUIRE(!rd->hierarchical || rng_u, "hierarchical render needs rng_u");
    FN_REQUIRE(!rd->hierarchical || rd->num_steps >= 3, "hierarchical render needs num_steps >= 3");
    FN_REQUIRE(rd->noise_std == 0.f || (rng_noise_f && (!rd->hierarchical || rng_noise_c)), "noise_std != 0 needs the noise draws");
    FN_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    const int C = L.out_dim;
    Workspace w = plan_workspace(rd, C);
    if (workspace_bytes < w.total) return fail(FENERF_E_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, w.total);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float* points_c = (float*)(ws + w.points_c);
    float* z_c = (float*)(ws + w.z_c);
    float* dirs = (float*)(ws + w.dirs);
    float* origins = (float*)(ws + w.origins);
    float* raw_c = (float*)(ws + w.raw_c);
    float* z_f = (float*)(ws + w.z_f);
    float* points_f = (float*)(ws + w.points_f);
    float* raw_f = (float*)(ws + w.raw_f);
    int32_t* guard = (int32_t*)(ws + w.guard);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rays = (long long)rd->img_h * rd->img_w;
    const long long ppb = rays * rd->num_steps;
    const float* noise_c = rd->noise_std != 0.f ? rng_noise_c : nullptr;
    const float* noise_f = rd->noise_std != 0.f ? rng_noise_f : nullptr;

    stage_mark(0, st);
    if (int e = ray_setup(rd, x_lin, y_lin, z_lin, cam2world, rng_perturb, points_c, z_c, dirs, origins, st)) return e;
    stage_mark(1, st);
    // the tcgen05 pass also leaves the densities as one float per point for the resampler (which never reads the far
    // sample, so the GUARD refinement of raw_c below does not concern that copy)
    float* sigma_c = (rd->hierarchical && rd->precision != FENERF_PRECISION_EXACT) ? (float*)(ws + w.sigma_c) : nullptr;
    if (int e = run_field(L, packed, points_c, dirs, film, rd->batch, ppb, rd->num_steps, rd->lock_view_dependence,
                          rd->precision, raw_c, st, 0, sigma_c)) return e;
    stage_mark(2, st);
    if (rd->precision == FENERF_PRECISION_GUARD) {
        float tau = rd->guard_tau > 0.f ? rd->guard_tau : 1.5e-3f;
        const int n_samples = rd->hierarchical ? 2 * rd->num_steps : rd->num_steps;
        if (int e = guard_refine(L, (const unsigned char*)packed, points_c, dirs, film, rd->batch, rays, rd->num_steps,
                                 rd->lock_view_dependence, tau, noise_f ? noise_f + (n_samples - 1) : nullptr, n_samples,
                                 rd->noise_std, raw_c, guard, (int32_t*)(ws + w.stats), st)) return e;
    }
    stage_mark(3, st);
    if (rd->hierarchical) {
        if (int e = resample(rd, C, raw_c, z_c, dirs, origins, noise_c, rng_u, z_f, points_f, (long long*)inds_dbg, st,
                             /*sort_fine=*/1, sigma_c)) return e;
        stage_mark(4, st);
        if (int e = run_field(L, packed, points_f, dirs, film, rd->batch, ppb, rd->num_steps, rd->lock_view_dependence,
                              rd->precision, raw_f, st)) return e;
    } else {
        stage_mark(4, st);
    }
    stage_mark(5, st);
    // both sample lists are depth-sorted here: one thread per ray, accumulators in registers (composite.cu)
    const int rc = composite_sorted(rd, C, raw_c, z_c, rd->hierarchical ? raw_f : nullptr, rd->hierarchical ? z_f : nullptr,
                                    noise_f, pixels, depth, weights_sum, weights, st);
    stage_mark(6, st);
    return rc;
}

int fenerf_mapping_film(const fenerf_mapping_params* net, const float* z, int32_t batch, int32_t n_layers, int32_t first_layer,
                        int32_t n_film_total, const float* avg_frequencies, const float* avg_phase_shifts, float psi,
                        float* h_scratch, float* film, void* stream) {
    FN_REQUIRE(net && z && h_scratch && film && batch >= 1 && n_layers >= 1 && first_layer >= 0 &&
               first_layer + n_layers <= n_film_total, "bad argument");
    FN_REQUIRE((avg_frequencies == nullptr) == (avg_phase_shifts == nullptr), "give both averages or neither");
    for (int i = 0; i < 5; ++i) FN_REQUIRE(net->weight[i] && net->bias[i], "mapping layer %d missing", i);
    FN_REQUIRE(net->hidden_dim == 256, "mapping network hidden width must be 256");
    return mapping_film(net->weight, net->bias, z, batch, net->z_dim, n_layers, first_layer, n_film_total, avg_frequencies,
                        avg_phase_shifts, psi, h_scratch, film, (cudaStream_t)stream);
}

// ---- frame consumers --------------------------------------------------------
int fenerf_mask2color(const float* masks, int32_t batch, int32_t n_labels, int64_t pixels_per_image, float* out, void* stream) {
    FN_REQUIRE(masks && out && batch >= 1 && n_labels >= 1 && pixels_per_image >= 1, "bad argument");
    return mask2color(masks, batch, n_labels, pixels_per_image, out, (cudaStream_t)stream);
}

int fenerf_frames_to_u8(const float* frames, int32_t batch, int32_t channels, int32_t first_channel, int32_t n_channels,
                        int64_t pixels_per_image, uint8_t* out, void* stream) {
    FN_REQUIRE(frames && out && batch >= 1 && n_channels >= 1 && first_channel >= 0 && first_channel + n_channels <= channels &&
               pixels_per_image >= 1, "bad argument");
    return frames_to_u8(frames, batch, channels, first_channel, n_channels, pixels_per_image, out, (cudaStream_t)stream);
}

// ---- backward ---------------------------------------------------------------
int fenerf_gemm_nt_f16(const void* A, const void* B, int64_t M, float* c_f32, void* c_f16, const void* gate_mul, void* stream) {
    FN_REQUIRE(A && B && M >= 0 && ((c_f32 != nullptr) != (c_f16 != nullptr)), "bad argument (exactly one of c_f32 / c_f16)");
    FN_REQUIRE(!gate_mul || c_f16, "gate_mul goes with the fp16 output");
    FN_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)c_f32 | (uintptr_t)c_f16 | (uintptr_t)gate_mul) & 15) == 0, "operands must be 16-byte aligned");
    return gemm_nt(A, B, M, c_f32, c_f16, nullptr, nullptr, nullptr, nullptr, 0, 1, (cudaStream_t)stream, gate_mul);
}

int fenerf_gemm_nt_film(const void* A, const void* W, int64_t M, const float* bias, const float* film_layer,
                        int64_t film_batch_stride, int64_t points_per_batch, const void* narrow_in, const void* narrow_w,
                        void* a_out, void* gate_out, void* stream) {
    FN_REQUIRE(A && W && bias && film_layer && a_out && gate_out && M >= 0 && points_per_batch >= 1, "bad argument");
    FN_REQUIRE((narrow_in == nullptr) == (narrow_w == nullptr), "narrow_in and narrow_w go together");
    FN_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)a_out | (uintptr_t)gate_out | (uintptr_t)narrow_in | (uintptr_t)narrow_w) & 15) == 0,
               "operands must be 16-byte aligned");
    return gemm_nt(A, W, M, nullptr, nullptr, a_out, gate_out, bias, film_layer, film_batch_stride, points_per_batch,
                   (cudaStream_t)stream, nullptr, narrow_in, narrow_w);
}

int fenerf_gemm_tn_f16(const void* X, const void* Y, int32_t batch, int64_t points_per_batch, int32_t slices, float* partial,
                       float* colsum, void* stream) {
    FN_REQUIRE(X && Y && partial && batch >= 1 && points_per_batch >= 1 && slices >= 1, "bad argument");
    FN_REQUIRE((((uintptr_t)X | (uintptr_t)Y | (uintptr_t)partial) & 15) == 0, "operands must be 16-byte aligned");
    return gemm_tn(X, Y, batch, points_per_batch, slices, partial, (cudaStream_t)stream, colsum);
}

int fenerf_composite_backward(const fenerf_render_desc* rd, int32_t out_dim, const float* raw_coarse, const float* z_coarse,
                              const float* raw_fine, const float* z_fine, const float* rng_noise, const float* d_pixels,
                              float* d_raw_coarse, float* d_raw_fine, void* stream) {
    if (int e = check_render_desc(rd)) return e;
    FN_REQUIRE(raw_coarse && z_coarse && d_pixels && d_raw_coarse, "NULL argument");
    FN_REQUIRE(rd->noise_std == 0.f || rng_noise, "noise_std != 0 needs rng_noise");
    return composite_backward(rd, out_dim, raw_coarse, z_coarse, raw_fine, z_fine, rd->noise_std != 0.f ? rng_noise : nullptr,
                              d_pixels, d_raw_coarse, d_raw_fine, (cudaStream_t)stream);
}

int fenerf_film_forward_stash(const float* z, const float* bias, const float* film_layer, int64_t film_batch_stride,
                              int64_t n_points, int64_t points_per_batch, const float* narrow_in, int32_t narrow_width,
                              const float* narrow_w, void* a_out, void* gate_out, int32_t dtype, void* stream) {
    FN_REQUIRE(bias && film_layer && a_out && gate_out && n_points > 0 && points_per_batch > 0, "bad argument");
    FN_REQUIRE(narrow_width == 0 || (narrow_in && narrow_w), "narrow inputs missing");
    FN_REQUIRE(dtype == FENERF_DTYPE_F16 || dtype == FENERF_DTYPE_F32, "dtype");
    return film_forward_stash(z, bias, film_layer, film_batch_stride, n_points, points_per_batch, narrow_in, narrow_width,
                              narrow_w, a_out, gate_out, dtype, (cudaStream_t)stream);
}

int fenerf_gate_backward(void* dA, const void* gate, int64_t n_points, int64_t points_per_batch, float* colsum, int32_t dtype,
                         void* stream) {
    FN_REQUIRE(dA && gate && colsum && n_points > 0 && points_per_batch > 0, "bad argument");
    FN_REQUIRE(dtype == FENERF_DTYPE_F16 || dtype == FENERF_DTYPE_F32, "dtype");
    return gate_backward(dA, gate, n_points, points_per_batch, colsum, dtype, (cudaStream_t)stream);
}

int fenerf_head_grads(const float* d_raw, const float* raw, int64_t n_points, int32_t out_dim, int32_t label_dim,
                      const float* scale, void* d_heads, void* d_rgb, int32_t dtype, void* stream) {
    FN_REQUIRE(d_raw && raw && scale && d_heads && d_rgb && n_points > 0, "bad argument");
    FN_REQUIRE(dtype == FENERF_DTYPE_F16 || dtype == FENERF_DTYPE_F32, "dtype");
    return head_grads(d_raw, raw, n_points, out_dim, label_dim, scale, d_heads, d_rgb, dtype, (cudaStream_t)stream);
}

int fenerf_extras_gather(const fenerf_field_desc* field, const void* packed, const float* points, const float* dirs,
                         int64_t n_points, int64_t points_per_batch, int32_t dir_group, int32_t lock_dirs, float* out,
                         void* stream) {
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(packed && points && dirs && out && n_points > 0 && points_per_batch > 0 && dir_group >= 1, "bad argument");
    return extras_gather(L, (const unsigned char*)packed, points, dirs, n_points, points_per_batch, dir_group, lock_dirs, out,
                         (cudaStream_t)stream);
}

int fenerf_grid_scatter_add(const fenerf_field_desc* field, const float* points, const void* d_feat, int32_t ld,
                            int64_t n_points, float* grad_channels_last, int32_t dtype, void* stream) {
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0, "unsupported field description");
    FN_REQUIRE(points && d_feat && grad_channels_last && n_points > 0 && ld >= 32 && ld % 8 == 0, "bad argument");
    FN_REQUIRE(dtype == FENERF_DTYPE_F16 || dtype == FENERF_DTYPE_F32, "dtype");
    return grid_scatter_add(L, points, d_feat, ld, n_points, grad_channels_last, dtype, (cudaStream_t)stream);
}

int fenerf_grid_unpack_grad(const fenerf_field_desc* field, const float* grad_channels_last, float* out,
                            const float* inv_scale, void* stream) {
    FnLayout L;
    FN_REQUIRE(fn_make_layout(field, &L) == 0 && L.grid_channels > 0, "field has no grid");
    FN_REQUIRE(grad_channels_last && out && inv_scale, "bad argument");
    return grid_unpack_grad(L, grad_channels_last, out, inv_scale, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
