// fenerf_pack_field: re-lay a field's raw torch parameters into the kernel layout (layout.h).
//
// Runs whenever the parameters change (optimizer step / EMA copy_to), so everything here is a
// handful of bandwidth-bound kernels: ~3.2 MB of weights plus, for the texture-embedding field,
// one channel-major -> channels-last transpose of the 113 MB feature grid (siren/siren.py:1546
// stores one voxel's 32 channels 3.5 MB apart; the trilinear gather wants them in one 128 B line).
#include "common.cuh"

namespace fn {

namespace {

__device__ __forceinline__ __half f16_hi(float w) { return __float2half_rn(w); }
__device__ __forceinline__ __half f16_lo(float w) { return __float2half_rn(w - __half2float(__float2half_rn(w))); }

// ---- first layer: Wt0, b0 and the input-chunk image (hi, hi, lo) -------------------------------
__global__ void pack_first_kernel(const float* __restrict__ w /*[256][3]*/, const float* __restrict__ b,
                                  float* __restrict__ wt, float* __restrict__ bo, unsigned char* __restrict__ img) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= FN_H) return;
    bo[n] = b[n];
    for (int k = 0; k < FN_KCHUNK; ++k) *reinterpret_cast<__half*>(img + fn_sw128_offset(n, k)) = __float2half_rn(0.f);
    for (int k = 0; k < 3; ++k) {
        float v = w[n * 3 + k];
        wt[k * FN_H + n] = v;
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_POS + k)) = f16_hi(v);
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_POS + 3 + k)) = f16_hi(v);
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_POS + 6 + k)) = f16_lo(v);
    }
}

// ---- hidden layer: k-major f32 copy (+ extra rows) and swizzled f16 images ---------------------
// grid: (K_total/16, 256/32 ... ) simple 2-D mapping: one thread per (k, n).
__global__ void pack_hidden_kernel(const float* __restrict__ w, const float* __restrict__ b, int in_dim, int x_off,
                                   int kx, int kx_pad, float* __restrict__ wt, float* __restrict__ bo,
                                   unsigned char* __restrict__ img) {
    // tile of 32 k x 32 n through shared memory so that both the read (k contiguous) and the
    // write (n contiguous) are coalesced
    __shared__ float tile[32][33];
    int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    int k_total = FN_H + kx_pad;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, k = k0 + threadIdx.x;
        float v = 0.f;
        if (k < FN_H) v = w[(size_t)n * in_dim + x_off + k];
        else if (k - FN_H < kx) v = w[(size_t)n * in_dim + (k - FN_H)];
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int k = k0 + i, n = n0 + threadIdx.x;
        if (k < k_total) wt[(size_t)k * FN_H + n] = tile[threadIdx.x][i];
    }
    // f16 image: thread (x = k within tile, y -> n rows)
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int n = n0 + i, k = k0 + threadIdx.x;
        if (k < FN_H)   // image order [feature half][k-chunk][128 rows][64 k]: a half's four k-chunks are contiguous
            *reinterpret_cast<__half*>(img + fn_hidden_img_offset(n, k)) = __float2half_rn(tile[i][threadIdx.x]);
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) bo[n0 + threadIdx.x] = b[n0 + threadIdx.x];
}

// input-chunk image of the first colour layer: dir (hi, hi, lo) in slots 16.., feat in 32..
__global__ void pack_color0_ximg_kernel(const float* __restrict__ w, int in_dim, int g, unsigned char* __restrict__ img) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= FN_H) return;
    for (int k = 0; k < FN_KCHUNK; ++k) *reinterpret_cast<__half*>(img + fn_sw128_offset(n, k)) = __float2half_rn(0.f);
    for (int k = 0; k < 3; ++k) {
        float v = w[(size_t)n * in_dim + k];
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_DIR + k)) = f16_hi(v);
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_DIR + 3 + k)) = f16_hi(v);
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_DIR + 6 + k)) = f16_lo(v);
    }
    for (int c = 0; c < g; ++c)
        *reinterpret_cast<__half*>(img + fn_sw128_offset(n, FN_SLOT_FEAT + c)) = __float2half_rn(w[(size_t)n * in_dim + 3 + c]);
}

__global__ void pack_heads_kernel(const float* __restrict__ sw, const float* __restrict__ sb,
                                  const float* __restrict__ rw, const float* __restrict__ rb,
                                  float* __restrict__ so, float* __restrict__ ro) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < FN_H) so[i] = sw[i];
    if (i == 0) so[FN_H] = sb[0];
    if (i < 3 * FN_H) ro[i] = rw[i];
    if (i < 3) ro[3 * FN_H + i] = rb[i];
}

// ---- label chain: Weff = W3 W2 W1, beff = W3 (W2 b1 + b2) + b3, in double ---------------------
// step 1: U = W3 W2 (L x 256), ub = W3 b2 + b3          grid L blocks x 256 threads
__global__ void label_step1_kernel(const float* __restrict__ w3, const float* __restrict__ b3,
                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                   double* __restrict__ u /*[L][257]*/) {
    int i = blockIdx.x, j = threadIdx.x;
    if (!w2) {                      // two-layer chain (siren.py:1189-1191): the middle layer is the identity
        u[i * (FN_H + 1) + j] = (double)w3[i * FN_H + j];
        if (j == 0) u[i * (FN_H + 1) + FN_H] = (double)b3[i];
        return;
    }
    double acc = 0.0;
    for (int m = 0; m < FN_H; ++m) acc += (double)w3[i * FN_H + m] * (double)w2[m * FN_H + j];
    u[i * (FN_H + 1) + j] = acc;
    if (j == 0) {
        double bb = (double)b3[i];
        for (int m = 0; m < FN_H; ++m) bb += (double)w3[i * FN_H + m] * (double)b2[m];
        u[i * (FN_H + 1) + FN_H] = bb;
    }
}
// step 2: Weff = U W1, beff = U b1 + ub
__global__ void label_step2_kernel(const double* __restrict__ u, const float* __restrict__ w1,
                                   const float* __restrict__ b1, int L, float* __restrict__ out /*[32*256 + 32 + 1]*/) {
    int i = blockIdx.x, j = threadIdx.x;
    double acc = 0.0;
    for (int m = 0; m < FN_H; ++m) acc += u[i * (FN_H + 1) + m] * (double)w1[m * FN_H + j];
    out[i * FN_H + j] = (float)acc;
    if (j == 0) {
        double bb = u[i * (FN_H + 1) + FN_H];
        for (int m = 0; m < FN_H; ++m) bb += u[i * (FN_H + 1) + m] * (double)b1[m];
        out[FENERF_MAX_LABEL * FN_H + i] = (float)bb;
    }
}
// step 3 (one block): power-of-two scale so that max|Weff| lands in [0.25, 0.5) -- random-init
// products of three 1/25-scaled layers sit in the fp16 subnormal range otherwise; 1/scale is stored
// after the biases for the epilogue.
__global__ void label_step3_kernel(float* __restrict__ lw, int L) {
    __shared__ float smax[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < L * FN_H; i += blockDim.x) m = fmaxf(m, fabsf(lw[i]));
    smax[threadIdx.x] = m;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
        __syncthreads();
    }
    float mx = smax[0];
    int e = 0;
    if (mx > 0.f && isfinite(mx)) { frexpf(mx, &e); }   // mx = f * 2^e, f in [0.5, 1)
    float scale = ldexpf(1.f, -e - 1);                    // mx * scale in [0.25, 0.5)
    if (!(mx > 0.f)) scale = 1.f;
    if (threadIdx.x == 0) lw[FENERF_MAX_LABEL * FN_H + FENERF_MAX_LABEL] = 1.f / scale;
}

// head images for the tcgen05 kernel (one block):
//   trunk head  [4][32 rows][64 k]: rows 0..L-1 = Weff * scale, row L = sigma weights
//   rgb head    [4][ 8 rows][64 k]: rows 0..2
__global__ void pack_head_imgs_kernel(const float* __restrict__ lw, int L, const float* __restrict__ sigma_w,
                                      const float* __restrict__ rgb_w, unsigned char* __restrict__ head_img,
                                      unsigned char* __restrict__ rgb_img) {
    const float scale = L > 0 ? 1.f / lw[FENERF_MAX_LABEL * FN_H + FENERF_MAX_LABEL] : 1.f;
    for (int i = threadIdx.x; i < 32 * FN_H; i += blockDim.x) {
        int row = i / FN_H, k = i % FN_H;
        float v = 0.f;
        if (row < L) v = lw[row * FN_H + k] * scale;
        else if (row == L) v = sigma_w[k];
        unsigned char* chunk = head_img + (size_t)(k / FN_KCHUNK) * (32 * FN_KCHUNK * 2);
        *reinterpret_cast<__half*>(chunk + fn_sw128_offset(row, k % FN_KCHUNK)) = __float2half_rn(v);
    }
    for (int i = threadIdx.x; i < 8 * FN_H; i += blockDim.x) {
        int row = i / FN_H, k = i % FN_H;
        float v = row < 3 ? rgb_w[row * FN_H + k] : 0.f;
        unsigned char* chunk = rgb_img + (size_t)(k / FN_KCHUNK) * (8 * FN_KCHUNK * 2);
        *reinterpret_cast<__half*>(chunk + fn_sw128_offset(row, k % FN_KCHUNK)) = __float2half_rn(v);
    }
}

// ---- grid: (G, D, H, W) channel-major -> [D][H][W][G] channels-last ---------------------------
// one block per (z, y) line: read G rows of R contiguous x, write R*G contiguous floats
__global__ void pack_grid_kernel(const float* __restrict__ in, float* __restrict__ out, __half* __restrict__ out16, int R, int G) {
    extern __shared__ float line[];  // [G][R + 1]
    size_t zy = blockIdx.x;          // z * R + y
    size_t plane = (size_t)R * R * R;
    for (int i = threadIdx.x; i < G * R; i += blockDim.x) {
        int c = i / R, x = i % R;
        line[c * (R + 1) + x] = in[(size_t)c * plane + zy * R + x];
    }
    __syncthreads();
    float* dst = out + zy * R * G;
    __half* dst16 = out16 + zy * R * G;
    for (int i = threadIdx.x; i < G * R; i += blockDim.x) {
        int x = i / G, c = i % G;
        const float v = line[c * (R + 1) + x];
        dst[i] = v;
        dst16[i] = __float2half_rn(v);
    }
}

// ---- fingerprint of the raw parameters (EMA copy_to / restore write through .data without bumping
// torch's version counter, so the host cannot see that the packed copy went stale) ------------------
struct FpSegments {
    const float* ptr[40];
    unsigned long long count[40];
    int n;
};

__global__ void fingerprint_kernel(FpSegments seg, unsigned long long* __restrict__ out) {
    unsigned long long h0 = 0, h1 = 0;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long base = 0;
    for (int s = 0; s < seg.n; ++s) {
        const unsigned int* p = reinterpret_cast<const unsigned int*>(seg.ptr[s]);
        for (unsigned long long i = tid; i < seg.count[s]; i += stride) {
            const unsigned long long x = p[i], pos = base + i;
            // position-weighted sums modulo 2^64 (commutative: atomics keep them deterministic)
            h0 += (x + 0x9E3779B97F4A7C15ull) * (2 * pos + 1);
            unsigned long long y = x * 0xBF58476D1CE4E5B9ull + pos;
            y ^= y >> 29;
            h1 += y * 0x94D049BB133111EBull;
        }
        base += seg.count[s];
    }
    for (int off = 16; off > 0; off >>= 1) {
        h0 += __shfl_xor_sync(0xffffffffu, h0, off);
        h1 += __shfl_xor_sync(0xffffffffu, h1, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, h0);
        atomicAdd(out + 1, h1);
    }
}

}  // namespace

int field_fingerprint(const FnLayout& L, const fenerf_field_params* p, unsigned long long* out, cudaStream_t st) {
    FpSegments seg;
    seg.n = 0;
    auto add = [&](const float* ptr, unsigned long long count) {
        if (ptr && count) { seg.ptr[seg.n] = ptr; seg.count[seg.n] = count; ++seg.n; }
    };
    const int n_trunk = L.trunk_hidden + 1, n_color = L.n_hidden - L.trunk_hidden;
    for (int i = 0; i < n_trunk; ++i) { add(p->trunk_w[i], (i == 0 ? 3 : FN_H) * FN_H); add(p->trunk_b[i], FN_H); }
    add(p->sigma_w, FN_H); add(p->sigma_b, 1);
    for (int i = 0; i < n_color; ++i) { add(p->color_w[i], (unsigned long long)(i == 0 ? L.kx + FN_H : FN_H) * FN_H); add(p->color_b[i], FN_H); }
    add(p->rgb_w, 3 * FN_H); add(p->rgb_b, 3);
    if (L.label_dim > 0) {
        add(p->label_w[0], FN_H * FN_H); add(p->label_b[0], FN_H);
        add(p->label_w[1], FN_H * FN_H); add(p->label_b[1], FN_H);
        add(p->label_w[2], (unsigned long long)L.label_dim * FN_H); add(p->label_b[2], L.label_dim);
    }
    if (L.grid_channels > 0) add(p->grid, (unsigned long long)L.grid_channels * L.grid_res * L.grid_res * L.grid_res);
    FN_CUDA_OK(cudaMemsetAsync(out, 0, 16, st));
    fingerprint_kernel<<<num_sms() * 4, 256, 0, st>>>(seg, out);
    FN_LAUNCH_OK("fingerprint_kernel");
    return 0;
}

int pack_field(const fenerf_field_desc* f, const FnLayout& L, const fenerf_field_params* p, void* packed_v,
               cudaStream_t st) {
    unsigned char* packed = static_cast<unsigned char*>(packed_v);
    FN_REQUIRE(p->trunk_w[0] && p->trunk_b[0], "trunk_w[0]/trunk_b[0] missing");
    pack_first_kernel<<<1, 256, 0, st>>>(p->trunk_w[0], p->trunk_b[0], (float*)(packed + L.first_w),
                                         (float*)(packed + L.first_b), packed + L.first_img);
    FN_LAUNCH_OK("pack_first_kernel");
    for (int l = 0; l < L.n_hidden; ++l) {
        bool is_c0 = (l == L.trunk_hidden);
        const float* w = l < L.trunk_hidden ? p->trunk_w[l + 1] : p->color_w[l - L.trunk_hidden];
        const float* b = l < L.trunk_hidden ? p->trunk_b[l + 1] : p->color_b[l - L.trunk_hidden];
        FN_REQUIRE(w && b, "weight/bias of hidden layer %d missing", l);
        int in_dim = is_c0 ? L.kx + FN_H : FN_H;
        int x_off = is_c0 ? L.kx : 0;
        int kx = is_c0 ? L.kx : 0, kx_pad = is_c0 ? L.kx_pad : 0;
        dim3 grid((FN_H + kx_pad + 31) / 32, FN_H / 32), block(32, 8);
        pack_hidden_kernel<<<grid, block, 0, st>>>(w, b, in_dim, x_off, kx, kx_pad, (float*)(packed + L.hid_w32[l]),
                                                   (float*)(packed + L.hid_b[l]), packed + L.hid_img[l]);
        FN_LAUNCH_OK("pack_hidden_kernel");
        if (is_c0) {
            pack_color0_ximg_kernel<<<1, 256, 0, st>>>(w, in_dim, L.grid_channels, packed + L.color0_ximg);
            FN_LAUNCH_OK("pack_color0_ximg_kernel");
        }
    }
    FN_REQUIRE(p->sigma_w && p->sigma_b && p->rgb_w && p->rgb_b, "sigma/rgb head parameters missing");
    pack_heads_kernel<<<3, 256, 0, st>>>(p->sigma_w, p->sigma_b, p->rgb_w, p->rgb_b, (float*)(packed + L.sigma_w),
                                         (float*)(packed + L.rgb_w));
    FN_LAUNCH_OK("pack_heads_kernel");
    if (L.label_dim > 0) {
        for (int i = 0; i < 3; i += 2) FN_REQUIRE(p->label_w[i] && p->label_b[i], "label layer %d missing", i);
        FN_REQUIRE((p->label_w[1] == nullptr) == (p->label_b[1] == nullptr), "label layer 1: weight and bias must both be given or both be NULL");
        double* u = reinterpret_cast<double*>(packed + L.label_scratch);
        label_step1_kernel<<<L.label_dim, FN_H, 0, st>>>(p->label_w[2], p->label_b[2], p->label_w[1], p->label_b[1], u);
        FN_LAUNCH_OK("label_step1_kernel");
        label_step2_kernel<<<L.label_dim, FN_H, 0, st>>>(u, p->label_w[0], p->label_b[0], L.label_dim,
                                                         (float*)(packed + L.label_w));
        FN_LAUNCH_OK("label_step2_kernel");
        label_step3_kernel<<<1, 256, 0, st>>>((float*)(packed + L.label_w), L.label_dim);
        FN_LAUNCH_OK("label_step3_kernel");
    }
    pack_head_imgs_kernel<<<1, 256, 0, st>>>((const float*)(packed + L.label_w), L.label_dim, p->sigma_w, p->rgb_w,
                                             packed + L.head_img, packed + L.rgb_img);
    FN_LAUNCH_OK("pack_head_imgs_kernel");
    if (L.grid_channels > 0) {
        FN_REQUIRE(p->grid, "grid missing");
        int R = L.grid_res, G = L.grid_channels;
        size_t smem = (size_t)G * (R + 1) * sizeof(float);
        FN_REQUIRE(smem <= 96 * 1024, "grid_res %d too large for the transpose tile", R);
        if (smem > 48 * 1024)
            FN_CUDA_OK(cudaFuncSetAttribute(pack_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pack_grid_kernel<<<R * R, 256, smem, st>>>(p->grid, (float*)(packed + L.grid), (__half*)(packed + L.grid16), R, G);
        FN_LAUNCH_OK("pack_grid_kernel");
    }
    (void)f;
    return 0;
}

}  // namespace fn
