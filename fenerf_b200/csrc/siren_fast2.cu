// Point network, FAST mode, second-generation kernel ("half-layer pipeline").
//
// Same math, orientation and operand layouts as siren_fast.cu (read its header first); what changes
// is the schedule.  One persistent CTA per SM owns ONE 128-point tile and overlaps the tensor pipe
// and the MUFU pipe on that single tile by splitting every FiLM layer into its two feature halves:
//
//     MMA(l, h0)  ->  epi(l, h0)  ||  MMA(l, h1)  ->  epi(l, h1)  ||  MMA(l+1, h0, k 0..127)  -> ...
//
// epi(l, h0) produces input features 0..127 of layer l+1, which is exactly what the first two
// k-chunks of MMA(l+1, h0) consume, so the next layer starts while epi(l, h1) is still running.
// Two things make this legal:
//   * six 16 KB activation chunk buffers used as a ring: layer l reads chunks (p .. p+3) mod 6,
//     epi(l,h0) writes (p+4, p+5) mod 6 -- free since layer l-1 -- and epi(l,h1) writes (p, p+1),
//     free once MMA(l,h1) has retired; then p += 4 (mod 6).  No buffer is ever overwritten while an
//     in-flight MMA can still read it.
//   * the accumulator halves are separate TMEM column ranges with their own "full" barriers, and
//     there are two accumulator SETS (stage parity): layer l+1 accumulates into one while the
//     epilogue drains the other, so the first two k-chunks of BOTH halves of layer l+1 run as soon
//     as epi(l,h0) is done and only the last two k-chunks wait for epi(l,h1).
// Activation chunks are stored MN-major (points contiguous): [k/8][point/64][k%8][64 points], 128B
// swizzle (tc5.cuh kDescHiMN).  The epilogue thread that owns feature k then writes its 64 points as
// eight 16-byte stores; the K-major layout of the first-generation kernel needs one 2-byte store per
// element, i.e. 2048 shared-store instructions per layer-tile, which the LSU (about one instruction
// per two cycles per SM) turns into the bottleneck (profiles/r01_trace_v12.txt: 1650 cycles per
// half-layer against a 1024-cycle MUFU bound).  The input chunk X stays K-major (one point per thread).
// With one tile per CTA the shared memory left over (227 KB - 96 KB - 16 KB) holds a 3 x 32 KB weight
// ring.  A slot carries TWO k-chunks of one feature half, i.e. 8 MMAs = 512 tensor cycles per
// full/empty barrier round trip: the issuer's wait + commit + bookkeeping costs ~400 cycles per
// round whatever the slot size (profiles/r01_trace_v9.txt, r01_mma_bench.txt: a tight issue loop
// holds 90 % of peak next to a busy epilogue), so 16 KB slots (256 cycles of work) starve the pipe.
//
//   warps 0..2    weight producers (one per ring slot)
//   warp  3       MMA issuer (warp-converged, elect.sync)
//   warps 4..11   epilogue: warp e serves TMEM lane quadrant e % 4 (features) and point half e / 4
#include "common.cuh"
#include "siren_common.cuh"
#include "tc5.cuh"

namespace fn {

namespace {

using namespace tc5;

constexpr int TILE = 128;
constexpr int NPROD = 3;
constexpr int RING = 3;
constexpr int MMA_WARP = NPROD;
constexpr int EPI_WARP0 = NPROD + 1;
constexpr int NEPI = 8;
constexpr int NTHREADS = (EPI_WARP0 + NEPI) * 32;    // 384
constexpr int NBUF = 6;                               // activation chunk buffers
constexpr uint32_t CHUNK_BYTES = 16384;
constexpr uint32_t STAGE_BYTES = 32768;      // one ring slot: two k-chunks of one feature half = 8 MMAs of work
constexpr uint32_t SMEM_A = 0;
constexpr uint32_t SMEM_X = NBUF * CHUNK_BYTES;
constexpr uint32_t SMEM_RING = SMEM_X + CHUNK_BYTES;
constexpr uint32_t SMEM_BAR = SMEM_RING + RING * STAGE_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 256;
constexpr int TMEM_COLS = 512;                       // two accumulator sets (stage parity) x two feature halves x 128 points
constexpr int MAX_LOADS = 128;
constexpr int MAX_STAGES = 16;

enum : uint8_t { EPI_FILM = 0, EPI_HEAD_TRUNK = 1, EPI_HEAD_RGB = 2 };

struct alignas(16) LoadOp {
    uint32_t src;          // byte offset in the packed buffer
    uint16_t bytes16;      // bytes / 16
    uint8_t x_chunk;       // first activation chunk (logical 0..3, mapped through the buffer ring) or 4 = input chunk
    uint8_t n_chunks;      // k-chunks in this load (heads: 4), else 1
    uint8_t k0, nk;        // K-steps inside a 64-wide chunk
    uint8_t n8;            // MMA N / 8
    uint8_t half;          // accumulator half (TMEM columns half * 128 ..)
    uint8_t first;         // 1: overwrite the accumulator with the first MMA of this load
    uint8_t w_is_a;        // 1: weights are the A operand (transposed FiLM layer); 0: B (head)
    uint8_t pad[2];
};
static_assert(sizeof(LoadOp) == 16, "LoadOp is read as one 16-byte constant-bank vector");

struct StageOp {
    uint8_t epi;           // EPI_*
    uint8_t film;          // FiLM layer index
    uint8_t n_loads;
    uint8_t split;         // loads [0, split) need only input features 0..127 (x_ready[0]); the rest all of them
    uint8_t acc0_end;      // loads [0, acc0_end) feed accumulator half 0
    uint8_t advance;       // 1: FiLM stage, the activation ring base advances by 4 afterwards
    uint8_t uniform;       // 1: plain 256x256 FiLM layer = exactly 4 loads (h0 k01, h0 k23, h1 k01, h1 k23)
    uint8_t pad;
};

struct Fast2Args {
    LoadOp loads[MAX_LOADS];
    StageOp stages[MAX_STAGES];
    int n_loads, n_stages;
    FnLayout L;
    const unsigned char* packed;
    const float* points;
    const float* dirs;
    const float* film;
    float* out;
    long long ppb, tiles_per_batch, n_tiles;
    int dir_group, lock_dirs;
    long long* trace;
};

__device__ __forceinline__ uint32_t buf_addr(uint32_t sbase, int p, int logical_chunk) {
    int b = p + logical_chunk;
    b = b >= NBUF ? b - NBUF : b;
    b = b >= NBUF ? b - NBUF : b;
    return sbase + SMEM_A + (uint32_t)b * CHUNK_BYTES;
}

template <bool kTrace>
__global__ void __launch_bounds__(NTHREADS, 1) siren_fast2_kernel(const __grid_constant__ Fast2Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + SMEM_BAR;             // [RING]
    const uint32_t bar_empty = bar_full + 8 * RING;         // [RING]
    const uint32_t bar_acc = bar_empty + 8 * RING;          // [2] accumulator half complete (MMA -> epilogue)
    const uint32_t bar_ready = bar_acc + 16;                // [2] next-layer features of half h written + acc half drained
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 8 * (2 * RING + 4));

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int h = 0; h < 2; ++h) { mbar_init(bar_acc + 8 * h, 1); mbar_init(bar_ready + 8 * h, NEPI); }
        fence_barrier_init();
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const FnLayout& L = a.L;

    if (warp < NPROD) {
        // ================= weight producers: load j of a stage -> slot j % RING = producer warp j % RING
        if (lane == 0) {
            uint32_t uses = 0;                       // completed uses of MY slot (phase parity)
            for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                int li = 0;
                for (int s = 0; s < a.n_stages; ++s) {
                    const int n = a.stages[s].n_loads;
                    for (int j = warp; j < n; j += RING, ++uses) {
                        mbar_wait(bar_empty + 8 * warp, (uses & 1) ^ 1);
                        const uint32_t bytes = (uint32_t)a.loads[li + j].bytes16 * 16u;
                        mbar_arrive_expect_tx(bar_full + 8 * warp, bytes);
                        bulk_g2s(sbase + SMEM_RING + warp * STAGE_BYTES, a.packed + a.loads[li + j].src, bytes, bar_full + 8 * warp);
                    }
                    li += n;
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer (warp-converged; an elected lane issues) =================
        uint32_t used[RING] = {0, 0, 0};            // per-slot use counts (phase parity), as in the producers
        uint32_t n_ready = 0;
        Tracer<kTrace> tr(lane == 0 ? a.trace : nullptr, 1);
        const uint32_t ring_lo = (sbase + SMEM_RING) >> 4;
        int tl = 0;
        for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tl) {
            int p = 2, li = 0;
            for (int s = 0; s < a.n_stages; ++s) {
                const StageOp sop = a.stages[s];
                mbar_wait(bar_ready, n_ready & 1);             // features 0..127 written, accumulator half 0 drained
                tc_fence_after();
                tr.log('A', tl, s, 0);
                if (sop.uniform) {
                    // straight-line: 4 loads x 8 MMAs in ring slots 0,1,2,0, order [h0 k01][h1 k01][h0 k23][h1 k23]:
                    // the first two need only input features 0..127, the last two all of them
                    uint32_t x_lo[4];
#pragma unroll
                    for (int kc = 0; kc < 4; ++kc) x_lo[kc] = buf_addr(sbase, p, kc) >> 4;
                    constexpr uint32_t idesc = umma_idesc_f16(TILE, 0, 1);          // B (activations) is MN-major
                    constexpr uint32_t kSlot16 = STAGE_BYTES >> 4, kChunk16 = CHUNK_BYTES >> 4;
                    const uint32_t d_set = tmem_base + (uint32_t)(s & 1) * 256u;
                    mbar_wait(bar_full, used[0] & 1);
                    tc_fence_after();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int slot = jj % RING;
                        const int half = jj & 1, kp = jj >> 1;  // feature half, k-chunk pair
                        if (jj == 2) {                          // k-chunks 2,3 need features 128..255
                            mbar_wait(bar_ready + 8, n_ready & 1);
                            tc_fence_after();
                            tr.log('B', tl, s, jj);
                        }
                        tr.log('F', tl, s, jj);
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (c == 1 && k == 0 && jj < 3) {   // next slot's wait overlaps this slot's MMAs
                                    const int ns = (jj + 1) % RING;
                                    mbar_wait(bar_full + 8 * ns, (used[ns] + (jj + 1 >= RING ? 1 : 0)) & 1);
                                    tc_fence_after();
                                }
                                tc_mma_f16_elect(d_set + half * 128, kDescHi | (uint64_t)(ring_lo + slot * kSlot16 + c * kChunk16 + 2 * k),
                                                 kDescHiMN | (uint64_t)(x_lo[kp * 2 + c] + 256 * k), idesc,
                                                 (kp == 0 && c == 0 && k == 0) ? 0u : 1u);
                            }
                        tc_commit_elect(bar_empty + 8 * slot);
                        if (jj == 2) { tc_commit_elect(bar_acc); tr.log('C', tl, s, 0); }
                    }
                    used[0] += 2; used[1] += 1; used[2] += 1;
                } else {
                    for (int j = 0; j < sop.n_loads; ++j) {
                        if (j == sop.split) {                  // the rest needs features 128..255 / half 1 drained
                            mbar_wait(bar_ready + 8, n_ready & 1);
                            tc_fence_after();
                            tr.log('B', tl, s, j);
                        }
                        const LoadOp op = a.loads[li + j];
                        const uint32_t slot = (uint32_t)j % RING;
                        const uint32_t cnt = slot == 0 ? used[0] : slot == 1 ? used[1] : used[2];
                        mbar_wait(bar_full + 8 * slot, cnt & 1);
                        tc_fence_after();
                        if (slot == 0) ++used[0]; else if (slot == 1) ++used[1]; else ++used[2];
                        tr.log('F', tl, s, j);
                        // activation chunks are MN-major, the input chunk X and all weight images K-major
                        const bool x_mn = op.x_chunk != 4;
                        const uint32_t idesc = umma_idesc_f16((uint32_t)op.n8 * 8u, (!op.w_is_a && x_mn) ? 1u : 0u, (op.w_is_a && x_mn) ? 1u : 0u);
                        const uint32_t w_stride = ((uint32_t)op.bytes16 * 16u) / op.n_chunks;
                        const uint32_t d_col = tmem_base + (uint32_t)(s & 1) * 256u + (uint32_t)op.half * 128u;
                        const uint64_t x_hi = x_mn ? kDescHiMN : kDescHi;
                        const uint32_t x_step = x_mn ? 256u : 2u;                    // per K-step, in 16 B units
#pragma unroll 1
                        for (int c = 0; c < op.n_chunks; ++c) {
                            const uint32_t x_lo = (op.x_chunk == 4 ? sbase + SMEM_X : buf_addr(sbase, p, op.x_chunk + c)) >> 4;
                            const uint32_t w_lo = (sbase + SMEM_RING + slot * STAGE_BYTES + c * w_stride) >> 4;
#pragma unroll 4
                            for (int k = 0; k < op.nk; ++k) {
                                const uint64_t xd = x_hi | (uint64_t)(x_lo + (uint32_t)(op.k0 + k) * x_step);
                                const uint64_t wd = kDescHi | (uint64_t)(w_lo + (uint32_t)(op.k0 + k) * 2u);
                                tc_mma_f16_elect(d_col, op.w_is_a ? wd : xd, op.w_is_a ? xd : wd, idesc,
                                                 (op.first && c == 0 && k == 0) ? 0u : 1u);
                            }
                        }
                        tc_commit_elect(bar_empty + 8 * slot);
                        if (j + 1 == sop.acc0_end) { tc_commit_elect(bar_acc); tr.log('C', tl, s, 0); }
                    }
                    if (sop.split >= sop.n_loads) {                // stage with no second part still consumes the phase
                        mbar_wait(bar_ready + 8, n_ready & 1);
                        tc_fence_after();
                    }
                }
                tc_commit_elect(bar_acc + 8);
                tr.log('C', tl, s, 1);
                ++n_ready;
                li += sop.n_loads;
                if (sop.advance) { p += 4; p = p >= NBUF ? p - NBUF : p; }
            }
        }
    } else {
        // ================= epilogue warps =================
        const int e = warp - EPI_WARP0;
        const int q = warp & 3;                        // TMEM lane quadrant this warp may access (hardware: warp id % 4)
        const int ch = e >> 2;                         // which 64-point half of the tile's columns this warp handles
        const int row = q * 32 + lane;                 // feature within a half (FiLM) / point (heads, input chunk)
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int C = L.out_dim;
        const float* sigma_w = reinterpret_cast<const float*>(a.packed + L.sigma_w);
        const float* rgb_w = reinterpret_cast<const float*>(a.packed + L.rgb_w);
        const float* label_w = reinterpret_cast<const float*>(a.packed + L.label_w);
        uint32_t n_acc = 0;
        Tracer<kTrace> tr(e == 0 && lane == 0 ? a.trace : nullptr, 2);
        int tl = 0;
        for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tl) {
            const long long b = tile / a.tiles_per_batch;
            const long long pnt = (tile % a.tiles_per_batch) * TILE + row;
            const bool valid = pnt < a.ppb;
            const long long flat = b * a.ppb + pnt;
            tr.log('T', tl, 0, 0);
            if (ch == 0) {
                // ---- build the input chunk: one point per thread (see layout.h for the slot order) ----
                float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) pos[i] = __fmul_rn(a.points[flat * 3 + i], L.input_scale);
                    if (a.lock_dirs) dir[2] = -1.f;
                    else {
                        const long long di = b * (a.ppb / a.dir_group) + pnt / a.dir_group;
#pragma unroll
                        for (int i = 0; i < 3; ++i) dir[i] = a.dirs[di * 3 + i];
                    }
                }
                __align__(16) __half slots[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) slots[i] = __float2half_rn(0.f);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    __half hi, lo;
                    split_f16(pos[i], hi, lo);
                    slots[FN_SLOT_POS + i] = hi; slots[FN_SLOT_POS + 3 + i] = lo; slots[FN_SLOT_POS + 6 + i] = hi;
                    split_f16(dir[i], hi, lo);
                    slots[FN_SLOT_DIR + i] = hi; slots[FN_SLOT_DIR + 3 + i] = lo; slots[FN_SLOT_DIR + 6 + i] = hi;
                }
                if (L.grid_channels > 0 && valid) {
                    float feat[32];
                    grid_features32(reinterpret_cast<const float*>(a.packed + L.grid), L.grid_res, pos[0], pos[1], pos[2], feat);
#pragma unroll
                    for (int i = 0; i < 32; ++i) slots[FN_SLOT_FEAT + i] = __float2half_rn(feat[i]);
                }
                const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
                const uint4* src = reinterpret_cast<const uint4*>(slots);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(smem + SMEM_X + row_off + (((uint32_t)j ^ (uint32_t)(row & 7)) << 4)) = src[j];
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_ready); mbar_arrive(bar_ready + 8); }

            int p = 2;
            for (int s = 0; s < a.n_stages; ++s) {
                const StageOp sop = a.stages[s];
                const bool last = (s + 1 == a.n_stages);
                const uint32_t t_set = t_lane + (uint32_t)(s & 1) * 256u;     // this stage's accumulator set
                if (sop.epi == EPI_FILM) {
                    const int fl = row;
                    const float* film_l = a.film + ((size_t)b * L.n_film + sop.film) * 2 * FN_H;
                    const float* bias = reinterpret_cast<const float*>(
                        a.packed + (sop.film == 0 ? L.first_b : L.hid_b[sop.film - 1]));
                    float fr[2], ph[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        fr[h] = __ldg(film_l + h * 128 + fl);
                        ph[h] = fmaf(fr[h], __ldg(bias + h * 128 + fl), __ldg(film_l + FN_H + h * 128 + fl));
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(bar_acc + 8 * h, n_acc & 1);
                        tc_fence_after();
                        tr.log('W', tl, s, h);
                        // output feature f = h*128 + fl of the next layer -> logical chunk f/64 of the NEXT input:
                        // buffers (p+4, p+5) for h = 0, (p+6, p+7) = (p, p+1) for h = 1   (all mod NBUF).
                        // MN-major chunk: [k/8][point/64][k%8][64 points]; this warp's 64 points are one 128-byte row.
                        const uint32_t kk = (uint32_t)(fl & 63);
                        unsigned char* rowp = smem + (buf_addr(sbase, p, 4 + h * 2 + (fl >> 6)) - sbase) + (kk >> 3) * 2048u +
                                              (uint32_t)ch * 1024u + (kk & 7u) * 128u;
                        const uint32_t sw = kk & 7u;
                        const float f_h = fr[h], p_h = ph[h];
                        // TMEM -> registers runs at ~64 B/clk/SM, i.e. as long as the MUFU work itself: keep
                        // the two overlapped with small (16-column) double-buffered loads
                        uint32_t r[2][16];
                        tc_ld16(t_set + h * 128 + ch * 64, r[0]);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {          // this warp's 64 points: 4 groups of 16
                            tc_wait_ld();
                            if (g < 3) tc_ld16(t_set + h * 128 + ch * 64 + (g + 1) * 16, r[(g + 1) & 1]);
#pragma unroll
                            for (int j8 = 0; j8 < 2; ++j8) {   // 8 points -> one 16-byte piece
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = __sinf(fmaf(f_h, __uint_as_float(r[g & 1][j8 * 8 + j]), p_h));
                                uint4 pk;
                                pk.x = pack_half2(v[0], v[1]); pk.y = pack_half2(v[2], v[3]);
                                pk.z = pack_half2(v[4], v[5]); pk.w = pack_half2(v[6], v[7]);
                                const uint32_t piece = (uint32_t)(g * 2 + j8);
                                *reinterpret_cast<uint4*>(rowp + ((piece ^ sw) << 4)) = pk;
                            }
                        }
                        tr.log('D', tl, s, h);
                        if (!last) {
                            fence_async_smem();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_ready + 8 * h);
                        }
                    }
                    ++n_acc;
                    p += 4; p = p >= NBUF ? p - NBUF : p;
                } else {
                    mbar_wait(bar_acc, n_acc & 1);
                    tc_fence_after();
                    tr.log('W', tl, s, 0);
                    if (ch == 0) {                 // lanes = points: the four ch == 0 warps cover the 128 points
                        if (sop.epi == EPI_HEAD_TRUNK) {
                            if (L.label_dim > 0) {
                                uint32_t r[32];
                                tc_ld32(t_set, r);
                                tc_wait_ld();
                                if (valid) {
                                    const float inv_scale = __ldg(label_w + FENERF_MAX_LABEL * FN_H + FENERF_MAX_LABEL);
#pragma unroll
                                    for (int o = 0; o < 32; ++o) {
                                        if (o < L.label_dim)
                                            a.out[flat * C + o] = fmaf(__uint_as_float(r[o]), inv_scale, __ldg(label_w + FENERF_MAX_LABEL * FN_H + o));
                                        else if (o == L.label_dim)
                                            a.out[flat * C + (C - 1)] = __uint_as_float(r[o]) + __ldg(sigma_w + FN_H);
                                    }
                                }
                            } else {
                                uint32_t r[8];
                                tc_ld8(t_set, r);
                                tc_wait_ld();
                                if (valid) a.out[flat * C + (C - 1)] = __uint_as_float(r[0]) + __ldg(sigma_w + FN_H);
                            }
                        } else {
                            uint32_t r[8];
                            tc_ld8(t_set, r);
                            tc_wait_ld();
                            if (valid) {
#pragma unroll
                                for (int o = 0; o < 3; ++o) {
                                    const float x = __uint_as_float(r[o]) + __ldg(rgb_w + 3 * FN_H + o);
                                    a.out[flat * C + L.label_dim + o] = __fdividef(1.f, 1.f + __expf(-x));
                                }
                            }
                        }
                    }
                    mbar_wait(bar_acc + 8, n_acc & 1);
                    ++n_acc;
                    tr.log('D', tl, s, 0);
                    if (!last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { mbar_arrive(bar_ready); mbar_arrive(bar_ready + 8); }
                    }
                }
            }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---- host: the per-tile stage / load program ------------------------------------------------------
LoadOp* push(Fast2Args& A) { LoadOp* op = &A.loads[A.n_loads++]; memset(op, 0, sizeof(*op)); op->n_chunks = 1; op->w_is_a = 1; return op; }

void push_film_pair(Fast2Args& A, size_t img_off, int half, int pair, bool first) {
    // image order [half][kc][128 rows][64 k]: one 32 KB load = k-chunks 2*pair, 2*pair+1 of a feature half
    LoadOp* op = push(A);
    op->src = (uint32_t)(img_off + (size_t)half * 65536 + (size_t)pair * STAGE_BYTES);
    op->bytes16 = STAGE_BYTES / 16; op->x_chunk = (uint8_t)(pair * 2); op->n_chunks = 2; op->k0 = 0; op->nk = 4; op->n8 = TILE / 8;
    op->half = (uint8_t)half; op->first = first ? 1 : 0;
}

void push_input_half(Fast2Args& A, size_t img_off, int half, int k0, int nk, bool first) {
    LoadOp* op = push(A);
    op->src = (uint32_t)(img_off + (size_t)half * 16384);
    op->bytes16 = 16384 / 16; op->x_chunk = 4; op->k0 = (uint8_t)k0; op->nk = (uint8_t)nk; op->n8 = TILE / 8;
    op->half = (uint8_t)half; op->first = first ? 1 : 0;
}

void push_head(Fast2Args& A, size_t img_off, int img_rows, int n) {
    LoadOp* op = push(A);
    op->src = (uint32_t)img_off;
    op->bytes16 = (uint16_t)((4 * img_rows * FN_KCHUNK * 2) / 16); op->x_chunk = 0; op->n_chunks = 4; op->k0 = 0; op->nk = 4;
    op->n8 = (uint8_t)(n / 8); op->half = 0; op->first = 1; op->w_is_a = 0;
}

bool build_program(const FnLayout& L, Fast2Args& A) {
    A.n_loads = 0; A.n_stages = 0;
    auto end_stage = [&](uint8_t epi, uint8_t film, int l0, int split, int acc0_end, int advance) {
        StageOp& st = A.stages[A.n_stages++];
        memset(&st, 0, sizeof(st));
        st.epi = epi; st.film = film; st.n_loads = (uint8_t)(A.n_loads - l0); st.split = (uint8_t)split;
        st.acc0_end = (uint8_t)acc0_end; st.advance = (uint8_t)advance; st.uniform = 0;
    };
    {   // first layer: K-step 0 of the input chunk; needs nothing from a previous layer but the input chunk itself
        int l0 = A.n_loads;
        push_input_half(A, L.first_img, 0, 0, 1, true);
        push_input_half(A, L.first_img, 1, 0, 1, true);
        end_stage(EPI_FILM, 0, l0, 0, 1, 1);
    }
    for (int l = 0; l < L.n_hidden; ++l) {
        if (l == L.trunk_hidden) {
            int l0 = A.n_loads;
            push_head(A, L.head_img, 32, L.label_dim > 0 ? 32 : 8);
            end_stage(EPI_HEAD_TRUNK, 0, l0, 0, 1, 0);
        }
        const bool c0 = (l == L.trunk_hidden);
        const int nkx = L.grid_channels > 0 ? 3 : 1;
        int l0 = A.n_loads;
        // order [h0 k01][h1 k01] | [h0 k23][+h0 x][h1 k23][+h1 x]: the first two need input features 0..127 only
        push_film_pair(A, L.hid_img[l], 0, 0, true);
        push_film_pair(A, L.hid_img[l], 1, 0, true);
        push_film_pair(A, L.hid_img[l], 0, 1, false);
        if (c0) push_input_half(A, L.color0_ximg, 0, 1, nkx, false);
        int acc0_end = A.n_loads - l0;
        push_film_pair(A, L.hid_img[l], 1, 1, false);
        if (c0) push_input_half(A, L.color0_ximg, 1, 1, nkx, false);
        end_stage(EPI_FILM, (uint8_t)(l + 1), l0, 2, acc0_end, 1);
        A.stages[A.n_stages - 1].uniform = c0 ? 0 : 1;
        if (A.n_loads > MAX_LOADS - 16 || A.n_stages > MAX_STAGES - 3) return false;
    }
    {
        int l0 = A.n_loads;
        push_head(A, L.rgb_img, 8, 8);
        end_stage(EPI_HEAD_RGB, 0, l0, 0, 1, 0);
    }
    return true;
}

}  // namespace

int siren_points_fast2(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                       const float* film, int batch, long long ppb, int dir_group, int lock_dirs, float* out,
                       long long* trace, cudaStream_t st) {
    static_assert(sizeof(Fast2Args) <= 4000, "kernel parameter block too large");
    static_assert(SMEM_TOTAL <= 232448, "one CTA per SM: 227 KB of shared memory");
    FN_REQUIRE(L.trunk_hidden >= 1 && L.n_hidden - L.trunk_hidden >= 1, "field needs >= 2 trunk and >= 1 colour layers");
    FN_REQUIRE(L.label_dim < 32, "the tcgen05 path packs labels and sigma into one 32-row head (label_dim <= 31)");
    Fast2Args a;
    memset(&a, 0, sizeof(a));
    FN_REQUIRE(build_program(L, a), "field too deep for the stage program");
    a.L = L; a.packed = packed; a.points = points; a.dirs = dirs; a.film = film; a.out = out;
    a.ppb = ppb; a.tiles_per_batch = (ppb + TILE - 1) / TILE; a.n_tiles = a.tiles_per_batch * batch;
    a.dir_group = dir_group < 1 ? 1 : dir_group; a.lock_dirs = lock_dirs; a.trace = trace;
    if (a.n_tiles <= 0) return 0;
    FN_REQUIRE(ppb % a.dir_group == 0, "points_per_batch %lld not a multiple of dir_group %d", ppb, a.dir_group);
    auto kernel = a.trace ? siren_fast2_kernel<true> : siren_fast2_kernel<false>;
    FN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    int blocks = (int)(a.n_tiles < (long long)num_sms() ? a.n_tiles : (long long)num_sms());
    kernel<<<blocks, NTHREADS, SMEM_TOTAL, st>>>(a);
    FN_LAUNCH_OK("siren_fast2_kernel");
    return 0;
}

}  // namespace fn
