// Point network, FAST mode: the FiLM-SIREN stack on the 5th-generation tensor cores.
//
// Replaces <SIREN>.forward_with_frequencies_phase_shifts (siren/siren.py:164-178, 1509-1530).
// One persistent CTA per SM walks PAIRS of 128-point tiles; per tile every FiLM layer is
//     tcgen05.mma (fp16 operands, fp32 accumulator in TMEM)  ->  epilogue warps:
//     tcgen05.ld, sin(freq * (acc + b) + phase), fp16, written back as the next layer's operand
// so activations never leave the SM.  Weights are pre-swizzled UMMA images in HBM/L2 (pack.cu) and
// stream through a shared-memory ring with 1-D bulk copies (cp.async.bulk, the TMA engine).
//
//   warps 0..3   weight producers    warp w owns ring slot w: cp.async.bulk global -> slot, arms full[w]
//   warp 4       MMA issuer          one elected lane issues tcgen05.mma / tcgen05.commit; owns TMEM
//   warps 5..8   epilogue of tile X  128 threads
//   warps 9..12  epilogue of tile Y  128 threads
// (One thread can start only one bulk copy per ~680 cycles whatever its size -- 24 B/clk for 16 KB
// copies -- while four issuing warps reach ~90 B/clk/SM: tools/bulk_bench.cu, profiles/r01_bulk_bench.txt.
// Hence one producer warp per ring slot.)
//
// Two tiles, one ring.  A bulk copy from L2 has a ~1600-cycle turnaround per ring slot regardless of
// its size (profiles/r01_trace_v2*.txt), so the weight stream is latency bound and what counts is
// bytes in flight.  The two tiles of a CTA run the same layer program half a layer apart: while X's
// epilogue warps turn its accumulator into the next activations, the MMA issuer streams Y's layer
// through the WHOLE 4-slot ring, and vice versa -- each tile sees a 64 KB ring during its MMA phase
// and the tensor pipe, the MUFU pipe and the copy engine overlap inside one CTA.
//
// Orientation.  The FiLM layers run TRANSPOSED: D^T[feature][point] = W[feature][k] . X[point][k],
// i.e. the weight image is the MMA's A operand (M = 128 features per half) and the activation tile
// its B operand (N = 128 points).  An epilogue thread then owns one TMEM lane = one output feature:
// its FiLM constants (freq, freq*bias + phase) sit in two registers, each of its 2 x 128 accumulator
// values costs FFMA + FMUL + MUFU.SIN + F2FP + a 2-byte shared store, and nothing is re-loaded per
// column (the first version of this kernel had points on lanes and spent ~5x the MUFU-bound time
// waiting on per-column FiLM loads; profiles/r01_trace_v1.txt).  The activation tile is K-major
// [point][k] in both roles, so the small heads run in the other orientation on the same buffer:
// D[point][head] with the tile as A (M = 128 points) and an 8- or 32-row head image as B.
//
// The 3-wide inputs (position for the first layer, view direction for the colour layer) go through
// the tensor cores as well, split hi/lo in fp16 on both operands (hi*hi + lo*hi + hi*lo) so they
// keep fp32-level accuracy; grid features ride in the same 64-wide "input chunk".
//
// Tensor-pipe bound by design (2*256*256 FLOP per point per layer); co-limited by MUFU (one sin per
// output element, 16/clk/SM) and by the L2->SM weight stream (128 KB per layer per 128-point tile).
//
// Shared memory (224 KB, one CTA per SM; TMEM 2 x 256 columns):
//   A     2 x 4 x 16 KB  activations [128 points][64 k] f16 x 4 k-chunks per tile, 128B swizzle, K-major
//   X     2 x 16 KB      input chunk [128 points][64 slots] per tile   (layout.h: slot order)
//   ring  4 x 16 KB      weight stages [128 feature rows][64 k] (half of one k-chunk image)
#include "common.cuh"
#include "siren_common.cuh"
#include "tc5.cuh"

namespace fn {

namespace {

constexpr int TILE = 128;
constexpr int NTHREADS = 416;
constexpr int MMA_WARP = 4, EPI_WARP0 = 5;
constexpr int RING = 4;
constexpr uint32_t STAGE_BYTES = 16384;
constexpr uint32_t A_CHUNK_BYTES = 16384;
constexpr uint32_t TILE_SMEM = 5 * A_CHUNK_BYTES;               // 4 activation chunks + the input chunk
constexpr uint32_t SMEM_A = 0;                                  // + t * TILE_SMEM
constexpr uint32_t SMEM_X = 4 * A_CHUNK_BYTES;                  // + t * TILE_SMEM
constexpr uint32_t SMEM_RING = 2 * TILE_SMEM;
constexpr uint32_t SMEM_BAR = SMEM_RING + RING * STAGE_BYTES;   // 229376
constexpr uint32_t SMEM_TOTAL = SMEM_BAR + 128;
constexpr int TMEM_COLS = 512;                                  // 256 accumulator columns per tile
constexpr int MAX_LOADS = 128;
constexpr int MAX_STAGES = 16;

enum : uint8_t { EPI_FILM = 0, EPI_HEAD_TRUNK = 1, EPI_HEAD_RGB = 2 };

struct LoadOp {            // one ring stage: a bulk copy and the MMAs that consume it
    uint32_t src;          // byte offset in the packed buffer
    uint16_t bytes;        // multiple of 16, <= STAGE_BYTES (stored / 16)
    uint8_t a_chunk;       // activation chunk 0..3, 4 = input chunk
    uint8_t k0, nk;        // K-steps (16 wide) inside the 64-wide chunk
    uint8_t n8;            // MMA N / 8
    uint16_t d_col;        // accumulator column offset
    uint8_t first;         // 1: first MMA of this accumulator range (overwrite instead of accumulate)
    uint8_t last;          // 1: last load of its stage -> commit the accumulator
    uint8_t w_is_a;        // 1: ring stage is the A operand (transposed FiLM layer); 0: it is B (head)
    uint8_t n_chunks;      // k-chunks packed in this one load (heads: 4, each `bytes / 4` apart), else 1
};

struct StageOp {
    uint8_t epi;           // EPI_*
    uint8_t film;          // FiLM layer index
    uint8_t n_loads;
    uint8_t uniform;       // 1: the first 8 loads are the canonical 4 k-chunks x 2 halves of a 256x256 image
};

struct FastArgs {
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
    long long* trace;   // diagnostics: per-role clock64 log of CTA 0 (fenerf_debug_trace), or NULL
};

using namespace tc5;

// ---- the kernel -----------------------------------------------------------------------------
template <bool kTrace>
__global__ void __launch_bounds__(NTHREADS, 1) siren_fast_kernel(const __grid_constant__ FastArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + SMEM_BAR;            // RING x 8 B
    const uint32_t bar_empty = bar_full + 8 * RING;        // RING x 8 B
    const uint32_t bar_acc = bar_empty + 8 * RING;         // [2] accumulator ready (MMA -> epilogue of tile t)
    const uint32_t bar_aready = bar_acc + 16;              // [2] operand ready + TMEM drained (epilogue t -> MMA)
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 8 * (2 * RING + 4));

    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(bar_acc + 8 * t, 1);
            mbar_init(bar_aready + 8 * t, 4);      // one arrival per epilogue warp of the tile
        }
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

    if (warp < RING) {
        // ================= weight producers =================
        // job order (shared with the MMA issuer): for each pair, for each stage, tile X then tile Y;
        // load number `it` lives in ring slot it % RING and is issued by producer warp it % RING
        if (lane == 0) {
            // Ring slots are assigned per stage: load j of a (stage, tile) job lives in slot j % RING, so
            // the MMA issuer's unrolled code addresses the ring with immediates.  Each slot keeps its own
            // use count (phase parity), identical here and in the issuer.
            uint32_t uses = 0;                       // completed uses of MY slot
            Tracer<kTrace> tr(warp == 0 ? a.trace : nullptr, 0);
            int tl = 0;
            for (long long pair = blockIdx.x; pair * 2 < a.n_tiles; pair += gridDim.x, ++tl) {
                const int nt = (pair * 2 + 1 < a.n_tiles) ? 2 : 1;
                int li = 0;
                for (int s = 0; s < a.n_stages; ++s) {
                    const int n = a.stages[s].n_loads;
                    for (int t = 0; t < nt; ++t)
                        for (int j = warp; j < n; j += RING, ++uses) {
                            const int i = li + j;
                            mbar_wait(bar_empty + 8 * warp, (uses & 1) ^ 1);
                            tr.log('E', tl, s, t * 64 + j);
                            const uint32_t bytes = (uint32_t)a.loads[i].bytes * 16u;
                            mbar_arrive_expect_tx(bar_full + 8 * warp, bytes);
                            bulk_g2s(sbase + SMEM_RING + warp * STAGE_BYTES, a.packed + a.loads[i].src, bytes, bar_full + 8 * warp);
                            tr.log('B', tl, s, t * 64 + j);
                        }
                    li += n;
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        // the whole warp runs this loop converged; an elected lane issues (tc_*_elect)
        {
            uint32_t used[RING] = {0, 0, 0, 0};      // per-slot use counts (phase parity), see the producers
            uint32_t n_ready[2] = {0, 0};
            Tracer<kTrace> tr(lane == 0 ? a.trace : nullptr, 1);
            const uint32_t ring_lo = (sbase + SMEM_RING) >> 4;
            int tl = 0;
            for (long long pair = blockIdx.x; pair * 2 < a.n_tiles; pair += gridDim.x, ++tl) {
                const int nt = (pair * 2 + 1 < a.n_tiles) ? 2 : 1;
                int li0 = 0;
                for (int s = 0; s < a.n_stages; ++s) {
                    const StageOp sop = a.stages[s];
                    for (int t = 0; t < nt; ++t) {
                        mbar_wait(bar_aready + 8 * t, n_ready[t] & 1);     // inputs written, accumulator drained
                        ++n_ready[t];
                        tc_fence_after();
                        tr.log('A', tl, s, t);
                        int j0 = 0;
                        if (sop.uniform) {
                            // straight-line issue for a 256x256 FiLM layer: 4 k-chunks x 2 feature halves in
                            // ring slots 0,1,2,3,0,1,2,3; every descriptor word is a base plus an immediate.
                            // The wait for load jj+1 sits between the MMAs of load jj.
                            const uint32_t x_lo = (sbase + t * TILE_SMEM + SMEM_A) >> 4;
                            const uint32_t d0 = tmem_base + t * 256;
                            constexpr uint32_t idesc = umma_idesc_f16(TILE);
                            mbar_wait(bar_full, used[0] & 1);
                            tc_fence_after();
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                constexpr uint32_t kStage16 = STAGE_BYTES >> 4, kChunk16 = A_CHUNK_BYTES >> 4;
                                const int slot = jj & 3;
                                tr.log('F', tl, s, t * 64 + jj);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (k == 2 && jj < 7) {
                                        const int nslot = (jj + 1) & 3;
                                        mbar_wait(bar_full + 8 * nslot, (used[nslot] + (jj + 1 >= 4 ? 1 : 0)) & 1);
                                        tc_fence_after();
                                    }
                                    tc_mma_f16_elect(d0 + (jj & 1) * 128, kDescHi | (uint64_t)(ring_lo + slot * kStage16 + 2 * k),
                                                     kDescHi | (uint64_t)(x_lo + (jj >> 1) * kChunk16 + 2 * k), idesc,
                                                     (jj < 2 && k == 0) ? 0u : 1u);
                                }
                                tr.log('M', tl, s, t * 64 + jj);
                                tc_commit_elect(bar_empty + 8 * slot);
                            }
#pragma unroll
                            for (int q = 0; q < RING; ++q) used[q] += 2;
                            if (sop.n_loads == 8) { tc_commit_elect(bar_acc + 8 * t); tr.log('C', tl, s, t); }
                            j0 = 8;
                        }
                        for (int j = j0; j < sop.n_loads; ++j) {
                            const LoadOp op = a.loads[li0 + j];
                            const uint32_t slot = (uint32_t)j % RING;
                            uint32_t cnt = slot == 0 ? used[0] : slot == 1 ? used[1] : slot == 2 ? used[2] : used[3];
                            mbar_wait(bar_full + 8 * slot, cnt & 1);
                            tc_fence_after();
                            if (slot == 0) ++used[0]; else if (slot == 1) ++used[1]; else if (slot == 2) ++used[2]; else ++used[3];
                            tr.log('F', tl, s, t * 64 + j);
                            const uint32_t idesc = umma_idesc_f16((uint32_t)op.n8 * 8u);
                            const uint32_t w_stride = op.n_chunks == 4 ? ((uint32_t)op.bytes * 4u) : ((uint32_t)op.bytes * 16u);
                            for (int c = 0; c < op.n_chunks; ++c) {
                                const int xc = op.a_chunk + c;
                                const uint32_t x_addr = sbase + t * TILE_SMEM + (xc < 4 ? SMEM_A + xc * A_CHUNK_BYTES : SMEM_X);
                                const uint32_t w_addr = sbase + SMEM_RING + slot * STAGE_BYTES + c * w_stride;
                                const uint32_t a_addr = op.w_is_a ? w_addr : x_addr;
                                const uint32_t b_addr = op.w_is_a ? x_addr : w_addr;
                                for (int k = 0; k < op.nk; ++k) {
                                    const uint32_t koff = (uint32_t)(op.k0 + k) * 32u;   // 16 f16 = 32 B inside the swizzle row
                                    tc_mma_f16_elect(tmem_base + t * 256 + op.d_col, umma_desc_sw128(a_addr + koff),
                                                     umma_desc_sw128(b_addr + koff), idesc, (op.first && c == 0 && k == 0) ? 0u : 1u);
                                }
                            }
                            tr.log('M', tl, s, t * 64 + j);
                            tc_commit_elect(bar_empty + 8 * slot);      // ring stage reusable once these MMAs retire
                            if (op.last) { tc_commit_elect(bar_acc + 8 * t); tr.log('C', tl, s, t); }
                        }
                    }
                    li0 += sop.n_loads;
                }
            }
        }
    } else {
        // ================= epilogue warps =================
        const int t = (warp - EPI_WARP0) >> 2;        // which tile of the pair this warp serves
        const int q = warp & 3;                       // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;                // point slot in the tile / feature within a half
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
        unsigned char* tsm = smem + t * TILE_SMEM;    // this tile's activation + input chunks
        const uint32_t my_acc = bar_acc + 8 * t, my_aready = bar_aready + 8 * t;
        const uint32_t row_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        const uint32_t sw = (uint32_t)(row & 7);
        const int C = L.out_dim;
        const float* sigma_w = reinterpret_cast<const float*>(a.packed + L.sigma_w);
        const float* rgb_w = reinterpret_cast<const float*>(a.packed + L.rgb_w);
        const float* label_w = reinterpret_cast<const float*>(a.packed + L.label_w);
        uint32_t n_acc = 0;
        Tracer<kTrace> tr((warp == EPI_WARP0 || warp == EPI_WARP0 + 4) && lane == 0 ? a.trace : nullptr, 2 + t);
        int tl = 0;
        for (long long pair = blockIdx.x; pair * 2 + t < a.n_tiles; pair += gridDim.x, ++tl) {
            const long long tile = pair * 2 + t;
            tr.log('T', tl, 0, 0);
            // ---- build the input chunk for this tile ----
            const long long b = tile / a.tiles_per_batch;
            const long long p = (tile % a.tiles_per_batch) * TILE + row;
            const bool valid = p < a.ppb;
            const long long flat = b * a.ppb + p;
            {
                float pos[3] = {0.f, 0.f, 0.f}, dir[3] = {0.f, 0.f, 0.f};
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) pos[i] = __fmul_rn(a.points[flat * 3 + i], L.input_scale);
                    if (a.lock_dirs) dir[2] = -1.f;
                    else {
                        const long long di = b * (a.ppb / a.dir_group) + p / a.dir_group;
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
                const uint4* src = reinterpret_cast<const uint4*>(slots);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(tsm + SMEM_X + row_off + (((uint32_t)j ^ sw) << 4)) = src[j];
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(my_aready);

            for (int s = 0; s < a.n_stages; ++s) {
                const StageOp sop = a.stages[s];
                if (sop.epi == EPI_FILM) {
                    // this thread's two output features: fl and 128 + fl (one per accumulator half);
                    // FiLM constants fetched before the accumulator wait so their latency is hidden
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
                    mbar_wait(my_acc, n_acc & 1);
                    ++n_acc;
                    tc_fence_after();
                    tr.log('W', tl, s, 0);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // element (point p, feature f = h*128 + fl) -> chunk f/64, k = f%64
                        const uint32_t kk = (uint32_t)(fl & 63);
                        unsigned char* chunk = tsm + SMEM_A + (uint32_t)(h * 2 + (fl >> 6)) * A_CHUNK_BYTES + (kk & 7u) * 2u;
                        const uint32_t kc = kk >> 3;
                        const float f_h = fr[h], p_h = ph[h];
                        uint32_t r[2][32];
                        tc_ld32(t_lane + h * 128, r[0]);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {          // 4 groups of 32 points
                            tc_wait_ld();
                            if (g + 1 < 4) tc_ld32(t_lane + h * 128 + (g + 1) * 32, r[(g + 1) & 1]);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const uint32_t p = (uint32_t)(g * 32 + j);
                                const float v = __sinf(fmaf(f_h, __uint_as_float(r[g & 1][j]), p_h));
                                const uint32_t off = (p >> 3) * 1024u + (p & 7u) * 128u + (((kc ^ (p & 7u)) & 7u) << 4);
                                *reinterpret_cast<__half*>(chunk + off) = __float2half_rn(v);
                            }
                        }
                    }
                } else {
                    mbar_wait(my_acc, n_acc & 1);
                    ++n_acc;
                    tc_fence_after();
                    tr.log('W', tl, s, 0);
                    if (sop.epi == EPI_HEAD_TRUNK) {
                        // columns: labels 0..L-1 (scaled), sigma at column L
                        if (L.label_dim > 0) {
                            uint32_t r[32];
                            tc_ld32(t_lane, r);
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
                            tc_ld8(t_lane, r);
                            tc_wait_ld();
                            if (valid) a.out[flat * C + (C - 1)] = __uint_as_float(r[0]) + __ldg(sigma_w + FN_H);
                        }
                    } else {
                        uint32_t r[8];
                        tc_ld8(t_lane, r);
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
                tr.log('D', tl, s, 0);
                if (s + 1 < a.n_stages) {
                    // next stage may overwrite the accumulator and read what was just written
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(my_aready);
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

// ---- host: the per-tile stage / load program --------------------------------------------------
struct Program {
    FastArgs args;
    bool ok;
};

void push_image_loads(FastArgs& A, size_t img_off, int n_chunks, bool first_of_stage) {
    // a [256 features][64 k] image per k-chunk, streamed as two 128-feature halves (A operand);
    // half h accumulates into TMEM columns [128 h, 128 h + 128) = the 128 points of the tile
    for (int kc = 0; kc < n_chunks; ++kc)
        for (int half = 0; half < 2; ++half) {
            LoadOp& op = A.loads[A.n_loads++];
            op.src = (uint32_t)(img_off + (size_t)half * 65536 + (size_t)kc * STAGE_BYTES);   // [half][kc][128 rows]
            op.bytes = STAGE_BYTES / 16;
            op.a_chunk = (uint8_t)kc;
            op.k0 = 0; op.nk = 4; op.n8 = TILE / 8; op.d_col = (uint16_t)(half * 128);
            op.first = (first_of_stage && kc == 0) ? 1 : 0;
            op.last = 0; op.w_is_a = 1; op.n_chunks = 1;
        }
}

void push_input_loads(FastArgs& A, size_t img_off, int k0, int nk, bool first) {
    for (int half = 0; half < 2; ++half) {
        LoadOp& op = A.loads[A.n_loads++];
        op.src = (uint32_t)(img_off + (size_t)half * STAGE_BYTES);
        op.bytes = STAGE_BYTES / 16; op.a_chunk = 4; op.k0 = (uint8_t)k0; op.nk = (uint8_t)nk; op.n8 = TILE / 8;
        op.d_col = (uint16_t)(half * 128); op.first = first ? 1 : 0; op.last = 0; op.w_is_a = 1; op.n_chunks = 1;
    }
}

void push_head_loads(FastArgs& A, size_t img_off, int img_rows, int n) {
    // the head image [4 k-chunks][img_rows][64 k] is the B operand (N = n <= img_rows), the activation
    // tile the A operand: D[point][head] in TMEM columns [0, n).  All four k-chunks travel in ONE
    // bulk copy: a slot's turnaround does not depend on its size.
    LoadOp& op = A.loads[A.n_loads++];
    op.src = (uint32_t)img_off;
    op.bytes = (uint16_t)((4 * img_rows * FN_KCHUNK * 2) / 16); op.a_chunk = 0; op.k0 = 0; op.nk = 4; op.n8 = (uint8_t)(n / 8);
    op.d_col = 0; op.first = 1; op.last = 0; op.w_is_a = 0; op.n_chunks = 4;
}

bool build_program(const FnLayout& L, FastArgs& A) {
    A.n_loads = 0; A.n_stages = 0;
    auto end_stage = [&](uint8_t epi, uint8_t film, int first_load) {
        StageOp& st = A.stages[A.n_stages++];
        st.epi = epi; st.film = film; st.n_loads = (uint8_t)(A.n_loads - first_load); st.uniform = 0;
        A.loads[A.n_loads - 1].last = 1;
    };
    // first layer: only K-step 0 of the input chunk (position hi/lo slots)
    {
        int l0 = A.n_loads;
        push_input_loads(A, L.first_img, 0, 1, true);
        end_stage(EPI_FILM, 0, l0);
    }
    for (int l = 0; l < L.n_hidden; ++l) {
        if (l == L.trunk_hidden) {
            // heads on the trunk output: labels (if any) + sigma
            int l0 = A.n_loads;
            push_head_loads(A, L.head_img, 32, L.label_dim > 0 ? 32 : 8);
            end_stage(EPI_HEAD_TRUNK, 0, l0);
        }
        int l0 = A.n_loads;
        push_image_loads(A, L.hid_img[l], 4, true);
        if (l == L.trunk_hidden)   // first colour layer: view direction (+ grid features) from the input chunk
            push_input_loads(A, L.color0_ximg, 1, L.grid_channels > 0 ? 3 : 1, false);
        end_stage(EPI_FILM, (uint8_t)(l + 1), l0);
        A.stages[A.n_stages - 1].uniform = 1;
        if (A.n_loads > MAX_LOADS - 16 || A.n_stages > MAX_STAGES - 3) return false;
    }
    {
        int l0 = A.n_loads;
        push_head_loads(A, L.rgb_img, 8, 8);
        end_stage(EPI_HEAD_RGB, 0, l0);
    }
    return true;
}

}  // namespace

namespace {
long long* g_trace = nullptr;
}  // namespace

void set_fast_trace(long long* buf) { g_trace = buf; }
long long* get_fast_trace() { return g_trace; }

int siren_points_fast(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                      const float* film, int batch, long long ppb, int dir_group, int lock_dirs, float* out,
                      cudaStream_t st) {
    static_assert(sizeof(FastArgs) <= 4000, "kernel parameter block too large");
    static_assert(SMEM_TOTAL <= 232448, "one CTA per SM: 227 KB of shared memory");
    FN_REQUIRE(L.trunk_hidden >= 1 && L.n_hidden - L.trunk_hidden >= 1, "field needs >= 2 trunk and >= 1 colour layers");
    FN_REQUIRE(L.label_dim < 32, "the tcgen05 path packs labels and sigma into one 32-row head (label_dim <= 31)");
    FastArgs a;
    memset(&a, 0, sizeof(a));
    FN_REQUIRE(build_program(L, a), "field too deep for the stage program");
    a.L = L; a.packed = packed; a.points = points; a.dirs = dirs; a.film = film; a.out = out;
    a.ppb = ppb; a.tiles_per_batch = (ppb + TILE - 1) / TILE; a.n_tiles = a.tiles_per_batch * batch;
    const long long n_pairs = (a.n_tiles + 1) / 2;
    a.dir_group = dir_group < 1 ? 1 : dir_group; a.lock_dirs = lock_dirs;
    a.trace = g_trace;
    if (a.n_tiles <= 0) return 0;
    FN_REQUIRE(ppb % a.dir_group == 0, "points_per_batch %lld not a multiple of dir_group %d", ppb, a.dir_group);
    auto kernel = a.trace ? siren_fast_kernel<true> : siren_fast_kernel<false>;
    FN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    int blocks = (int)(n_pairs < (long long)num_sms() ? n_pairs : (long long)num_sms());
    kernel<<<blocks, NTHREADS, SMEM_TOTAL, st>>>(a);
    FN_LAUNCH_OK("siren_fast_kernel");
    return 0;
}

}  // namespace fn
