// Point network, FAST mode (tcgen05), third-generation kernel: two tiles per CTA, three 32 KB ring slots.
//
// What the measurements say (round 1: profiles/r01_*.txt; round 2: profiles/r02_fast3_ablation.txt, r02_microbench.txt,
// DESIGN.md section 5): per 128-point layer-tile the tensor pipe needs 2048 cycles (32 MMAs at the measured 64 cycles) and
// the epilogue ~2300-3500 (one sin per output element: MUFU.SIN is 8 cycles per warp instruction and sub-partition, the fp16
// pack shares that pipe, and a lone warp per sub-partition only reaches 13.7 cycles per element; two reach 10.3); a layer
// is a dependency chain MMA -> epilogue -> MMA; a ring slot's turnaround is ~1500 cycles.  Hence:
//   * TWO tiles per CTA (one CTA per SM) run the same layer program, each with its own MMA-issuer warp
//     and its own four epilogue warps, so one tile's epilogue overlaps the other tile's MMAs;
//   * ring slots are 32 KB = two k-chunks of one feature half = 8 MMAs = 512 tensor cycles per
//     barrier round; three slots (96 KB in flight), reused in global round-robin over both tiles' loads
//     (six 16 KB slots: measured 2-4 % slower);
//   * per layer the rounds run [h0 k01][h1 k01][h0 k23][h1 k23] and accumulator / operand hand-offs are
//     per feature half: half 0 is committed after round 3 (its epilogue overlaps round 4) and the next
//     layer's first round only needs half 0 of the previous epilogue;
//   * that ring only fits because the per-tile 16 KB input chunk is gone: the position slots of the
//     first layer live in activation chunk 3 (free at tile start), and the view-direction / grid
//     feature slots of the first colour layer are written into activation chunk 0 once that layer's
//     k01 rounds have retired; their MMAs go in front of the last round, so half 0 still completes a round early;
//   * activations are MN-major (points contiguous), so the feature-per-thread epilogue stores 16
//     bytes at a time; TMEM is drained in 16-column double-buffered pieces (a tcgen05.ld + wait is ~21 cycles);
//   * the trunk head and the first colour layer of a tile are issued back to back, the head accumulating into
//     half 1 so that the colour layer's first round does not wait for the head's epilogue; the head hands its
//     accumulator back before its global stores, and a tile's last [r, g, b, sigma] store is deferred past the next
//     tile's hand-off (a fence.proxy.async is a MEMBAR: it would wait for those stores);
//   * [r2] the warp index comes from a lane-0 shuffle: the compiler then keeps every descriptor / barrier address of
//     the issuers in uniform registers (6 instead of 18 instructions per MMA, no R2UR / ELECT): -5 % kernel time.
//     With that, tcgen05.mma / commit become warp-level UTCHMMA / UTCBAR and every wait in front of them must END
//     CONVERGED (mbar_wait_warp_spin: the loop exit is a warp vote);
//   * [r2] issuers and epilogue warps poll without the suspend-time hint (+2 %).
//
//   warps 0..2    weight producers (producer w serves ring loads it % 3 == w)
//   warps 3, 4    MMA issuers of tile X / tile Y (warp-converged, uniform-register operands)
//   warps 5..8    epilogue of tile X      warps 9..12   epilogue of tile Y
#include "common.cuh"
#include "siren_common.cuh"
#include "tc5.cuh"
#include <type_traits>

namespace fn {

namespace {

using namespace tc5;

constexpr int TILE = 128;
#ifdef FENERF_AB_SLAB16
constexpr int SLAB_CHUNKS = 1;                       // 6 slots of 16 KB (a slot recycles after 4 MMAs): measured 2-4 % slower
#else
constexpr int SLAB_CHUNKS = 2;                       // k-chunks (16 KB each) per ring slot: 3 slots of 32 KB
#endif
constexpr int RING = 6 / SLAB_CHUNKS;
constexpr int PROD = 3;                              // producer warps; producer w serves loads it % PROD == w
constexpr int MMA_WARP = PROD;                       // issuer of tile X; MMA_WARP + 1 issues tile Y
constexpr int EPI_WARP0 = PROD + 2;
#ifdef FENERF_AB_EPI8
constexpr int EPI_SPLIT = 2;                         // measured: no faster for model A (the epilogue is not latency-bound per warp),
#else                                                // slower for model B (672 threads cap the kernel at 80 registers)
constexpr int EPI_SPLIT = 1;                         // epilogue warps per TMEM lane quadrant and tile
#endif
constexpr int EPI_WARPS = 4 * EPI_SPLIT;             // per tile
constexpr int NTHREADS = (EPI_WARP0 + 2 * EPI_WARPS) * 32;      // 416
constexpr uint32_t CHUNK_BYTES = 16384;
constexpr uint32_t STAGE_BYTES = SLAB_CHUNKS * CHUNK_BYTES;
constexpr uint32_t TILE_SMEM = 4 * CHUNK_BYTES;      // four activation chunks per tile
constexpr uint32_t SMEM_RING = 2 * TILE_SMEM;
constexpr uint32_t SMEM_BAR = SMEM_RING + RING * STAGE_BYTES;   // 229376
constexpr uint32_t SMEM_TAB = SMEM_BAR + 256;                     // per-tile program: loads, then stages
constexpr int TMEM_COLS = 512;
constexpr int MAX_LOADS = 112;
constexpr int MAX_STAGES = 24;
constexpr uint32_t SMEM_TOTAL = SMEM_TAB + MAX_LOADS * 16 + MAX_STAGES * 8;   // 231616 (of 232448)
// The MMA issuers poll their barriers without the suspend-time hint, and every wait ends in a warp vote
// (tc5.cuh: mbar_wait_warp_spin).  Measured (tools/ablate_fast3.sh): 2 % faster than the sleeping try_wait -- and with six
// 16 KB ring slots the hinted try_wait on the ring's transaction barriers produced wrong results from a single tile on
// (cause not identified; no wait of this kernel uses the hinted form any more).
#define FN_CTRL_WAIT mbar_wait_warp_spin
// Ring order.  Both tiles of a pair run the same flat load program L[0..n); the ring interleaves them load by load with tile Y
// SEQ_LAG loads behind tile X:  X0 .. X(lag-1), then X(lag) Y0 X(lag+1) Y1 ... and Y's tail.  A slot is reused three entries later,
// so the slab of a tile's round r waits for the OTHER tile's round r-1-lag/2... of half a layer earlier, which that tile
// consumed long before (with the stage-by-stage order X r1..r4 Y r1..r4 a tile's third round waited for the other tile's last).
#ifndef FENERF_AB_SEQ_LAG
#define FENERF_AB_SEQ_LAG 2
#endif
constexpr uint32_t SEQ_LAG = FENERF_AB_SEQ_LAG;
// Shared ring (experiment switch).  Every ring load is weight data, the same for both tiles of a pair: with SHARE_RING each slab is
// loaded ONCE per pair and read by both issuers (one `full` barrier per slot that both wait on, `empty` counts two commits), which
// halves the L2 -> shared-memory weight stream and doubles the time a slot may take to turn around.  The two tiles then run at
// most RING rounds apart (the leader waits for the follower to release the slot).  With a single tile in the pair the idle issuer
// releases the slots without issuing MMAs.
#ifndef FENERF_AB_SHARE
#define FENERF_AB_SHARE 0
#endif
constexpr bool SHARE_RING = FENERF_AB_SHARE != 0;
// The flat (per-load) ring orders above are compiled in only with -DFENERF_AB_FLAT=1; the default build keeps the stage-by-stage
// order  X s, Y s, X s+1, ...  that every committed measurement was taken with.
#ifndef FENERF_AB_FLAT
#define FENERF_AB_FLAT (FENERF_AB_SHARE != 0)
#endif
#define FN_PROD_WAIT mbar_wait_poll             // (single lane; measured no different from the hinted form, kept uniform with the rest)
#define FN_EPI_WAIT mbar_wait_warp_spin     // epilogue warps wait converged as well (tcgen05.ld is .sync.aligned); +0.5-1 % over the hinted form
#ifdef FENERF_AB_LD32
constexpr int GW = 32;                               // TMEM columns per tcgen05.ld in the FiLM epilogue
#else
constexpr int GW = 16;
#endif

enum : uint8_t { EPI_FILM = 0, EPI_HEAD_TRUNK = 1, EPI_HEAD_RGB = 2 };
enum : uint8_t { X_NONE = 0, X_POS = 1, X_EXTRA = 2 };   // load reads the K-major input slots instead of an activation chunk

struct alignas(16) LoadOp {
    uint32_t src;          // byte offset in the packed buffer
    uint16_t bytes16;      // bytes / 16
    uint8_t x_chunk;       // first activation chunk read (0..3)
    uint8_t n_chunks;      // k-chunks in this load, `bytes / n_chunks` apart in the slot
    uint8_t k0, nk;        // K-steps inside a 64-wide chunk
    uint8_t n8;            // MMA N / 8
    uint8_t half;          // accumulator half; for X loads: both halves (weights at 16 KB stride), half ignored
    uint8_t first;         // 1: the first MMA of this load overwrites the accumulator
    uint8_t w_is_a;        // 1: weights are the A operand (transposed FiLM layer); 0: B (head)
    uint8_t xkind;         // X_*
    uint8_t pad;
};
static_assert(sizeof(LoadOp) == 16, "LoadOp is 16 bytes");

struct StageOp {
    uint8_t epi;           // EPI_*
    uint8_t film;          // FiLM layer index
    uint8_t n_loads;
    uint8_t uniform;       // 1: 256x256 FiLM layer = 8 / SLAB_CHUNKS slab loads [+ the X_EXTRA loads]
    uint8_t xsync;         // 1: a fifth, X_EXTRA load follows: before it the issuer commits `xmain` and waits `xready`
    uint8_t l0;            // index of the stage's first load in Fast3Args::loads
    uint8_t fuse_next;     // 1: the issuer runs the next stage of the SAME tile before turning to the other tile
    uint8_t pad;
};

struct Fast3Args {
    LoadOp loads[MAX_LOADS];
    StageOp stages[MAX_STAGES];
    int n_loads, n_stages;
    FnLayout L;
    const unsigned char* packed;
    const float* points;
    const float* dirs;
    const float* film;
    float* out;
    float* sigma_out;      // optional compact copy of the density channel, one float per point (the resampler's input)
    long long ppb, tiles_per_batch, n_tiles;
    int dir_group, lock_dirs;
    int sigma_only;        // the program stops after the trunk head; only out[..., C-1] is written
    long long* trace;
};

// Timing experiments (WRONG results by design; tools/ablate_fast3.sh): -DFENERF_ABLATE=<mask> builds a variant without
//   1 the weight bytes (1 KB loads)   2 the FiLM epilogue's work (hand-offs only)   4 tcgen05.mma (commits only)
//   8 sin   16 the activation stores   32 tcgen05.ld
#ifdef FENERF_ABLATE
#define FN_DBG(bit) ((FENERF_ABLATE) & (bit))
#else
#define FN_DBG(bit) 0
#endif

template <bool kTrace>
#ifdef FENERF_AB_MAXNREG
__global__ void __maxnreg__(FENERF_AB_MAXNREG) siren_fast3_kernel(const __grid_constant__ Fast3Args a) {
#else
__global__ void __launch_bounds__(NTHREADS, 1) siren_fast3_kernel(const __grid_constant__ Fast3Args a) {
#endif
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    // the warp index through a lane-0 broadcast: the compiler then knows it (and every ring / descriptor address derived
    // from the role and tile index) is warp-uniform and keeps the issuer's operand math in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t bar_full = sbase + SMEM_BAR;             // [RING][2]: slot filled with a load of tile t (each issuer
                                                            // only ever waits on its own tile's barriers, in order)
    const uint32_t bar_empty = bar_full + 16 * RING;        // [RING]
    // Accumulator / operand hand-offs are per tile AND per feature half h (index t * 2 + h): half h of the
    // accumulator feeds activation chunks 2h, 2h+1 of the next layer, so the next layer's [h0 k01] MMAs can
    // start while the epilogue is still working on half 1, and the epilogue of half 0 starts while the
    // [h1 k23] MMAs are still running.
    const uint32_t bar_acc = bar_empty + 8 * RING;          // [2][2] accumulator half complete (issuer -> epilogue t)
    const uint32_t bar_aready = bar_acc + 32;               // [2][2] chunks 2h,2h+1 written + half h drained (epilogue t -> issuer)
    const uint32_t bar_xmain = bar_aready + 32;             // [2] colour layer 0: k01 MMAs retired, chunk 0 reusable
    const uint32_t bar_xready = bar_xmain + 16;             // [2] colour layer 0: extra input slots written into chunk 0
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SMEM_BAR + 8 * (3 * RING + 12));

    // The program tables are read once per load / stage by single threads on latency-critical paths
    // (producer turnaround, issuer phase changes): keep them in shared memory, not in the constant bank.
    LoadOp* s_loads = reinterpret_cast<LoadOp*>(smem + SMEM_TAB);
    StageOp* s_stages = reinterpret_cast<StageOp*>(smem + SMEM_TAB + MAX_LOADS * sizeof(LoadOp));
    for (int i = threadIdx.x; i < a.n_loads; i += NTHREADS) s_loads[i] = a.loads[i];
    for (int i = threadIdx.x; i < a.n_stages; i += NTHREADS) s_stages[i] = a.stages[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) {
#if FENERF_AB_FLAT
            mbar_init(bar_full + 16 * i, 1); mbar_init(bar_full + 16 * i + 8, 1); mbar_init(bar_empty + 8 * i, SHARE_RING ? 2 : 1);
#else
            mbar_init(bar_full + 16 * i, 1); mbar_init(bar_full + 16 * i + 8, 1); mbar_init(bar_empty + 8 * i, 1);
#endif
        }
        for (int t = 0; t < 2; ++t) {
            for (int h = 0; h < 2; ++h) {
                mbar_init(bar_acc + 8 * (t * 2 + h), 1);
                mbar_init(bar_aready + 8 * (t * 2 + h), EPI_WARPS);
            }
            mbar_init(bar_xmain + 8 * t, 1);
            mbar_init(bar_xready + 8 * t, 4);
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

    if (warp < PROD) {
        // ================= weight producers: global load number `it` -> slot it % RING = producer warp it % RING.
        // (Round-robin over ALL loads, not per stage: a slot is then always reused RING loads later, so the
        // first load of a phase was requested two rounds before the previous phase ended.)
        if (lane == 0) {
            uint32_t it = 0, empty_par = 0;         // bit s: parity of the next wait on empty[s] (starts "free")
#if FENERF_AB_FLAT
            const uint32_t n_flat = (uint32_t)a.n_loads;
            auto emit = [&](int t, uint32_t li) {
                if ((int)(it % PROD) == warp) {
                    const uint32_t slot = it % RING;
                    const LoadOp op = s_loads[li];                      // before the wait: off the turnaround path
                    uint32_t bytes = (uint32_t)op.bytes16 * 16u;
                    if (FN_DBG(1)) bytes = 1024u;                       // timing experiment only: wrong results
                    const unsigned char* src = a.packed + op.src;
                    FN_PROD_WAIT(bar_empty + 8 * slot, ((empty_par >> slot) & 1u) ^ 1u);
                    empty_par ^= 1u << slot;
                    mbar_arrive_expect_tx(bar_full + 16 * slot + 8 * t, bytes);
                    bulk_g2s(sbase + SMEM_RING + slot * STAGE_BYTES, src, bytes, bar_full + 16 * slot + 8 * t);
                }
                ++it;
            };
#endif
            for (long long pair = blockIdx.x; pair * 2 < a.n_tiles; pair += gridDim.x) {
                const int nt = (pair * 2 + 1 < a.n_tiles) ? 2 : 1;
#if FENERF_AB_FLAT
                const uint32_t n_iter = n_flat + ((nt == 2 && !SHARE_RING) ? SEQ_LAG : 0u);
                if (SHARE_RING) {
                    for (uint32_t i = 0; i < n_flat; ++i) emit(0, i);
                } else
                for (uint32_t i = 0; i < n_iter; ++i) {
                    if (i < n_flat) emit(0, i);
                    if (nt == 2 && i >= SEQ_LAG && i - SEQ_LAG < n_flat) emit(1, i - SEQ_LAG);
#else
                for (int s = 0; s < a.n_stages;) {
                    int s_end = s + 1;
                    while (s_stages[s_end - 1].fuse_next) ++s_end;
                    for (int t = 0; t < nt; ++t)
                        for (int ss = s; ss < s_end; ++ss) {
                            const int n = s_stages[ss].n_loads, li = s_stages[ss].l0;
                            for (int j = 0; j < n; ++j, ++it) {
                                if ((int)(it % PROD) != warp) continue;
                                const uint32_t slot = it % RING;
                                const LoadOp op = s_loads[li + j];          // before the wait: off the turnaround path
                                uint32_t bytes = (uint32_t)op.bytes16 * 16u;
                                if (FN_DBG(1)) bytes = 1024u;               // timing experiment only: wrong results
                                const unsigned char* src = a.packed + op.src;
                                FN_PROD_WAIT(bar_empty + 8 * slot, ((empty_par >> slot) & 1u) ^ 1u);
                                empty_par ^= 1u << slot;
                                mbar_arrive_expect_tx(bar_full + 16 * slot + 8 * t, bytes);
                                bulk_g2s(sbase + SMEM_RING + slot * STAGE_BYTES, src, bytes, bar_full + 16 * slot + 8 * t);
                            }
                        }
                    s = s_end;
#endif
                }
            }
        }
    } else if (warp == MMA_WARP || warp == MMA_WARP + 1) {
        // ================= MMA issuers, one per tile (warp-converged; an elected lane issues) =================
        // Both walk the same global load sequence (the ring order); each issues only its own tile's phases,
        // so the ~1000 cycles of commit / barrier latency between two phases of one tile overlap with the
        // other issuer's MMAs instead of idling the tensor pipe.
        const int t = warp - MMA_WARP;
        uint32_t full_par = 0;                      // bit s: phase parity of full[s][t] this tile waits for next
#if FENERF_AB_FLAT
        uint32_t it = 0;                            // flat load index of this tile's program within the current pair
        uint32_t seq_base = 0;                      // ring position of the pair's first entry
        int nt_cur = 2;
        const uint32_t n_flat = (uint32_t)a.n_loads;
        // ring slot of this tile's load `it` (the producers' emission order, see SEQ_LAG)
        auto slot_now = [&]() -> uint32_t {
            uint32_t pos;
            if (SHARE_RING || nt_cur == 1) pos = it;
            else if (t == 0) pos = it < SEQ_LAG ? it : 2u * it - SEQ_LAG;
            else pos = (it + SEQ_LAG + 1u < n_flat ? it + SEQ_LAG + 1u : n_flat) + it;
            return (seq_base + pos) % RING;
        };
#else
        uint32_t it = 0;                            // global load number (slot = it % RING), as in the producers
#endif
        uint32_t n_ready = 0, n_x = 0;
        Tracer<kTrace> tr(lane == 0 ? a.trace : nullptr, t == 0 ? 1 : 0);
        const uint32_t ring_lo = (sbase + SMEM_RING) >> 4;
        constexpr uint32_t kSlot16 = STAGE_BYTES >> 4, kChunk16 = CHUNK_BYTES >> 4;
        // one ring slot's "full" wait for a single-load step (slot number is a runtime value here)
        // Every wait of the issuer ends with a warp vote (mbar_wait_warp_spin): with the operand math in uniform registers ptxas
        // emits tcgen05.mma / commit as warp-level UTCHMMA / UTCBAR without the elect.sync, and lanes that left a per-thread
        // try_wait loop one by one issued them once per warp fragment (seen: K-steps accumulated twice).
        auto ctrl_wait = [&](uint32_t bar, uint32_t parity) { FN_CTRL_WAIT(bar, parity); };
        auto wait_full = [&](uint32_t slot) {
#if FENERF_AB_FLAT
            ctrl_wait(bar_full + 16 * slot + (SHARE_RING ? 0 : 8 * t), (full_par >> slot) & 1u);
#else
            ctrl_wait(bar_full + 16 * slot + 8 * t, (full_par >> slot) & 1u);
#endif
            tc_fence_after();
            full_par ^= 1u << slot;
        };
        // stage flags as register bit masks: a phase change costs no memory access
        uint32_t m_uniform = 0, m_xsync = 0, m_fuse = 0;
        for (int i = 0; i < a.n_stages; ++i) {
            m_uniform |= (uint32_t)(s_stages[i].uniform != 0) << i;
            m_xsync |= (uint32_t)(s_stages[i].xsync != 0) << i;
            m_fuse |= (uint32_t)(s_stages[i].fuse_next != 0) << i;
        }
        int tl = 0;
        for (long long pair = blockIdx.x; pair * 2 < a.n_tiles; pair += gridDim.x, ++tl) {
            const int nt = (pair * 2 + 1 < a.n_tiles) ? 2 : 1;
#if FENERF_AB_FLAT
            nt_cur = nt;
            it = 0;
            if (t < nt) {
                {
                    for (int ss = 0; ss < a.n_stages; ++ss) {
#else
            for (int s = 0; s < a.n_stages;) {
                int s_end = s + 1;
                while ((m_fuse >> (s_end - 1)) & 1u) ++s_end;
                for (int tt = 0; tt < nt; ++tt) {
                    for (int ss = s; ss < s_end; ++ss) {
#endif
                        const bool st_uniform = (m_uniform >> ss) & 1u, st_xsync = (m_xsync >> ss) & 1u;
#if FENERF_AB_FLAT
#else
                        if (tt != t) {                       // the other issuer's phase: only the ring position moves
                            it += s_stages[ss].n_loads;
                            continue;
                        }
#endif
                        tr.log('B', tl, ss, t);
                        const uint32_t rdy_par = n_ready & 1;
                        ++n_ready;
                        ctrl_wait(bar_aready + 8 * (t * 2), rdy_par);
                        // a plain FiLM layer observes half 1 only before its [h1 k01] round (inside `issue`)
                        if (!st_uniform) ctrl_wait(bar_aready + 8 * (t * 2 + 1), rdy_par);
                        tc_fence_after();
                        tr.log('A', tl, ss, t);
                        const uint32_t x_lo0 = (sbase + t * TILE_SMEM) >> 4;      // activation chunk 0 of this tile
                        const uint32_t d0 = tmem_base + (uint32_t)t * 256u;
                        // K-major input slots (positions in chunk 3 / direction + grid features in chunk 0) against
                        // the [256 features][64 slots] weight image: one MMA group per feature half
                        auto x_load = [&](const LoadOp op) {
#if FENERF_AB_FLAT
                            const uint32_t slot = slot_now();
#else
                            const uint32_t slot = it % RING;
#endif
                            wait_full(slot);
                            tr.log('F', tl, ss, t * 64 + 4);
                            const uint32_t x_lo = x_lo0 + (op.xkind == X_POS ? 3u : 0u) * kChunk16;
                            constexpr uint32_t idesc = umma_idesc_f16(TILE, 0, 0);
#pragma unroll
                            for (int hh = 0; hh < SLAB_CHUNKS; ++hh) {       // a 32 KB slot holds both halves, a 16 KB slot op.half
                                const uint32_t h = SLAB_CHUNKS == 2 ? (uint32_t)hh : (uint32_t)op.half;
                                const uint32_t w_lo = ring_lo + slot * kSlot16 + (uint32_t)hh * kChunk16;
#pragma unroll
                                for (int k = 0; k < 3; ++k)
                                    if (k < op.nk) {
                                        const uint32_t ko = (uint32_t)(op.k0 + k) * 2u;
                                        tc_mma_f16_elect(d0 + h * 128, kDescHi | (uint64_t)(w_lo + ko), kDescHi | (uint64_t)(x_lo + ko),
                                                         idesc, (op.first && k == 0) ? 0u : 1u);
                                    }
                            }
                            tc_commit_elect(bar_empty + 8 * slot);
                            ++it;
                        };
                        if (st_uniform) {
                            // The layer's eight (half, k-chunk) pieces in the order [h0 c0][h0 c1][h1 c0][h1 c1][h0 c2][h0 c3][h1 c2][h1 c3],
                            // SLAB_CHUNKS of them per ring slot, 4 MMAs each.  One compact loop for every starting slot and both
                            // issuers (all operand math stays in uniform registers).
                            constexpr uint32_t idesc = umma_idesc_f16(TILE, 0, 1);          // B (activations) MN-major
#pragma unroll 1
                            for (int r = 0; r < 8 / SLAB_CHUNKS; ++r) {
                                // the first h1 piece needs accumulator half 1 drained (and, later, chunks 2,3)
                                if (r * SLAB_CHUNKS == 2) ctrl_wait(bar_aready + 8 * (t * 2 + 1), rdy_par);
#if FENERF_AB_FLAT
                                const uint32_t slot = slot_now();
#else
                                const uint32_t slot = it % RING;
#endif
                                wait_full(slot);
                                tr.log('F', tl, ss, t * 64 + r);
#pragma unroll
                                for (int cc = 0; cc < SLAB_CHUNKS; ++cc) {
                                    const uint32_t idx = (uint32_t)(r * SLAB_CHUNKS + cc);
                                    const uint32_t half = (idx >> 1) & 1u, chunk = (idx >> 2) * 2u + (idx & 1u);
                                    const uint32_t w_lo = ring_lo + slot * kSlot16 + (uint32_t)cc * kChunk16;
                                    const uint32_t x_lo = x_lo0 + chunk * kChunk16;
                                    const uint32_t d = d0 + half * 128u;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (!FN_DBG(4))
                                            tc_mma_f16_elect(d, kDescHi | (uint64_t)(w_lo + 2 * k), kDescHiMN | (uint64_t)(x_lo + 256 * k), idesc,
                                                             (chunk == 0 && k == 0) ? 0u : 1u);
                                }
                                tc_commit_elect(bar_empty + 8 * slot);
                                ++it;
                                tr.log('I', tl, ss, t * 64 + r);
                                const int done = (r + 1) * SLAB_CHUNKS;      // pieces issued so far
                                // first colour layer: chunks 0/1 have been read for the last time once the k01 pieces
                                // retire; the epilogue overwrites chunk 0 with the extra input slots while [h0 c2][h0 c3] run
                                if (done == 4 && st_xsync) tc_commit_elect(bar_xmain + 8 * t);
                                if (done == 6) {
                                    if (st_xsync) {
                                        // the extra-input MMAs (both halves) go in front of the last h1 pieces, so that half 0 still
                                        // completes one round before the stage ends and its epilogue overlaps that round
                                        ctrl_wait(bar_xready + 8 * t, n_x & 1);
                                        ++n_x;
                                        tc_fence_after();
                                        tr.log('X', tl, ss, t);
                                        for (int xl = 0; xl < 2 / SLAB_CHUNKS; ++xl) x_load(s_loads[s_stages[ss].l0 + 6 / SLAB_CHUNKS + xl]);
                                    }
                                    // half 0 is complete after [h0 c3] (and chunks 0,1 were last read by [h1 c1])
                                    tc_commit_elect(bar_acc + 8 * (t * 2));
                                }
                            }
                        } else if (s_loads[s_stages[ss].l0].xkind != X_NONE) {
                            for (int xl = 0; xl < 2 / SLAB_CHUNKS; ++xl) x_load(s_loads[s_stages[ss].l0 + xl]);
                        } else {
                            // head: activations are the A operand (M = 128 points, MN-major), the [n rows][256] head
                            // image the B operand; 4 k-chunks x 4 K-steps, fully unrolled
                            const LoadOp op = s_loads[s_stages[ss].l0];
#if FENERF_AB_FLAT
                            const uint32_t slot = slot_now();
#else
                            const uint32_t slot = it % RING;
#endif
                            wait_full(slot);
                            tr.log('F', tl, ss, t * 64);
                            const uint32_t idesc = umma_idesc_f16((uint32_t)op.n8 * 8u, 1u, 0u);
                            const uint32_t w_stride16 = ((uint32_t)op.bytes16) >> 2;      // bytes / 4 chunks / 16
                            const uint32_t w_lo = ring_lo + slot * kSlot16;
#pragma unroll
                            for (int c = 0; c < 4; ++c)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc_mma_f16_elect(d0 + (uint32_t)op.half * 128u, kDescHiMN | (uint64_t)(x_lo0 + c * kChunk16 + 256 * k),
                                                     kDescHi | (uint64_t)(w_lo + c * w_stride16 + 2 * k), idesc, (c | k) ? 1u : 0u);
                            tc_commit_elect(bar_empty + 8 * slot);
                            ++it;
                        }
                        if (!st_uniform) tc_commit_elect(bar_acc + 8 * (t * 2));
                        tc_commit_elect(bar_acc + 8 * (t * 2 + 1));
                        tr.log('C', tl, ss, t);
                    }
                }
#if FENERF_AB_FLAT
#else
                s = s_end;
#endif
            }
#if FENERF_AB_FLAT
            else if (SHARE_RING) {
                // single tile in this pair: the idle issuer still releases every slot (empty counts two arrivals)
                for (uint32_t i = 0; i < n_flat; ++i, ++it) {
                    const uint32_t slot = slot_now();
                    wait_full(slot);
                    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
                    __syncwarp();
                }
            }
            seq_base += (nt == 2 && !SHARE_RING) ? 2u * n_flat : n_flat;
#endif
        }
    } else {
        // ================= epilogue warps =================
        const int t = (warp - EPI_WARP0) / EPI_WARPS;  // which tile of the pair
        const int j = ((warp - EPI_WARP0) >> 2) % EPI_SPLIT;   // which 128 / EPI_SPLIT points of the tile in the FiLM epilogues;
                                                       // everything per point (input slots, heads) is done by the j == 0 warps
        const int q = warp & 3;                        // TMEM lane quadrant (hardware: warp id % 4)
        const int row = q * 32 + lane;                 // feature within a half (FiLM) / point (heads, input slots)
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)t * 256u;
        unsigned char* tsm = smem + t * TILE_SMEM;
        const uint32_t my_acc = bar_acc + 16 * t, my_aready = bar_aready + 16 * t;     // + 8 * h
        const uint32_t my_xmain = bar_xmain + 8 * t, my_xready = bar_xready + 8 * t;
        const uint32_t xrow_off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;   // K-major row of this point
        const uint32_t xsw = (uint32_t)(row & 7);
        const int C = L.out_dim;
        const float* sigma_w = reinterpret_cast<const float*>(a.packed + L.sigma_w);
        const float* rgb_w = reinterpret_cast<const float*>(a.packed + L.rgb_w);
        const float* label_w = reinterpret_cast<const float*>(a.packed + L.label_w);
        uint32_t n_acc = 0, n_x = 0;
        // traced: quadrant-0 warp of each tile (roles 2, 3); roles 1 / 0 are the issuers of tile X / Y
        Tracer<kTrace> tr(q == 0 && j == 0 && lane == 0 ? a.trace : nullptr, 2 + t);
        int tl = 0;
        // position (already box-warped) and view direction of this thread's point in tile `tile_i`; zeros past the end
        auto load_inputs = [&](long long tile_i, float (&pos)[3], float (&dir)[3]) {
            const long long bb = tile_i / a.tiles_per_batch;
            const long long pp = (tile_i % a.tiles_per_batch) * TILE + row;
#pragma unroll
            for (int i = 0; i < 3; ++i) { pos[i] = 0.f; dir[i] = 0.f; }
            if (pp < a.ppb) {
                const long long ff = bb * a.ppb + pp;
#pragma unroll
                for (int i = 0; i < 3; ++i) pos[i] = __fmul_rn(a.points[ff * 3 + i], L.input_scale);
                if (a.lock_dirs) dir[2] = -1.f;
                else {
                    const long long di = bb * (a.ppb / a.dir_group) + pp / a.dir_group;
#pragma unroll
                    for (int i = 0; i < 3; ++i) dir[i] = a.dirs[di * 3 + i];
                }
            }
        };
        float next_pos[3], next_dir[3];      // fetched during the previous tile's last stage (off the tile-start path)
        bool have_next = false;
        float sig_keep = 0.f;                // density of this thread's point: held from the trunk head to the rgb head so
                                             // that [.., r, g, b, sigma] leaves in 16-byte (C = 4) / 8-byte stores
        // The [r, g, b, sigma] tail of the point's output row is the tail of a tile: its store is deferred until the NEXT tile's
        // input slots have been handed to the issuer (the fence in front of that hand-off would wait for the store otherwise).
        // 16 bytes at once when the row is 16 bytes (C = 4), two 8-byte stores when C is even, scalars otherwise.
        float4 pend_v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pend_p = nullptr;
        auto flush_tail = [&]() {
            if (pend_p == nullptr) return;
            const uintptr_t ob = reinterpret_cast<uintptr_t>(a.out);
#ifdef FENERF_AB_SCALAR_STORES
            if (false) {
#else
            if (L.out_dim == 4 && (ob & 15) == 0) {
#endif
                *reinterpret_cast<float4*>(pend_p) = pend_v;
            } else if ((L.out_dim & 1) == 0 && (ob & 7) == 0) {
                *reinterpret_cast<float2*>(pend_p) = make_float2(pend_v.x, pend_v.y);
                *reinterpret_cast<float2*>(pend_p + 2) = make_float2(pend_v.z, pend_v.w);
            } else {
                pend_p[0] = pend_v.x; pend_p[1] = pend_v.y; pend_p[2] = pend_v.z; pend_p[3] = pend_v.w;
            }
            pend_p = nullptr;
        };
        for (long long pair = blockIdx.x; pair * 2 + t < a.n_tiles; pair += gridDim.x, ++tl) {
            const long long tile = pair * 2 + t;
            const long long b = tile / a.tiles_per_batch;
            const long long pnt = (tile % a.tiles_per_batch) * TILE + row;
            const bool valid = pnt < a.ppb;
            const long long flat = b * a.ppb + pnt;
            tr.log('T', tl, 0, 0);
            // ---- input slots of this thread's point (layout.h): positions now (chunk 3), direction + grid
            //      features kept in registers until the first colour layer ----
            uint4 xslots[8];
            if (j == 0) {
                float pos[3], dir[3];
                if (have_next) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) { pos[i] = next_pos[i]; dir[i] = next_dir[i]; }
                } else {
                    load_inputs(tile, pos, dir);
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
                if (L.grid_channels > 0 && valid && !a.sigma_only) {     // density needs no grid features
                    float feat[32];
                    grid_features32_h(reinterpret_cast<const __half*>(a.packed + L.grid16), L.grid_res, pos[0], pos[1], pos[2], feat);
#pragma unroll
                    for (int i = 0; i < 32; ++i) slots[FN_SLOT_FEAT + i] = __float2half_rn(feat[i]);
                }
                const uint4* src = reinterpret_cast<const uint4*>(slots);
#pragma unroll
                for (int i = 0; i < 8; ++i) xslots[i] = src[i];
                // positions: K-step 0 = pieces 0,1 of the row, into activation chunk 3 (free at tile start)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    *reinterpret_cast<uint4*>(tsm + 3 * CHUNK_BYTES + xrow_off + (((uint32_t)i ^ xsw) << 4)) = xslots[i];
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(my_aready); mbar_arrive(my_aready + 8); }
            flush_tail();

            for (int s = 0; s < a.n_stages; ++s) {
                const StageOp sop = s_stages[s];
                const bool last = (s + 1 == a.n_stages);
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
                    if (sop.xsync && j == 0) {
                        // first colour layer: once its main MMAs have retired, chunk 0 takes the K-major extra
                        // input slots (pieces 2..7 of the row = direction and grid features)
                        FN_EPI_WAIT(my_xmain, n_x & 1);
                        tc_fence_after();
#pragma unroll
                        for (int i = 2; i < 8; ++i)
                            *reinterpret_cast<uint4*>(tsm + xrow_off + (((uint32_t)i ^ xsw) << 4)) = xslots[i];
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(my_xready);
                        ++n_x;
                    }
                    const uint32_t kk = (uint32_t)(fl & 63);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        FN_EPI_WAIT(my_acc + 8 * h, n_acc & 1);
                        tc_fence_after();
                        if (h == 0) tr.log('W', tl, s, 0);
                        // feature f = h*128 + fl -> chunk f/64, row k = f%64 of the MN-major chunk
                        // [k/8][point/64][k%8][64 points]
                        unsigned char* rowp = tsm + (uint32_t)(h * 2 + (fl >> 6)) * CHUNK_BYTES + (kk >> 3) * 2048u + (kk & 7u) * 128u;
                        const uint32_t sw = kk & 7u;
                        const float f_h = fr[h], p_h = ph[h];
                        uint32_t r[2][GW];
                        constexpr int NG = 128 / GW / EPI_SPLIT;             // groups of GW TMEM columns (= points) per warp
                        const uint32_t col0 = (uint32_t)(j * (128 / EPI_SPLIT));
                        if (FN_DBG(2 | 32)) {
#pragma unroll
                            for (int i = 0; i < GW; ++i) r[0][i] = r[1][i] = 0x3f000000u + (uint32_t)(i + lane);
                        }
                        if (!FN_DBG(2)) {
                        if (!FN_DBG(32)) tc_ld(t_lane + h * 128 + col0, r[0]);
#pragma unroll
                        for (int g = 0; g < NG; ++g) {
                            if (!FN_DBG(32)) {
                                tc_wait_ld();
                                if (g + 1 < NG) tc_ld(t_lane + h * 128 + col0 + (g + 1) * GW, r[(g + 1) & 1]);
                            }
#pragma unroll
                            for (int j8 = 0; j8 < GW / 8; ++j8) {
                                float v[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    // (sin on the FMA pipe -- range reduction in turns + a degree-7 polynomial -- for 2 or 3 of
                                    // every 8 outputs was measured 8 % / 28 % SLOWER: the epilogue is bound by issue slots and
                                    // single-warp latency at least as much as by the MUFU pipe)
                                    const float u = fmaf(f_h, __uint_as_float(r[g & 1][j8 * 8 + i]), p_h);
                                    v[i] = FN_DBG(8) ? u : __sinf(u);
                                }
                                uint4 pk;
                                pk.x = pack_half2(v[0], v[1]); pk.y = pack_half2(v[2], v[3]);
                                pk.z = pack_half2(v[4], v[5]); pk.w = pack_half2(v[6], v[7]);
                                // which group of 8 points (0..15): 64-point atoms are 1024 B apart, 16-byte pieces swizzled inside
                                const uint32_t pt8 = (col0 >> 3) + (uint32_t)(g * (GW / 8) + j8);
                                if (!FN_DBG(16) || pk.x == 0x12345678u)
                                    *reinterpret_cast<uint4*>(rowp + (pt8 >> 3) * 1024u + (((pt8 & 7u) ^ sw) << 4)) = pk;
                            }
                        }
                        }
                        // chunks 2h, 2h+1 written, accumulator half h drained
                        if (h == 0) tr.log('H', tl, s, 0);
                        fence_async_smem();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(my_aready + 8 * h);
                    }
                    ++n_acc;
                } else {
                    if (last && j == 0) {
                        // the inputs of this thread's next point: requested now, consumed at the next tile start
                        const long long ntile = (pair + gridDim.x) * 2 + t;
                        have_next = ntile < a.n_tiles;
                        if (have_next) load_inputs(ntile, next_pos, next_dir);
                    }
                    const bool early_h0 = !last && sop.epi == EPI_HEAD_TRUNK;
                    if (early_h0) {
                        // the trunk head's result sits in accumulator half 1: chunks 0,1 and accumulator half 0 stay as the last
                        // FiLM layer left them, so the colour layer's [h0 k01] round may start without waiting for this epilogue
                        __syncwarp();
                        if (lane == 0) mbar_arrive(my_aready);
                    }
                    FN_EPI_WAIT(my_acc, n_acc & 1);
                    FN_EPI_WAIT(my_acc + 8, n_acc & 1);
                    ++n_acc;
                    tc_fence_after();
                    tr.log('W', tl, s, 0);
                    if (j != 0) {
                        // heads are per point: the j == 0 warps own them; the others only keep the hand-off protocol
                        if (!last && sop.epi == EPI_HEAD_TRUNK) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(my_aready + 8);
                        }
                    } else if (sop.epi == EPI_HEAD_TRUNK) {
                        if (L.label_dim > 0) {
                            uint32_t r[32];
                            tc_ld32(t_lane + 128, r);
                            tc_wait_ld();
                            if (!last) {      // accumulator half 1 is drained: hand it back before the global stores (their latency
                                              // would otherwise sit in the colour layer's [h1 k01] round via the MEMBAR of a fence)
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(my_aready + 8);
                            }
                            if (valid) {
                                const float inv_scale = __ldg(label_w + FENERF_MAX_LABEL * FN_H + FENERF_MAX_LABEL);
                                const float* lb = label_w + FENERF_MAX_LABEL * FN_H;
                                float* orow = a.out + flat * C;
                                // a point's row is 4C bytes: with an even C the label pairs go out as 8-byte stores
                                const bool pair_ok = ((C & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 7) == 0);
#pragma unroll
                                for (int o = 0; o < 32; o += 2) {
                                    const float v0 = fmaf(__uint_as_float(r[o]), inv_scale, __ldg(lb + o));
                                    const float v1 = fmaf(__uint_as_float(r[o + 1]), inv_scale, __ldg(lb + o + 1));
                                    if (a.sigma_only) {
                                    } else if (o + 1 < L.label_dim && pair_ok) {
                                        *reinterpret_cast<float2*>(orow + o) = make_float2(v0, v1);
                                    } else {
                                        if (o < L.label_dim) orow[o] = v0;
                                        if (o + 1 < L.label_dim) orow[o + 1] = v1;
                                    }
                                    if (o == L.label_dim) sig_keep = __uint_as_float(r[o]) + __ldg(sigma_w + FN_H);
                                    if (o + 1 == L.label_dim) sig_keep = __uint_as_float(r[o + 1]) + __ldg(sigma_w + FN_H);
                                }
                                if (a.sigma_only) orow[C - 1] = sig_keep;
                                if (a.sigma_out) a.sigma_out[flat] = sig_keep;
                            }
                        } else {
                            uint32_t r[8];
                            tc_ld8(t_lane + 128, r);
                            tc_wait_ld();
                            if (!last) {      // accumulator half 1 is drained: hand it back before the global stores (their latency
                                              // would otherwise sit in the colour layer's [h1 k01] round via the MEMBAR of a fence)
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(my_aready + 8);
                            }
                            sig_keep = __uint_as_float(r[0]) + __ldg(sigma_w + FN_H);
                            if (valid && a.sigma_only) a.out[flat * C + (C - 1)] = sig_keep;
                            if (valid && a.sigma_out) a.sigma_out[flat] = sig_keep;
                        }
                    } else {
                        uint32_t r[8];
                        tc_ld8(t_lane, r);
                        tc_wait_ld();
                        if (valid) {
                            float c3[3];
#pragma unroll
                            for (int o = 0; o < 3; ++o) {
                                const float x = __uint_as_float(r[o]) + __ldg(rgb_w + 3 * FN_H + o);
                                c3[o] = __fdividef(1.f, 1.f + __expf(-x));
                            }
                            pend_v = make_float4(c3[0], c3[1], c3[2], sig_keep);
                            pend_p = a.out + flat * C + L.label_dim;
                        }
                    }
                }
                tr.log('D', tl, s, 0);
                if (!last && sop.epi != EPI_FILM && sop.epi != EPI_HEAD_TRUNK) {      // (the trunk head hands both halves back itself)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(my_aready); mbar_arrive(my_aready + 8); }
                }
            }
        }
        flush_tail();
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
LoadOp* push(Fast3Args& A) { LoadOp* op = &A.loads[A.n_loads++]; memset(op, 0, sizeof(*op)); op->n_chunks = 1; op->w_is_a = 1; return op; }

// one ring slot of a 256x256 FiLM layer image ([half][k-chunk][16 KB]): SLAB_CHUNKS consecutive k-chunks of one half
void push_film_slab(Fast3Args& A, size_t img_off, int half, int chunk) {
    LoadOp* op = push(A);
    op->src = (uint32_t)(img_off + (size_t)half * 65536 + (size_t)chunk * CHUNK_BYTES);
    op->bytes16 = STAGE_BYTES / 16; op->x_chunk = (uint8_t)chunk; op->n_chunks = SLAB_CHUNKS; op->k0 = 0; op->nk = 4; op->n8 = TILE / 8;
    op->half = (uint8_t)half; op->first = chunk == 0 ? 1 : 0;
}

// the [256 features][64 slots] input image (32 KB, halves 16 KB apart): one load per ring slot
void push_x(Fast3Args& A, size_t img_off, uint8_t kind, int k0, int nk, bool first) {
    for (int h = 0; h < 2 / SLAB_CHUNKS; ++h) {
        LoadOp* op = push(A);
        op->src = (uint32_t)(img_off + (size_t)h * CHUNK_BYTES); op->bytes16 = STAGE_BYTES / 16; op->k0 = (uint8_t)k0; op->nk = (uint8_t)nk;
        op->n8 = TILE / 8; op->half = (uint8_t)h; op->first = first ? 1 : 0; op->xkind = kind;
    }
}

// `half`: which 128-column half of the tile's accumulator takes the [128 points][n] result.  The trunk head uses half 1, so
// that the colour layer's first round ([h0 k01], issued right behind it) does not wait for the head's epilogue.
void push_head(Fast3Args& A, size_t img_off, int img_rows, int n, int half) {
    LoadOp* op = push(A);
    op->src = (uint32_t)img_off;
    op->bytes16 = (uint16_t)((4 * img_rows * FN_KCHUNK * 2) / 16); op->x_chunk = 0; op->n_chunks = 4; op->k0 = 0; op->nk = 4;
    op->n8 = (uint8_t)(n / 8); op->half = (uint8_t)half; op->first = 1; op->w_is_a = 0;
}

bool build_program(const FnLayout& L, Fast3Args& A, bool sigma_only) {
    A.n_loads = 0; A.n_stages = 0;
    auto end_stage = [&](uint8_t epi, uint8_t film, int l0) -> StageOp& {
        StageOp& st = A.stages[A.n_stages++];
        memset(&st, 0, sizeof(st));
        st.epi = epi; st.film = film; st.n_loads = (uint8_t)(A.n_loads - l0); st.l0 = (uint8_t)l0;
        return st;
    };
    {
        int l0 = A.n_loads;
        push_x(A, L.first_img, X_POS, 0, 1, true);
        end_stage(EPI_FILM, 0, l0);
    }
    for (int l = 0; l < L.n_hidden; ++l) {
        if (l == L.trunk_hidden) {
            int l0 = A.n_loads;
            push_head(A, L.head_img, 32, L.label_dim > 0 ? 32 : 8, 1);
            // the head and the first colour layer of a tile are issued back to back: the other tile is in
            // its long epilogue meanwhile, and would otherwise hold the in-order issuer at its own head
            end_stage(EPI_HEAD_TRUNK, 0, l0).fuse_next = sigma_only ? 0 : 1;
            if (sigma_only) return true;      // density only: the program ends at the trunk head
        }
        const bool c0 = (l == L.trunk_hidden);
        int l0 = A.n_loads;
        for (int idx = 0; idx < 8; idx += SLAB_CHUNKS) {     // [h0 c0][h0 c1][h1 c0][h1 c1][h0 c2][h0 c3] (extras) [h1 c2][h1 c3]
            if (c0 && idx == 6) push_x(A, L.color0_ximg, X_EXTRA, 1, L.grid_channels > 0 ? 3 : 1, false);
            push_film_slab(A, L.hid_img[l], (idx >> 1) & 1, (idx >> 2) * 2 + (idx & 1));
        }
        StageOp& st = end_stage(EPI_FILM, (uint8_t)(l + 1), l0);
        st.uniform = 1;
        st.xsync = c0 ? 1 : 0;
        if (A.n_loads > MAX_LOADS - 14 || A.n_stages > MAX_STAGES - 3) return false;
    }
    {
        int l0 = A.n_loads;
        push_head(A, L.rgb_img, 8, 8, 0);
        end_stage(EPI_HEAD_RGB, 0, l0);
    }
    return true;
}

long long* g_trace = nullptr;

}  // namespace

// fenerf_debug_trace: device buffer (4 roles x 4096 int64) that CTA 0 of the next launches logs into, or NULL
void set_fast_trace(long long* buf) { g_trace = buf; }
long long* get_fast_trace() { return g_trace; }

int siren_points_fast3(const FnLayout& L, const unsigned char* packed, const float* points, const float* dirs,
                       const float* film, int batch, long long ppb, int dir_group, int lock_dirs, float* out,
                       long long* trace, int sigma_only, cudaStream_t st, float* sigma_out) {
    static_assert(sizeof(Fast3Args) <= 4000, "kernel parameter block too large");
    static_assert(SMEM_TOTAL <= 232448, "one CTA per SM: 227 KB of shared memory");
    FN_REQUIRE(L.trunk_hidden >= 1 && L.n_hidden - L.trunk_hidden >= 1, "field needs >= 2 trunk and >= 1 colour layers");
    FN_REQUIRE(L.label_dim < 32, "the tcgen05 path packs labels and sigma into one 32-row head (label_dim <= 31)");
    Fast3Args a;
    memset(&a, 0, sizeof(a));
    FN_REQUIRE(build_program(L, a, sigma_only != 0), "field too deep for the stage program");
    a.sigma_only = sigma_only ? 1 : 0;
    a.L = L; a.packed = packed; a.points = points; a.dirs = dirs; a.film = film; a.out = out; a.sigma_out = sigma_out;
    a.ppb = ppb; a.tiles_per_batch = (ppb + TILE - 1) / TILE; a.n_tiles = a.tiles_per_batch * batch;
    a.dir_group = dir_group < 1 ? 1 : dir_group; a.lock_dirs = lock_dirs; a.trace = trace;
#ifdef FENERF_ABLATE
    fprintf(stderr, "fenerf_b200: built with FENERF_ABLATE=%d -- RESULTS ARE WRONG (timing experiment)\n", (int)(FENERF_ABLATE));
#endif
    if (a.n_tiles <= 0) return 0;
    FN_REQUIRE(ppb % a.dir_group == 0, "points_per_batch %lld not a multiple of dir_group %d", ppb, a.dir_group);
    const long long n_pairs = (a.n_tiles + 1) / 2;
    auto kernel = a.trace ? siren_fast3_kernel<true> : siren_fast3_kernel<false>;
    static std::atomic<int> attr_set[2][kMaxDevices];      // per kernel instantiation and device
    FN_CUDA_OK(ensure_dynamic_smem(kernel, attr_set[a.trace ? 1 : 0], (int)SMEM_TOTAL));
    int blocks = (int)(n_pairs < (long long)num_sms() ? n_pairs : (long long)num_sms());
    kernel<<<blocks, NTHREADS, SMEM_TOTAL, st>>>(a);
    FN_LAUNCH_OK("siren_fast3_kernel");
    return 0;
}

}  // namespace fn
