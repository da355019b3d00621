// aw_persistent.cu — KP: the persistent warp-specialised block kernel (64 <= B <= 2048), the default hot path.
//
// One CTA per SM walks over work items (tile, block): a tile = up to T streams (T = 4 up to B = 512, 2 above), a block = one
// of the nb B-frame blocks of the call (RealtimeAudioProcessor.swift:88-116 feeds processPendingBlock once per B frames; a call
// of k*B frames is k blocks, rendered here by ONE launch).  For every item it does what 2*S ConvolutionEngine.process calls plus
// the RealtimeAudioProcessor mix do for T streams (ConvolutionEngine.swift:232-367, RealtimeAudioProcessor.swift:146-163):
// forward real FFT of the T*S overlap-save frames, frequency-domain delay-line multiply-accumulate over all (speaker,
// partition) pairs for both ears, inverse real FFT with the overlap-save discard.  One launch serves every stream range
// ("segment": streams bound to the same bank) of the engine; a tile looks its segment up.
//
//   producer warps      (4; 3 at B >= 512 with T = 4) one elected lane each; producer w fills the ring slots w, w+4, ...: FDL rows
//                       and the matching filter rows go into a shared-memory ring.  A stage = RS consecutive partitions of one
//                       speaker x C bin pairs for the tile's streams, moved by ONE tensor-map TMA copy (cp.async.bulk.tensor.5d,
//                       SASS UTMALDG: box = rows x streams x bins, laid down as [row][stream][bins]; two boxes where the ring
//                       wraps) + one bulk copy for the filter rows, all completing on the slot's `full` mbarrier.  The issue path
//                       is what limits a CTA's supply rate (a copy costs its issuing thread ~155 cycles whatever its size,
//                       tools/tmabw.cu; three producers still match four, two lose 6 %), hence few, large copies.  Without
//                       tensor maps (driver without cuTensorMapEncodeTiled, AW_KP_TENSOR_TMA=0) the FDL rows go by 1-D bulk
//                       copies, one per stream ([stream][row][bins]).  Rows wider than C bin pairs are walked in column chunks
//                       (accumulators stay in registers for the whole chunk).
//   MAC warps           (8; 4 from B = 1024) sets of 128 threads that take alternate stages; a thread owns one bin pair (two at
//                       B >= 512) of RT rows of the stage for ALL T streams and both ears, so a filter value is read from shared
//                       memory once per T streams (shared-memory bandwidth is the resource next to HBM here).  full/empty
//                       mbarriers per ring slot.  The partial sums of the sets (and of the R rows side by side) are reduced
//                       through shared memory in a fixed order (deterministic).
//   FFT warps           (4, 8 or 12) run the forward transforms ONE ITEM AHEAD of the MAC warps (nothing but the head slot of
//                       the FDL depends on them), and the inverse transforms of the item the MAC warps just finished
//                       (accumulators handed over through shared memory, acc_ready/acc_free mbarriers).
//
// The head partition (p = 0) is streamed like any other row: the FFT warps publish it with a generic->async proxy fence
// + a shared-memory count per item parity (release/acquire), which a producer checks before it issues a head row of an item —
// and, for block b > 0 of a call, before any history row (they include head rows written earlier in the same launch).
// The FDL ring of a (stream, speaker) has P + 1 slots: the forward transform of block b+1 (one item ahead) lands in the slot
// block b does not read, so the k blocks of a call need no extra synchronisation — forward(i+2) is only issued after
// inverse(i), i.e. after item i's multiply-accumulate has consumed all its rows.
// Stage order inside an item (and column chunk) — B >= 256: for every speaker its history partitions p = 1..P-1 in groups of RS,
// then the S head rows; B <= 128 (wide stages): for every speaker its partitions p = 0..P-1 in groups of RS — the same for every
// tile size, stream count and blocks-per-call, so a stream's output does not depend on how many streams the engine renders, on
// which GPU it lives, or on how a caller cuts its audio into calls (tested bit-exactly).
// The last (partial) round of tiles is cut into tiles of T/2 or T/4 streams so that every SM gets a share of it.
#include <stdlib.h>
#include <string.h>

#include "aw_fft_blocks.cuh"

namespace aw {

template <int LOG2M, int T> struct PGeo {
    static_assert(T == 2 || T == 4, "tile of 2 or 4 streams");
    static constexpr int M = 1 << LOG2M, halfB = M / 2;
    // Stage shape.  A bulk copy costs its issuing thread ~155 cycles and the copy engine ~40 cycles whatever its size
    // (tools/tmabw.cu), so a stage is built from few, large copies: RS consecutive partitions of one speaker x C bin pairs.
    //   B <= 256 : whole rows (C = B/2), RS = 2 * 128/C rows  -> 4 KB of FDL per stream and copy, 8 KB of filter in one copy
    //   B  = 512 : whole rows (C = 256), one row              -> the same sizes; a MAC thread owns CW = 2 bin pairs
    //   B >= 1024: column chunks of 256 bin pairs, one row    -> 4 KB copies for the FDL and for each filter plane
    static constexpr int C = halfB < 128 ? halfB : (LOG2M >= 9 ? 256 : 128);   // bin pairs per column chunk
    static constexpr int NC = halfB / C;                     // column chunks per row
    static constexpr int CW = C > 128 ? C / 128 : 1;         // bin pairs per MAC thread
    static constexpr int R = C < 128 ? 128 / C : 1;          // rows side by side in a MAC set (one per C threads)
    static constexpr int RT = LOG2M <= 8 ? 2 : 1;            // rows per MAC thread
    static constexpr int RS = R * RT;                        // rows ((speaker, partition) pairs, consecutive partitions) per stage
    // Where does a speaker's head row (p = 0) go?  Wide stages (B <= 128: 8 or 4 rows) would waste a whole stage on it, so there it
    // rides in the speaker's first stage: rows are walked in ring-slot order head, head+1, ..., head+P-1 (consecutive slots),
    // i.e. partitions 0..P-1 (the forward transforms run a whole item ahead, so only a launch's very first stage waits for them).
    // From B = 256 the heads are the last S stages of the item, as late as possible, so that a launch's first item never
    // waits for its forward transforms.
    static constexpr bool MERGED = RS >= 4;
    // sets of 128 MAC threads; set q drains the ring slots q, q + MAC_SETS, ...  From B = 1024 the transforms, not the MAC, set the
    // pace (P is small, 10 transforms of 2B points per stream and block): one MAC set, and the threads go to the FFT warps.
    static constexpr int MAC_SETS = LOG2M >= 10 ? 1 : 2;
    static constexpr int MAC_THREADS = MAC_SETS * 128;
    static constexpr int G = RegFft<LOG2M>::G;               // threads per transform
#ifndef AW_KP_LARGE_FFT_THREADS
#define AW_KP_LARGE_FFT_THREADS 384
#endif
#ifndef AW_KP_LARGE_PRODUCERS
#define AW_KP_LARGE_PRODUCERS 4
#endif
#ifndef AW_KP_LARGE_PREFETCH
#define AW_KP_LARGE_PREFETCH 0
#endif
    static constexpr int FFT_THREADS = LOG2M >= 10 ? AW_KP_LARGE_FFT_THREADS : (8 * G <= 128 ? 128 : 256);
    static constexpr int NFT = FFT_THREADS / G;              // transforms side by side
    // producer warps, one issuing lane each.  B >= 512 with T = 4: three, so that the 19 warps get 104 registers each (the MAC
    // threads hold 2 bin pairs x 4 streams x 2 ears of accumulators); at B = 512 the 6 ring slots divide evenly among them.
#ifndef AW_KP_SMALL_PRODUCERS
#define AW_KP_SMALL_PRODUCERS 4
#endif
    static constexpr int PRODUCERS = LOG2M >= 10 ? AW_KP_LARGE_PRODUCERS : (LOG2M >= 9 && T == 4 ? 3 : (LOG2M <= 8 ? AW_KP_SMALL_PRODUCERS : 4));
    static constexpr int THREADS = 32 * PRODUCERS + MAC_THREADS + FFT_THREADS;
    // Registers per thread.  The CTA owns the SM, but up to B = 256 it leaves 8 K registers (and 40 KB of shared memory) free:
    // room for one 128-thread CTA of the equalizer kernel, whose float64 recurrence then runs next to the next call's convolution.
#ifndef AW_KP_SMALL_MAXNREG
#define AW_KP_SMALL_MAXNREG 112
#endif
    static constexpr int MAXNREG = LOG2M <= 8 ? AW_KP_SMALL_MAXNREG : 96;
    static_assert(THREADS * MAXNREG <= 65536, "register file");
    static constexpr int PS = PaddedSize<LOG2M>::value;
    static constexpr int stage_f4 = RS * (T + 2) * C;        // FDL [T][RS][C] + filter [RS][2 planes][C] float4
    static constexpr size_t stage_bytes = (size_t)stage_f4 * sizeof(float4);
    static constexpr int red_f4 = R > 1 ? (MAC_SETS * R - 1) * T * 2 * C : 0;   // partial sums of every (set, row) but the first
                                                                                 // (R == 1: exchanged through the accumulator buffers)
    static constexpr int TW = pt_total_entries(LOG2M);       // per-pass twiddle tables (conflict-free reads)
    static constexpr size_t fixed_bytes = (size_t)TW * sizeof(float2)                   // twiddles
                                          + (size_t)(NFT + 2 * T) * PS * sizeof(float2)  // forward buffers + accumulator buffers
                                          + (size_t)red_f4 * sizeof(float4) + (size_t)FFT_THREADS * sizeof(float) + 1024;
    static constexpr int max_stages = (int)((226 * 1024 - fixed_bytes) / stage_bytes);
    // A ring slot is always filled by the same producer warp (slot index mod PRODUCERS) and drained by the same MAC set (slot
    // index mod MAC_SETS): mbarrier parity waits are only safe for a waiter that is at most one phase behind.  Even depth keeps
    // two sets alternating across the wrap.
    static constexpr int STAGES = (max_stages > 32 ? 32 : max_stages) / 2 * 2;
    static constexpr size_t smem = fixed_bytes + (size_t)STAGES * stage_bytes;
    // next round's operands fetched while this round transforms — only where the registers exist (with 96-104 registers per
    // thread at B >= 512 the spills cost more than the exposed latency: measured)
    static constexpr bool PREFETCH = LOG2M <= 8 || (LOG2M >= 10 && AW_KP_LARGE_PREFETCH);
    static_assert(STAGES >= PRODUCERS && STAGES % MAC_SETS == 0 && (MAC_SETS == 2 || R == 1), "ring geometry");
    static_assert(NC == 1 || RS == 1, "column chunks carry one row per stage");
};

// Timing experiments (skip the transforms, their traffic or their arithmetic) exist only in builds made with
// -DAW_TIMING_EXPERIMENTS; the shipped kernel cannot be told to skip work.
#ifdef AW_TIMING_EXPERIMENTS
#define AW_DBG(a) ((a).debug)
#else
#define AW_DBG(a) 0
#endif

// One launch serves up to kKpMaxSegments stream ranges ("segments": streams bound to the same bank — per-device profiles,
// DeviceProfileManager.swift:4-12).  Tiles are numbered across the segments; a tile never straddles two of them.
struct PersistArgs {
    int n_segs, n_tiles;
    int small;                   // streams per tile behind a segment's first n_big tiles (T, T/2 or T/4)
    int nb, order, keep_pct, prev_is_rows;   // KpCall
    int Se, P_cap;               // engine-wide layout of the state arrays (speakers per stream, FDL slots per (stream, speaker))
    KpSegment seg[kKpMaxSegments];
    StridedIn cur, prev;         // block b of the call is cur + b*B; the block before block 0 is prev (inputOverlapBuffer)
    float *overlap_save;
    float2 *fdl;
    float *fdl_ny;
    const unsigned char *tmaps;  // CUtensorMap[rows - 1][3 tile sizes] of the FDL (see aw_kernels.h), or nullptr: 1-D bulk copies per stream
    StridedOut out;
    const float2 *tw;
    int debug;   // AW_TIMING_EXPERIMENTS builds only: bit 0 skips the forward transforms, bit 1 the inverse transforms, bit 2 keeps the
                 // forward transforms' arithmetic but not their memory traffic, bit 3 their traffic but not their arithmetic
    EqFuse eq;   // steady-state equalizer applied to the block before it is stored (n_filters == 0: none); nb == 1 only
};

struct TileCtx {                 // what a role needs to know about the tile it works on
    int s0, nvalid, ts;          // first stream, valid streams, streams of this tile's shape (T, or `small` in the cut last round)
    int S, P, Pm, head, hs;      // renderers, partitions, ring modulus, FDL head slot of block 0, stages per speaker
    const float4 *bank;
    const float *bank_ny;
    const KpRowTable *rt;        // which input channels a row sums, whose filter rows it uses
};

template <int T, int RS, bool MERGED>
__device__ __forceinline__ TileCtx tile_ctx(const PersistArgs &a, int tile)
{
    int i = 0;
    while (i + 1 < a.n_segs && tile >= a.seg[i + 1].tile0) ++i;
    const KpSegment &d = a.seg[i];
    TileCtx c;
    const int k = tile - d.tile0;
    int size = T;
    if (k < d.n_big) c.s0 = d.first_stream + k * T;
    else { c.s0 = d.first_stream + d.n_big * T + (k - d.n_big) * a.small; size = a.small; }
    c.nvalid = min(size, d.first_stream + d.n_streams - c.s0);
    c.ts = size;
    c.S = d.S; c.P = d.P; c.Pm = d.Pm; c.head = d.head;
    c.hs = MERGED ? (d.P + RS - 1) / RS : (d.P - 1 + RS - 1) / RS;   // stages per speaker (MERGED: head row included)
    c.bank = d.bank; c.bank_ny = d.bank_ny; c.rt = d.rows;
    return c;
}

template <int LOG2M, int T>
__global__ void __maxnreg__((PGeo<LOG2M, T>::MAXNREG)) k_persistent(const __grid_constant__ PersistArgs a)
{
    using PG = PGeo<LOG2M, T>;
    constexpr int M = PG::M, halfB = PG::halfB, C = PG::C, NC = PG::NC, R = PG::R, RT = PG::RT, RS = PG::RS, CW = PG::CW, G = PG::G, NFT = PG::NFT;
    constexpr int STAGES = PG::STAGES, PS = PG::PS, stage_f4 = PG::stage_f4, PRODUCERS = PG::PRODUCERS;
    constexpr int MAC_WARPS = PG::MAC_THREADS / 32, SET_WARPS = 4, FFT_WARPS = PG::FFT_THREADS / 32;
    constexpr int BAR_RED_A = 1, BAR_RED_B = 2, BAR_EQ = 3, BAR_FFT0 = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);
    float2 *tw = reinterpret_cast<float2 *>(smem_raw + (size_t)STAGES * PG::stage_bytes);
    float2 *fftbuf = tw + PG::TW;
    float2 *accbuf = fftbuf + (size_t)NFT * PS;
    float4 *red = reinterpret_cast<float4 *>(accbuf + (size_t)2 * T * PS);
    float *part = reinterpret_cast<float *>(red + PG::red_f4);
    uint64_t *full = reinterpret_cast<uint64_t *>(part + PG::FFT_THREADS);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_ready = empty + STAGES;
    uint64_t *acc_free = acc_ready + 1;
    // FFT warps that have published their head rows, counted separately for even and odd items: a warp without work in an item
    // runs ahead of the others — by at most one item (it waits for acc_ready before the next forward pass), so with two
    // counters its early arrival for item i+1 cannot be mistaken for another warp's arrival for item i
    unsigned *heads_done = reinterpret_cast<unsigned *>(acc_free + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nb = a.nb, n_items = my_tiles * nb;
    const int dbg = AW_DBG(a);
    constexpr bool MERGED = PG::MERGED;
    auto my_tile = [&](int lt) { return tile_ctx<T, RS, MERGED>(a, (int)blockIdx.x + lt * (int)gridDim.x); };
    // items of this CTA in walk order: (local tile lt, block b)
    auto next_item = [&](int &lt, int &b) {
        if (a.order) { if (++b == nb) { b = 0; ++lt; } }
        else if (++lt == my_tiles) { lt = 0; ++b; }
    };
    auto head_of = [](const TileCtx &tc, int b) {                // fdlIndex of block b: it moves down one slot per block (:256-259)
        int h = tc.head - (b < tc.Pm ? b : b % tc.Pm);
        return h < 0 ? h + tc.Pm : h;
    };

    for (int k = tid; k < PG::TW; k += PG::THREADS) tw[k] = RegFft<LOG2M>::pt_entry(a.tw, k);
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], SET_WARPS); }
        heads_done[0] = 0;
        heads_done[1] = 0;
        mbar_init(acc_ready, MAC_WARPS);
        mbar_init(acc_free, FFT_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // Programmatic dependent launch: the next call's launch may be scheduled onto SMs as this grid's CTAs retire (its prologue
    // above touches nothing a previous launch writes), and this grid goes no further until the previous one has completed and
    // flushed — the FDL head slots, the overlap buffer and the output it wrote are read/overwritten below.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < PRODUCERS) {
        // ===== producers: warp w fills the ring slots w, w + PRODUCERS, ... every time the stage sequence comes round to them =====
        if (lane == 0 && n_items > 0) {
            const size_t stream_stride = (size_t)a.Se * a.P_cap * halfB;
            const float4 *fdl4 = reinterpret_cast<const float4 *>(a.fdl);
            // L2 priorities: history rows are read once per block (evict first) — they must not push out the head rows the FFT
            // warps have just written, nor the filter bank every tile re-reads (evict last).  Tile-major calls re-read a tile's
            // rows in its next block: there a share of them (keep_pct) is loaded evict_last in all but the call's last block.
            const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
            const uint64_t pol_reuse = l2_policy_evict_last_share(a.keep_pct);
            // position of this producer in the stage sequence, advanced one stage at a time (no divisions on the issue path):
            // item (lt, b), column chunk c, and inside the chunk either history group jj of speaker s or the head row of speaker s
            int item = 0, lt = 0, b = 0, c = 0, s = 0, jj = 0, stage = 0;
            unsigned phase = 0, seen0 = 0, seen1 = 0;        // last values read from heads_done[0], [1]
            TileCtx tc = my_tile(0);
            int hb = head_of(tc, 0);
            int bs = tc.rt->spk[0];                          // bank speaker whose filter rows row s uses
            bool hist = MERGED || tc.hs > 0;
            auto new_item = [&]() {
                const int lt0 = lt;
                ++item;
                next_item(lt, b);
                if (item < n_items) {
                    if (lt != lt0) tc = my_tile(lt);
                    hb = head_of(tc, b);
                }
            };
            auto advance = [&]() {
                if (MERGED) {
                    if (++jj == tc.hs) {
                        jj = 0;
                        if (++s == tc.S) { s = 0; if (++c == NC) { c = 0; new_item(); } }
                    }
                } else if (hist) {
                    if (++jj == tc.hs) { jj = 0; if (++s == tc.S) { s = 0; hist = false; } }
                } else if (++s == tc.S) {
                    s = 0;
                    if (++c == NC) { c = 0; new_item(); }
                    hist = tc.hs > 0;
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            };
            auto advance_to_my_slot = [&](int steps) {
                const int s0_ = s, item0 = item;
                for (int i = 0; i < steps && item < n_items; ++i) advance();
                if (item < n_items && (s != s0_ || item != item0)) bs = tc.rt->spk[s];
            };
            advance_to_my_slot(warp);   // producer w owns the ring slots w, w + PRODUCERS, ... < STAGES
            while (item < n_items) {
                // rows of this stage: partitions p0, p0+1, ... in consecutive ring slots (MERGED: the first stage of a speaker starts
                // with the head row, p = 0)
                const int p0 = MERGED ? jj * RS : (hist ? 1 + jj * RS : 0);
                const int nrows = MERGED ? min(RS, tc.P - p0) : (hist ? min(RS, tc.P - p0) : 1);
                const bool has_head = MERGED ? (jj == 0) : !hist;
                // Head rows of an item exist once every FFT warp has published them (a monotonic count: no phase to alias).  The
                // history rows of block b > 0 include the head rows this launch wrote for blocks < b of the same tile: the newest
                // of them belongs to the tile's previous item — and this producer may be a whole ring ahead of the producer
                // that waited for it, so it checks for itself.
                const int dep = has_head ? item : (b > 0 ? (a.order ? item - 1 : item - my_tiles) : -1);
                const unsigned need = dep < 0 ? 0u : (unsigned)(FFT_WARPS * ((dep >> 1) + 1));
                if (dep >= 0 && ((dep & 1) ? seen1 : seen0) < need) {
                    unsigned got;
                    for (;;) {
                        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(got) : "r"(smem_u32(heads_done + (dep & 1))) : "memory");
                        if (got >= need) break;
                        __nanosleep(128);
                    }
                    if (dep & 1) seen1 = got; else seen0 = got;
                    // the acquire above is a generic-proxy operation; the bulk copies below read through the async proxy: order them
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                int slot = hb + p0;
                if (slot >= tc.Pm) slot -= tc.Pm;            // the ring has Pm = P + 1 slots (Q4: the reference's has P)
                const int n1 = min(nrows, tc.Pm - slot);     // rows before the ring wraps
                mbar_wait(&empty[stage], phase ^ 1u);
                float4 *dst = ring + (size_t)stage * stage_f4;
                const uint64_t pol = has_head ? pol_keep : ((a.order && b + 1 < nb) ? pol_reuse : pol_stream);
                if (a.tmaps) {
                    // one tensor-map copy for the rows of all the tile's streams (two where the ring wraps): [row][stream][bins]
                    mbar_expect_tx(&full[stage], (unsigned)(nrows * (tc.ts + 2) * C * sizeof(float4)));
                    const int sidx = tc.ts == 4 ? 0 : (tc.ts == 2 ? 1 : 2);
                    constexpr int chunks = 2 * C > 256 ? 2 * C / 256 : 1;     // 256-bin chunks per stage column
                    tma_load_5d_hint(dst, a.tmaps + (size_t)((n1 - 1) * 3 + sidx) * 128, 0, c * chunks, tc.s0, slot, s, &full[stage], pol);
                    if (RS > 1 && n1 < nrows)
                        tma_load_5d_hint(dst + (size_t)n1 * tc.ts * C, a.tmaps + (size_t)((nrows - n1 - 1) * 3 + sidx) * 128, 0, c * chunks, tc.s0, 0,
                                         s, &full[stage], pol);
                } else {
                    mbar_expect_tx(&full[stage], (unsigned)(nrows * (tc.nvalid + 2) * C * sizeof(float4)));
#pragma unroll
                    for (int u = 0; u < T; ++u) {
                        if (u < tc.nvalid) {                 // a partial tile moves (and waits for) only the rows it has
                            const float4 *row = fdl4 + (size_t)(tc.s0 + u) * stream_stride + (size_t)s * a.P_cap * halfB + c * C;
                            bulk_g2s_hint(dst + (u * RS) * C, row + (size_t)slot * halfB, (unsigned)(n1 * C * sizeof(float4)), &full[stage], pol);
                            if (RS > 1 && n1 < nrows)
                                bulk_g2s_hint(dst + (u * RS + n1) * C, row, (unsigned)((nrows - n1) * C * sizeof(float4)), &full[stage], pol);
                        }
                    }
                }
                const float4 *frow = tc.bank + ((size_t)bs * tc.P + p0) * M;
                if (NC == 1) {                               // whole rows: both planes of RS consecutive partitions are contiguous
                    bulk_g2s_hint(dst + T * RS * C, frow, (unsigned)(nrows * 2 * C * sizeof(float4)), &full[stage], pol_keep);
                } else {
                    bulk_g2s_hint(dst + T * C, frow + c * C, (unsigned)(C * sizeof(float4)), &full[stage], pol_keep);
                    bulk_g2s_hint(dst + T * C + C, frow + halfB + c * C, (unsigned)(C * sizeof(float4)), &full[stage], pol_keep);
                }
                const int step = stage + PRODUCERS < STAGES ? PRODUCERS : STAGES - stage + warp;   // to my next slot
                advance_to_my_slot(step);
            }
        }
    } else if (warp < PRODUCERS + MAC_WARPS) {
        // ===== MAC warps: set q consumes the stages k = q (mod 2); a thread owns one bin pair of one row, all T streams =====
        const int mt = tid - 32 * PRODUCERS;
        const int set = mt >> 7, w = mt & 127;
        const int r = C < 128 ? w / C : 0, jp = C < 128 ? w - r * C : w;
        // set q drains the ring slots of parity q, i.e. the stages k = q, q + 2, ... of the CTA's stage sequence (STAGES is even)
        int stage = set;                                     // ring slot of my next stage
        unsigned phase = 0;
        int m = set;                                         // my next stage, relative to the current chunk
        constexpr int SETS = PG::MAC_SETS;
        int lt = 0, b = 0;
        TileCtx tc = my_tiles > 0 ? my_tile(0) : TileCtx{};
        for (int item = 0; item < n_items; ++item) {
            // first head stage of / stages per column chunk (MERGED: no separate head stages)
            const int hs = tc.hs, head0 = tc.S * hs, spc = MERGED ? head0 : head0 + tc.S;
            // where stream u's copy of row r sits in a stage: tensor-map loads give [row][stream], bulk copies [stream][row]
            const int x_us = a.tmaps ? C : RS * C, x_rs = a.tmaps ? tc.ts * C : C;
            for (int c = 0; c < NC; ++c) {
                int jj = hs > 0 ? m % hs : 0;                // history group of my next stage (RS > 1 only)
                // Partial sums are reduced in the order (stage parity within the item, row): with an odd number of stages per item
                // (7 rows) the two sets swap the even and the odd stages from item to item, and the order of the sum must not
                const int contributor = (PG::MAC_SETS == 2 ? (m & 1) : 0) * R + r;
                float4 aL[CW][T], aR[CW][T];
#pragma unroll
                for (int v = 0; v < CW; ++v)
#pragma unroll
                    for (int u = 0; u < T; ++u) { aL[v][u] = make_float4(0.f, 0.f, 0.f, 0.f); aR[v][u] = aL[v][u]; }
                for (; m < spc; m += SETS) {
                    int nrows = 1;                           // head stages carry one row
                    if (MERGED) nrows = min(RS, tc.P - jj * RS);
                    else if (RS > 1 && m < head0) nrows = min(RS, tc.P - 1 - jj * RS);
                    mbar_wait(&full[stage], phase);
                    const float4 *src = ring + stage * stage_f4;
#pragma unroll
                    for (int i = 0; i < RT; ++i) {
                        const int row = r + R * i;
                        if (row < nrows) {
#pragma unroll
                            for (int v = 0; v < CW; ++v) {
                                const int col = jp + 128 * v;
                                const float4 h0 = src[T * RS * C + (row * 2) * C + col], h1 = src[T * RS * C + (row * 2 + 1) * C + col];
                                float4 x[T];
#pragma unroll
                                for (int u = 0; u < T; ++u) x[u] = src[u * x_us + row * x_rs + col];
#pragma unroll
                                for (int u = 0; u < T; ++u) {
                                    cmac2f(aL[v][u], x[u], h0.x, h0.y, h1.x, h1.y);
                                    cmac2f(aR[v][u], x[u], h0.z, h0.w, h1.z, h1.w);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    stage += SETS;
                    if (stage >= STAGES) { stage -= STAGES; phase ^= 1u; }
                    if (RS > 1 && hs > 0) { jj += SETS; while (jj >= hs) jj -= hs; }
                }
                m -= spc;                                    // position in the next chunk
                // the FFT warps must be done with the previous item's accumulators before they are overwritten
                if (c == 0 && item > 0) mbar_wait(acc_free, (unsigned)((item - 1) & 1));
                if constexpr (R == 1 && SETS == 1) {
#pragma unroll
                    for (int v = 0; v < CW; ++v) {
                        const int J = c * C + jp + 128 * v;          // bins 2J, 2J+1
#pragma unroll
                        for (int u = 0; u < T; ++u) {
                            float2 *bl = accbuf + (size_t)(2 * u) * PS, *br = bl + PS;
                            bl[pad16(2 * J)] = make_float2(aL[v][u].x, aL[v][u].y);
                            bl[pad16(2 * J + 1)] = make_float2(aL[v][u].z, aL[v][u].w);
                            br[pad16(2 * J)] = make_float2(aR[v][u].x, aR[v][u].y);
                            br[pad16(2 * J + 1)] = make_float2(aR[v][u].z, aR[v][u].w);
                        }
                    }
                } else if constexpr (R == 1) {
                    // set 1 hands its partial sums over in the accumulator buffers; set 0 adds its own and leaves the result there
                    if (set == 1) {
#pragma unroll
                        for (int v = 0; v < CW; ++v) {
                            const int J = c * C + jp + 128 * v;      // bins 2J, 2J+1
#pragma unroll
                            for (int u = 0; u < T; ++u) {
                                float2 *bl = accbuf + (size_t)(2 * u) * PS, *br = bl + PS;
                                bl[pad16(2 * J)] = make_float2(aL[v][u].x, aL[v][u].y);
                                bl[pad16(2 * J + 1)] = make_float2(aL[v][u].z, aL[v][u].w);
                                br[pad16(2 * J)] = make_float2(aR[v][u].x, aR[v][u].y);
                                br[pad16(2 * J + 1)] = make_float2(aR[v][u].z, aR[v][u].w);
                            }
                        }
                    }
                    named_sync(BAR_RED_A, PG::MAC_THREADS);
                    if (set == 0) {
#pragma unroll
                        for (int v = 0; v < CW; ++v) {
                            const int J = c * C + jp + 128 * v;
#pragma unroll
                            for (int u = 0; u < T; ++u) {
                                float2 *bl = accbuf + (size_t)(2 * u) * PS, *br = bl + PS;
                                const float2 l0 = bl[pad16(2 * J)], l1 = bl[pad16(2 * J + 1)], r0 = br[pad16(2 * J)], r1 = br[pad16(2 * J + 1)];
                                bl[pad16(2 * J)] = make_float2(aL[v][u].x + l0.x, aL[v][u].y + l0.y);
                                bl[pad16(2 * J + 1)] = make_float2(aL[v][u].z + l1.x, aL[v][u].w + l1.y);
                                br[pad16(2 * J)] = make_float2(aR[v][u].x + r0.x, aR[v][u].y + r0.y);
                                br[pad16(2 * J + 1)] = make_float2(aR[v][u].z + r1.x, aR[v][u].w + r1.y);
                            }
                        }
                    }
                } else {
                    static_assert(R == 1 || CW == 1, "narrow rows: one bin pair per thread");
                    const int J = c * C + jp;                // bins 2J, 2J+1
                    if (contributor > 0) {
#pragma unroll
                        for (int u = 0; u < T; ++u) {
                            float4 *d = red + ((size_t)((contributor - 1) * T + u) * 2) * C + jp;
                            d[0] = aL[0][u];
                            d[C] = aR[0][u];
                        }
                    }
                    named_sync(BAR_RED_A, PG::MAC_THREADS);
                    if (contributor == 0) {
#pragma unroll
                        for (int q = 1; q < PG::MAC_SETS * R; ++q) {
#pragma unroll
                            for (int u = 0; u < T; ++u) {
                                const float4 *d = red + ((size_t)((q - 1) * T + u) * 2) * C + jp;
                                const float4 l = d[0], rt = d[C];
                                aL[0][u].x += l.x; aL[0][u].y += l.y; aL[0][u].z += l.z; aL[0][u].w += l.w;
                                aR[0][u].x += rt.x; aR[0][u].y += rt.y; aR[0][u].z += rt.z; aR[0][u].w += rt.w;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < T; ++u) {
                            float2 *bl = accbuf + (size_t)(2 * u) * PS, *br = bl + PS;
                            bl[pad16(2 * J)] = make_float2(aL[0][u].x, aL[0][u].y);
                            bl[pad16(2 * J + 1)] = make_float2(aL[0][u].z, aL[0][u].w);
                            br[pad16(2 * J)] = make_float2(aR[0][u].x, aR[0][u].y);
                            br[pad16(2 * J + 1)] = make_float2(aR[0][u].z, aR[0][u].w);
                        }
                    }
                    named_sync(BAR_RED_B, PG::MAC_THREADS);  // the partial sums may be overwritten by the next chunk
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_ready);
            const int lt0 = lt;
            next_item(lt, b);
            if (item + 1 < n_items && lt != lt0) tc = my_tile(lt);
        }
    } else {
        // ===== FFT warps =====
        using F = RegFft<LOG2M>;
        const int ft = tid - 32 * PRODUCERS - PG::MAC_THREADS;
        const int f = ft / G, t = ft - f * G;
        const int fwarp = ft >> 5;
        const int warp_first_f = G >= 32 ? f : (ft - lane) / G;   // first transform handled by this warp
        const GroupBar gb{BAR_FFT0 + f, G};

        // publish the head slots of an item: generic-proxy global writes -> visible to the producer's async-proxy reads.  The fence
        // waits for this thread's outstanding stores; from B = 256 (heads are an item's LAST stages) it is issued a little later,
        // inside the inverse pass of the previous item, when most of them have landed.
        auto publish_heads = [&](int item) {
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(heads_done + (item & 1))) : "memory");
        };

        auto forward_item = [&](int lt, int b, int item) {
            const TileCtx tc = my_tile(lt);
            const int s0 = tc.s0, nvalid = tc.nvalid;
            const int nfft = T * tc.S;
            const int hb = head_of(tc, b);
            // overlap-save frame of block b = [block b-1 | block b] (:237-248); block -1 is the engine's inputOverlapBuffer
            const StridedIn cur_b{a.cur.ptr + (size_t)b * M, a.cur.ss, a.cur.cs};
            const StridedIn prev_b = b == 0 ? a.prev : StridedIn{a.cur.ptr + (size_t)(b - 1) * M, a.cur.ss, a.cur.cs};
            float *ov_save = (b == nb - 1) ? a.overlap_save : nullptr;   // inputOverlapBuffer <- the call's last block (:243)
            const bool prev_rows = b == 0 && a.prev_is_rows;             // the overlap buffer holds one (summed) block per row
            if (!(dbg & 1)) {
                auto fetch = [&](int base, float2 (&v)[F::E], bool &active, int &stream, int &s) {
                    const int idx = base + f;
                    const int ls = idx / tc.S;
                    s = idx - ls * tc.S;
                    active = idx < nfft && ls < nvalid;
                    stream = s0 + (active ? ls : 0);
                    // a row sums the input channels that share its filter pair (KpRowTable); the overlap buffer holds the row's
                    // previous block already summed, the call's own input does not
                    const signed char *src = tc.rt->src[active ? s : 0];
                    const int c0 = src[0];
                    const float *prev = prev_b.ptr + stream * prev_b.ss + (prev_rows ? s : c0) * prev_b.cs;
                    const float *cur = cur_b.ptr + stream * cur_b.ss + c0 * cur_b.cs;
#pragma unroll
                    for (int e = 0; e < F::E; ++e) {
                        const int i = F::template load_index<0>(t, e);   // frame = [previous block | current block] (:237-248)
                        v[e] = (!active || (dbg & (4 | 32))) ? make_float2(0.f, 0.f)
                                       : (i < M / 2 ? *reinterpret_cast<const float2 *>(prev + 2 * i) : *reinterpret_cast<const float2 *>(cur + 2 * (i - M / 2)));
                    }
                    if (active && src[1] >= 0) {                         // warp-uniform per transform: rare (FC + LFE)
                        for (int q = 1; q < kKpMaxRowSources && src[q] >= 0; ++q) {
                            const float *pq = prev_b.ptr + stream * prev_b.ss + src[q] * prev_b.cs;
                            const float *cq = cur_b.ptr + stream * cur_b.ss + src[q] * cur_b.cs;
#pragma unroll
                            for (int e = 0; e < F::E; ++e) {
                                const int i = F::template load_index<0>(t, e);
                                if (i < M / 2) {
                                    if (!prev_rows) { const float2 x = *reinterpret_cast<const float2 *>(pq + 2 * i); v[e].x += x.x; v[e].y += x.y; }
                                } else {
                                    const float2 x = *reinterpret_cast<const float2 *>(cq + 2 * (i - M / 2));
                                    v[e].x += x.x; v[e].y += x.y;
                                }
                            }
                        }
                    }
                };
                auto transform = [&](float2 (&v)[F::E], bool active, int stream, int sp) {
                    if (dbg & 8) {                       // timing experiment: the forward transform's memory traffic without its arithmetic
                        if (active) {
                            float2 *dst = a.fdl + (((size_t)stream * a.Se + sp) * a.P_cap + hb) * M;
#pragma unroll
                            for (int e = 0; e < F::E; ++e) {
                                const int i = F::template load_index<0>(t, e);
                                dst[i] = v[e];
                                if (ov_save && i >= M / 2) *reinterpret_cast<float2 *>(ov_save + ((size_t)stream * a.Se + sp) * M + 2 * (i - M / 2)) = v[e];
                            }
                        }
                        return;
                    }
                    if (dbg & (4 | 16)) active = false;       // timing experiments: 4 = arithmetic without traffic, 16 = loads but no stores,
                                                                // 32 = stores but no loads
                    if (active && ov_save) {
                        float *ov = ov_save + ((size_t)stream * a.Se + sp) * M;
#pragma unroll
                        for (int e = 0; e < F::E; ++e) {
                            const int i = F::template load_index<0>(t, e);
                            if (i >= M / 2) *reinterpret_cast<float2 *>(ov + 2 * (i - M / 2)) = v[e];
                        }
                    }
                    const size_t row = ((size_t)stream * a.Se + sp) * a.P_cap + hb;
                    float2 *dst = a.fdl + row * M;
                    float *dst_ny = a.fdl_ny + row;
                    forward_frame_regs<LOG2M, true>(fftbuf + (size_t)f * PS, tw, t, active, v,
                                              [&](int k, float2 x) { dst[k] = x; }, [&](float ny) { *dst_ny = ny; }, gb);   // FDL[head] <- spectrum (:256-264)
                };
                float2 v[F::E];
                bool active = false;
                int stream = 0, sp = 0;
                if constexpr (PG::PREFETCH) {
                    float2 vn[F::E];
                    bool active_n = false;
                    int stream_n = 0, sp_n = 0;
                    fetch(0, v, active, stream, sp);
                    for (int base = 0; base < nfft; base += NFT) {
                        if (base + NFT < nfft) fetch(base + NFT, vn, active_n, stream_n, sp_n);
                        transform(v, active, stream, sp);
#pragma unroll
                        for (int e = 0; e < F::E; ++e) v[e] = vn[e];
                        active = active_n; stream = stream_n; sp = sp_n;
                    }
                } else {
                    // no registers for a second frame: ask L2 for the next round's frame while this one is transformed (one request
                    // per 128-byte line: 16 consecutive threads of a transform cover one)
                    auto prefetch = [&](int base) {
                        const int idx = base + f;
                        const int ls = idx / tc.S, s = idx - ls * tc.S;
                        if (idx < nfft && ls < nvalid && (t & 15) == 0) {
                            const int c0 = tc.rt->src[s][0];
                            const float *prev = prev_b.ptr + (s0 + ls) * prev_b.ss + (prev_rows ? s : c0) * prev_b.cs;
                            const float *cur = cur_b.ptr + (s0 + ls) * cur_b.ss + c0 * cur_b.cs;
#pragma unroll
                            for (int e = 0; e < F::E; ++e) {
                                const int i = F::template load_index<0>(t, e);
                                const float *p = i < M / 2 ? prev + 2 * i : cur + 2 * (i - M / 2);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                            }
                        }
                    };
                    for (int base = 0; base < nfft; base += NFT) {
                        fetch(base, v, active, stream, sp);
                        if (base + NFT < nfft) prefetch(base + NFT);
                        transform(v, active, stream, sp);
                    }
                }
            }
            if (MERGED || item == 0) publish_heads(item);       // wide stages need the head rows first; otherwise see inverse_item
        };

        // fused equalizer (ParametricEqualizerState.process, ParametricEqualizerProcessor.swift:58-91): a systolic array across the
        // lanes of a warp, lane f of a group of n_filters lanes owning biquad f of one (stream, ear) channel — see k_eq_systolic
        const int gw = a.eq.n_filters;
        const int eq_groups = gw > 0 ? 32 / gw : 0;
        const int eg = gw > 0 ? lane / gw : 0, ef = gw > 0 ? lane - eg * gw : 0;
        double eb0 = 0, eb1 = 0, eb2 = 0, ea1 = 0, ea2 = 0, epre = 1.0;
        if (gw > 0 && eg < eq_groups) {
            const double *cf = a.eq.prog->coef[ef];
            eb0 = cf[0]; eb1 = cf[1]; eb2 = cf[2]; ea1 = cf[3]; ea2 = cf[4];
            epre = a.eq.prog->preamp_linear;
        }

        auto inverse_item = [&](int lt, int b, int item) {
            const TileCtx tc = my_tile(lt);
            const int s0 = tc.s0, nvalid = tc.nvalid;
            BlockGeom g;                                         // the Nyquist sum's view of this tile's segment
            g.S = tc.S; g.Se = a.Se; g.P = tc.P; g.Pm = tc.Pm; g.P_cap = a.P_cap; g.head = head_of(tc, b);
            // The Nyquist products of the item's 2T outputs (S*P each) are summed by whole warps, outputs dealt round-robin to the FFT
            // warps, and handed over through `part` (double-buffered by item parity).  The barrier also orders every read of
            // the Nyquist side array before the forward transforms two items ahead, which overwrite this block's oldest slot.
            float *ny_part = part + (item & 1) * 2 * T;
            for (int idx = fwarp; idx < 2 * T; idx += FFT_WARPS) {
                const int ls = idx >> 1, ear = idx & 1;
                float sum = 0.f;
                if (ls < nvalid) {
                    const float *xs = a.fdl_ny + (size_t)(s0 + ls) * a.Se * a.P_cap;
                    const int n = tc.S * tc.P;
                    int sp = 0, p = lane;                        // k = sp * P + p, k = lane, lane + 32, ...
                    while (p >= tc.P) { p -= tc.P; ++sp; }
                    for (int k = lane; k < n; k += 32) {
                        int slot = g.head + p;
                        if (slot >= tc.Pm) slot -= tc.Pm;
                        sum = fmaf(xs[(size_t)sp * a.P_cap + slot], tc.bank_ny[((size_t)tc.rt->spk[sp] * tc.P + p) * 2 + ear], sum);
                        p += 32;
                        while (p >= tc.P) { p -= tc.P; ++sp; }
                    }
                }
                sum = group_sum(sum, 32);
                if (lane == 0) ny_part[idx] = sum;
            }
            named_sync(BAR_EQ, PG::FFT_THREADS);
            if (!MERGED && item + 1 < n_items) publish_heads(item + 1);   // the forward pass of the next item was issued before this call
            for (int base = 0; base < 2 * T; base += NFT) {
                if (base + warp_first_f >= 2 * T) continue;          // warp-uniform: no transform of this warp has work
                const int idx = base + f;
                const int ls = idx >> 1, ear = idx & 1;
                const bool active = idx < 2 * T && ls < nvalid;
                const int stream = s0 + (active ? ls : 0);
                // idle transforms of a working warp run on their (free) forward buffer so the warp stays converged
                float2 *buf = idx < 2 * T ? accbuf + (size_t)idx * PS : fftbuf + (size_t)f * PS;
                const float ny = idx < 2 * T ? ny_part[idx] : 0.f;
                if (gw == 0) {
                    float *row = a.out.ptr + stream * a.out.ss + ear * a.out.cs + (size_t)b * M;
                    inverse_frame<LOG2M, false, true>(buf, ny, tw, t, active, [&](int i, float x0, float x1) { store_pair(a.out, row, 2 * i, x0, x1); }, gb);
                } else {                                             // keep the block in shared memory (over the spectrum) for the EQ
                    float *eb = reinterpret_cast<float *>(buf);
                    inverse_frame<LOG2M, true, true>(buf, ny, tw, t, active, [&](int i, float x0, float x1) { eb[2 * i] = x0; eb[2 * i + 1] = x1; }, gb);
                }
            }
            if (gw > 0) {
                named_sync(BAR_EQ, PG::FFT_THREADS);
                for (int ch0 = fwarp * eq_groups; ch0 < 2 * T; ch0 += FFT_WARPS * eq_groups) {   // warp-uniform
                    const int ch = ch0 + eg;
                    const bool live = eg < eq_groups && ch < 2 * T && (ch >> 1) < nvalid;
                    float *eb = reinterpret_cast<float *>(accbuf + (size_t)(live ? ch : 0) * PS);
                    double *zp = a.eq.z + ((((size_t)(s0 + (live ? ch >> 1 : 0)) * 2 + a.eq.voice) * 2 + (ch & 1)) * 64 + ef) * 2;
                    double z1 = 0, z2 = 0, y = 0;
                    if (live) { z1 = zp[0]; z2 = zp[1]; }
                    for (int tt = 0; tt < M + gw - 1; ++tt) {
                        const double up = __shfl_up_sync(0xffffffffu, y, 1);
                        const int i = tt - ef;
                        if (live && i >= 0 && i < M) {
                            const double x = ef == 0 ? __dmul_rn((double)eb[i], epre) : up;   // preamp first (:66)
                            y = __dadd_rn(__dmul_rn(eb0, x), z1);                                // no FMA contraction (:73-75)
                            const double n1 = __dadd_rn(__dsub_rn(__dmul_rn(eb1, x), __dmul_rn(ea1, y)), z2);
                            const double n2 = __dsub_rn(__dmul_rn(eb2, x), __dmul_rn(ea2, y));
                            z1 = fabs(n1) < 1e-30 ? 0.0 : n1;                                    // :94-97
                            z2 = fabs(n2) < 1e-30 ? 0.0 : n2;
                            if (ef == gw - 1) eb[i] = (float)y;                                  // :88-89
                        }
                    }
                    if (live) { zp[0] = z1; zp[1] = z2; }
                }
                named_sync(BAR_EQ, PG::FFT_THREADS);
                for (int i = ft; i < 2 * T * (M / 2); i += PG::FFT_THREADS) {                    // coalesced store of the tile's block
                    const int ch = i / (M / 2), j = i - ch * (M / 2);
                    if ((ch >> 1) < nvalid) {
                        const float *eb = reinterpret_cast<const float *>(accbuf + (size_t)ch * PS);
                        float *row = a.out.ptr + (size_t)(s0 + (ch >> 1)) * a.out.ss + (ch & 1) * a.out.cs + (size_t)b * M;
                        store_pair(a.out, row, 2 * j, eb[2 * j], eb[2 * j + 1]);
                    }
                }
            }
        };

        int lt = 0, b = 0, ltn = 0, bn = 0;                      // item being finished, item being prepared
        if (n_items > 0) forward_item(0, 0, 0);
        for (int item = 0; item < n_items; ++item) {
            next_item(ltn, bn);
            if (item + 1 < n_items) forward_item(ltn, bn, item + 1);
            mbar_wait_relaxed(acc_ready, (unsigned)(item & 1));
            if (!(dbg & 2)) inverse_item(lt, b, item);
            else if (!MERGED && item + 1 < n_items) publish_heads(item + 1);   // (timing experiments that skip the inverse pass)
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
            lt = ltn; b = bn;
        }
    }
}

namespace {

template <int LOG2M, int T>
cudaError_t launch_persistent_lt(const PersistArgs &a, int ctas, cudaStream_t st)
{
    static bool configured[64] = {};   // opt in to the large dynamic shared memory once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_persistent<LOG2M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PGeo<LOG2M, T>::smem);
        if (e != cudaSuccess) return e;
        // the SM's shared-memory carve-out at its maximum: whatever this CTA leaves (40 KB up to B = 256) stays usable by a
        // co-resident CTA of the equalizer kernel (AW_ENGINE_OVERLAP_EQ)
        e = cudaFuncSetAttribute(k_persistent<LOG2M, T>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    static const bool pdl = !(getenv("AW_PDL") && atoi(getenv("AW_PDL")) == 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3((unsigned)PGeo<LOG2M, T>::THREADS);
    cfg.dynamicSmemBytes = PGeo<LOG2M, T>::smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_persistent<LOG2M, T>, a);
}

}  // namespace

// The fused equalizer runs on the FFT warps: fuse only when one systolic round covers the 2T channels of a tile.
bool persistent_can_fuse_eq(int log2m, int tile, int n_filters)
{
    if (n_filters < 1 || n_filters > 32 || !(persistent_tiles(log2m) & tile)) return false;
    const int G = (1 << log2m) / 16, fft_warps = (8 * G <= 128 ? 128 : 256) / 32;
    return (32 / n_filters) * fft_warps >= 2 * tile;
}

// tiles supported for a transform size, as a bit mask of T
int persistent_tiles(int log2m)
{
    if (log2m >= 6 && log2m <= 9) return 4 | 2;
    if (log2m == 10 || log2m == 11) return 2;   // the transform buffers of 12 FFT warps leave no room for a T = 4 ring
    return 0;
}

int persistent_stage_rows(int log2m)
{
    switch (log2m) {
    case 6: return PGeo<6, 2>::RS;
    case 7: return PGeo<7, 2>::RS;
    case 8: return PGeo<8, 2>::RS;
    case 9: return PGeo<9, 2>::RS;
    case 10: return PGeo<10, 2>::RS;
    case 11: return PGeo<11, 2>::RS;
    default: return 0;
    }
}

int persistent_stage_bins(int log2m)
{
    switch (log2m) {
    case 6: return 2 * PGeo<6, 2>::C;
    case 7: return 2 * PGeo<7, 2>::C;
    case 8: return 2 * PGeo<8, 2>::C;
    case 9: return 2 * PGeo<9, 2>::C;
    case 10: return 2 * PGeo<10, 2>::C;
    case 11: return 2 * PGeo<11, 2>::C;
    default: return 0;
    }
}

cudaError_t launch_persistent(const KpSegment *segs, int n_segs, int Se, int P_cap, int log2m, StridedIn cur, StridedIn prev,
                              float *overlap_save, float2 *fdl, float *fdl_ny, const void *tmaps, StridedOut out, const float2 *tw,
                              int tile, int max_ctas, const KpCall &call, const EqFuse &eq, cudaStream_t st)
{
    if (n_segs <= 0) return cudaSuccess;
    if (n_segs > kKpMaxSegments || !(persistent_tiles(log2m) & tile) || call.nb < 1 || max_ctas < 1) return cudaErrorInvalidValue;
    if (call.nb > 1 && (eq.n_filters != 0 || out.ring_cap > 0)) return cudaErrorInvalidValue;
    if (eq.n_filters != 0 && (n_segs != 1 || !persistent_can_fuse_eq(log2m, tile, eq.n_filters) || out.ring_cap > 0)) return cudaErrorInvalidValue;
    PersistArgs a;
    memset(&a, 0, sizeof(a));
    // Tiles of `tile` streams, numbered across the segments.  The last, partial round of tiles (big % ctas of them) would leave
    // SMs idle while the others stream: it is cut into tiles of tile/2 or tile/4 streams, as many as still fit in one round.
    int big = 0;
    for (int i = 0; i < n_segs; ++i) big += (segs[i].n_streams + tile - 1) / tile;
    if (big <= 0) return cudaSuccess;
    const int rem = big % max_ctas;                        // tiles of the last round (all of them when there is only one)
    int cut = 1;
    if (eq.n_filters == 0)                                  // (the fused equalizer needs whole tiles per systolic round)
        while (rem > 0 && cut * 2 <= tile && rem * cut * 2 <= max_ctas) cut *= 2;
    const int small = tile / cut, keep = big - (cut > 1 ? rem : 0);   // the first `keep` tiles stay whole
    int tiles = 0, seen = 0;
    for (int i = 0; i < n_segs; ++i) {
        a.seg[i] = segs[i];
        const int nt = (segs[i].n_streams + tile - 1) / tile;
        const int nbig = keep - seen < 0 ? 0 : (keep - seen < nt ? keep - seen : nt);
        const int rest = segs[i].n_streams - nbig * tile;
        a.seg[i].tile0 = tiles;
        a.seg[i].n_big = nbig;
        tiles += nbig + (rest > 0 ? (rest + small - 1) / small : 0);
        seen += nt;
    }
    a.n_segs = n_segs; a.n_tiles = tiles; a.small = small; a.Se = Se; a.P_cap = P_cap;
    a.nb = call.nb; a.order = call.order ? 1 : 0; a.keep_pct = call.keep_pct; a.prev_is_rows = call.prev_is_rows;
    a.cur = cur; a.prev = prev; a.overlap_save = overlap_save; a.fdl = fdl; a.fdl_ny = fdl_ny; a.out = out; a.tw = tw;
    a.tmaps = static_cast<const unsigned char *>(tmaps);
    a.debug = call.debug; a.eq = eq;
    const int grid = tiles < max_ctas ? tiles : max_ctas;
#define AW_KP(L, TT) return launch_persistent_lt<L, TT>(a, grid, st)
    if (tile == 4) {
        switch (log2m) {
        case 6: AW_KP(6, 4);
        case 7: AW_KP(7, 4);
        case 8: AW_KP(8, 4);
        case 9: AW_KP(9, 4);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (log2m) {
    case 6: AW_KP(6, 2);
    case 7: AW_KP(7, 2);
    case 8: AW_KP(8, 2);
    case 9: AW_KP(9, 2);
    case 10: AW_KP(10, 2);
    case 11: AW_KP(11, 2);
    default: return cudaErrorInvalidValue;
    }
#undef AW_KP
}

}  // namespace aw
