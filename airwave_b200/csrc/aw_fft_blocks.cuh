// aw_fft_blocks.cuh — device building blocks shared by the FFT-bearing kernels (aw_fft_kernels.cu, aw_persistent.cu):
// padded shared-memory passes of the register-radix FFT core, the forward/inverse real-FFT frames with fused
// overlap-save assembly/discard (ConvolutionEngine.swift:237-252, 353-366), the Nyquist product sum, the complex
// multiply-accumulate, and the cp.async / mbarrier / bulk-copy (TMA) / named-barrier PTX wrappers.
#pragma once
#include <stdint.h>

#include "aw_fft_reg.cuh"
#include "aw_kernels.h"

namespace aw {

using namespace awfft;

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Wait that is expected to last (a whole tile): back off between polls so that the waiting warps leave the issue slots — and the
// power budget — to the warps that stream.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITR_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONER_%=;\n\t"
        "nanosleep.u32 256;\n\t"
        "bra WAITR_%=;\n\t"
        "DONER_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bulk copy with an L2 eviction-priority hint (createpolicy): evict_first for data that is read once per launch
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// evict_last for `pct` percent of the lines a copy touches (hardware picks which), evict_first for the rest: keeps a chosen
// share of a streamed working set resident when all of it would not fit
__device__ __forceinline__ uint64_t l2_policy_evict_last_share(int pct)
{
    uint64_t pol;
    if (pct >= 88) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (pct >= 63) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.75;" : "=l"(pol));
    else if (pct >= 38) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.5;" : "=l"(pol));
    else if (pct >= 13) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.25;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// Tensor-map TMA load of a 5-D box (cp.async.bulk.tensor, SASS UTMALDG): one instruction moves rows x streams x bins of the
// FDL into a stage, whatever the strides between them in HBM.  `tmap` is the generic address of a CUtensorMap in global memory.
__device__ __forceinline__ void tma_load_5d_hint(void *smem_dst, const void *tmap, int c0, int c1, int c2, int c3, int c4, uint64_t *bar,
                                                 uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::
            "r"(smem_u32(smem_dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
// 8-byte load of data that is read once (frame operands, accumulators): does not allocate in L1, which K2 / K4 keep for the
// twiddle tables they read from global memory
__device__ __forceinline__ float2 ld_once2(const float *p)
{
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cmac2f(float4 &acc, const float4 x, const float hr0, const float hi0, const float hr1, const float hi1)
{
    acc.x = fmaf(x.x, hr0, acc.x); acc.x = fmaf(-x.y, hi0, acc.x);
    acc.y = fmaf(x.x, hi0, acc.y); acc.y = fmaf(x.y, hr0, acc.y);
    acc.z = fmaf(x.z, hr1, acc.z); acc.z = fmaf(-x.w, hi1, acc.z);
    acc.w = fmaf(x.z, hi1, acc.w); acc.w = fmaf(x.w, hr1, acc.w);
}

// Barrier used between the phases of a transform computed by more than one warp: named barrier `id` over `count`
// threads.  {0, 0} = the whole CTA (__syncthreads), which is what the stand-alone kernels use.
struct GroupBar {
    int id, count;
};

// ------------------------------------------------------------------------------------------------
// building blocks
// ------------------------------------------------------------------------------------------------
template <int LOG2M, int P>
__device__ __forceinline__ void smem_load(const float2 *buf, float2 (&v)[RegFft<LOG2M>::E], int t)
{
#pragma unroll
    for (int e = 0; e < RegFft<LOG2M>::E; ++e) v[e] = buf[pad16(RegFft<LOG2M>::template load_index<P>(t, e))];
}

template <int LOG2M, int P>
__device__ __forceinline__ void smem_store(float2 *buf, const float2 (&v)[RegFft<LOG2M>::E], int t)
{
#pragma unroll
    for (int e = 0; e < RegFft<LOG2M>::E; ++e) buf[pad16(RegFft<LOG2M>::template store_index<P>(t, e))] = v[e];
}

// Barrier between the phases of a pass.  A transform is computed by G consecutive threads; when G <= 32 they all sit in
// one warp, so a warp-level barrier (plus its memory ordering) is enough and the warps of a CTA run their transforms
// independently of each other — the loads of one warp overlap the butterflies of another.
template <int LOG2M>
__device__ __forceinline__ void group_sync(GroupBar gb = GroupBar{0, 0})
{
    if constexpr (RegFft<LOG2M>::G <= 32) __syncwarp();
    else if (gb.count == 0) __syncthreads();
    else named_sync(gb.id, gb.count);
}

// Passes [P, LAST] entirely in shared memory (in place); every thread of the transform's group must call it.
// PT: `tw` is a per-pass twiddle table (RegFft::pt_entry layout) instead of the half-circle table.
template <int LOG2M, int P, int PT>
__device__ __forceinline__ void pass_compute(float2 (&v)[RegFft<LOG2M>::E], const float2 *tw, int t)
{
    if constexpr (PT != 0) RegFft<LOG2M>::template compute_pt<P, PT == 2>(v, tw, t);
    else RegFft<LOG2M>::template compute<P>(v, tw, t);
}

template <int LOG2M, int P, int LAST, int PT = 0>
struct SmemPasses {
    __device__ __forceinline__ static void run(float2 *buf, const float2 *tw, int t, GroupBar gb = GroupBar{0, 0})
    {
        if constexpr (P <= LAST) {
            float2 v[RegFft<LOG2M>::E];
            smem_load<LOG2M, P>(buf, v, t);
            group_sync<LOG2M>(gb);
            pass_compute<LOG2M, P, PT>(v, tw, t);
            smem_store<LOG2M, P>(buf, v, t);
            group_sync<LOG2M>(gb);
            SmemPasses<LOG2M, P + 1, LAST, PT>::run(buf, tw, t, gb);
        }
    }
};

// Forward real FFT of one frame.  `load(i)` returns z[i] = x[2i] + i*x[2i+1] of the frame (i < M); on return the
// padded buffer holds Z and (after the trailing barrier) `emit(k, X)` has been called by the owning threads for
// k = 0..M-1 with the packed spectrum 2*X[k] (k = 0: (2*DC, 0)) and `emit_ny(2*X[M])` once.
// Body of the forward real FFT of one frame whose pass-0 operands are already in registers (v[e] = z[load_index<0>(t, e)]):
// lets a caller fetch the next frame from global memory while this one is being transformed.
template <int LOG2M, int PT = 0, class Emit, class EmitNy>
__device__ __forceinline__ void forward_frame_regs(float2 *buf, const float2 *tw, int t, bool active, float2 (&v)[RegFft<LOG2M>::E],
                                                   Emit emit, EmitNy emit_ny, GroupBar gb = GroupBar{0, 0})
{
    using F = RegFft<LOG2M>;
    constexpr int M = F::M;
    F::template compute<0>(v, tw, t);   // pass 0 has no twiddles
    smem_store<LOG2M, 0>(buf, v, t);
    group_sync<LOG2M>(gb);
    SmemPasses<LOG2M, 1, F::PASSES - 1, PT>::run(buf, tw, t, gb);
    if (active) {
        for (int k = t; k <= M / 2; k += F::G) {
            if (k == 0) {
                const float2 z0 = buf[0];
                emit(0, make_float2(2.0f * (z0.x + z0.y), 0.0f));
                emit_ny(2.0f * (z0.x - z0.y));
            } else {
                const int j = M - k;
                const float2 a = buf[pad16(k)], b = buf[pad16(j)];
                const float er = a.x + b.x, ei = a.y - b.y;   // E = Z[k] + conj(Z[M-k])
                const float dr = a.x - b.x, di = a.y + b.y;   // D = Z[k] - conj(Z[M-k])
                const float2 w = tw[k];                       // exp(-2*pi*i*k/N)
                const float tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;
                emit(k, make_float2(er + ti, ei - tr));
                if (j != k) emit(j, make_float2(er - ti, -ei - tr));
            }
        }
    }
    group_sync<LOG2M>(gb);
}

// Forward real FFT of one frame.  `load(i)` returns z[i] = x[2i] + i*x[2i+1] of the frame (i < M); `emit(k, X)` is called
// by the owning threads for k = 0..M-1 with the packed spectrum 2*X[k] (k = 0: (2*DC, 0)) and `emit_ny(2*X[M])` once.
template <int LOG2M, int PT = 0, class Load, class Emit, class EmitNy>
__device__ __forceinline__ void forward_frame(float2 *buf, const float2 *tw, int t, bool active, Load load, Emit emit, EmitNy emit_ny,
                                              GroupBar gb = GroupBar{0, 0})
{
    using F = RegFft<LOG2M>;
    float2 v[F::E];
#pragma unroll
    for (int e = 0; e < F::E; ++e) v[e] = active ? load(F::template load_index<0>(t, e), e) : make_float2(0.f, 0.f);
    forward_frame_regs<LOG2M, PT>(buf, tw, t, active, v, emit, emit_ny, gb);
}

// Inverse real FFT: the padded buffer holds the packed spectrum acc[0..M) (acc[0].x = DC) and ny the Nyquist value;
// `emit(i, x0, x1)` receives the time samples x[2i], x[2i+1] for i in [M/2, M) — the overlap-save "second half".
// All threads of the CTA must call it (barriers); `active` masks the stores only.
// SYNC_BEFORE_EMIT: barrier between the last pass's loads and `emit`, for callers whose emit overwrites `buf`.
template <int LOG2M, bool SYNC_BEFORE_EMIT = false, int PT = 0, class Emit>
__device__ __forceinline__ void inverse_frame(float2 *buf, float ny, const float2 *tw, int t, bool active, Emit emit,
                                              GroupBar gb = GroupBar{0, 0})
{
    using F = RegFft<LOG2M>;
    constexpr int M = F::M;
    // inverse split, in place; stores conj(Z) so that the forward passes compute conj(IFFT(Z))
    for (int k = t; k <= M / 2; k += F::G) {
        if (k == 0) {
            const float dc = buf[0].x;
            buf[0] = make_float2(dc + ny, -(dc - ny));
        } else {
            const int j = M - k;
            const float2 a = buf[pad16(k)], b = buf[pad16(j)];
            const float er = a.x + b.x, ei = a.y - b.y;
            const float dr = a.x - b.x, di = a.y + b.y;
            const float2 w = make_float2(tw[k].x, -tw[k].y);
            const float tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;
            buf[pad16(k)] = make_float2(er - ti, -(ei + tr));
            if (j != k) buf[pad16(j)] = make_float2(er + ti, -(-ei + tr));
        }
    }
    group_sync<LOG2M>(gb);
    SmemPasses<LOG2M, 0, F::PASSES - 2, PT>::run(buf, tw, t, gb);
    {
        constexpr int P = F::PASSES - 1;
        float2 v[F::E];
        smem_load<LOG2M, P>(buf, v, t);
        if constexpr (SYNC_BEFORE_EMIT) group_sync<LOG2M>(gb);
        pass_compute<LOG2M, P, PT>(v, tw, t);
        if (active) {
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int i = F::template store_index<P>(t, e);
                if (i >= M / 2) emit(i - M / 2, v[e].x, -v[e].y);
            }
        }
    }
    group_sync<LOG2M>(gb);
}

// inverse_frame with the packed spectrum still in global memory (K4): the inverse split is applied on the way into the transform
// buffer (one pass over shared memory and one barrier less than copying first).  `ny_after_barrier()` is called by every thread
// after the first barrier and returns the Nyquist value in thread 0 of the transform, which owns bin 0 here and in pass 0.
template <int LOG2M, int PT = 0, class Ny, class Emit>
__device__ __forceinline__ void inverse_frame_global(float2 *buf, const float2 *src, const float2 *tw, int t, bool active, Ny ny_after_barrier,
                                                     Emit emit, GroupBar gb = GroupBar{0, 0})
{
    using F = RegFft<LOG2M>;
    constexpr int M = F::M, G = F::G;
    constexpr int NP = (M / 2 + G) / G;            // bins k = t + n*G <= M/2 per thread, at most
    if (active) {
        float2 a[NP], b[NP];
#pragma unroll
        for (int n = 0; n < NP; ++n) {
            const int k = t + n * G;
            if (k <= M / 2) {
                a[n] = ld_once2(reinterpret_cast<const float *>(src + k));
                b[n] = k == 0 ? a[n] : ld_once2(reinterpret_cast<const float *>(src + (M - k)));
            }
        }
#pragma unroll
        for (int n = 0; n < NP; ++n) {
            const int k = t + n * G;
            if (k == 0) {
                buf[0] = a[n];                     // (DC, -): combined with the Nyquist value below
            } else if (k <= M / 2) {
                const int j = M - k;
                const float er = a[n].x + b[n].x, ei = a[n].y - b[n].y;
                const float dr = a[n].x - b[n].x, di = a[n].y + b[n].y;
                const float2 w = make_float2(tw[k].x, -tw[k].y);
                const float tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;
                buf[pad16(k)] = make_float2(er - ti, -(ei + tr));
                if (j != k) buf[pad16(j)] = make_float2(er + ti, -(-ei + tr));
            }
        }
    }
    group_sync<LOG2M>(gb);
    const float ny = ny_after_barrier();
    if (t == 0) {                                  // the same thread loads bin 0 in pass 0: no barrier needed
        const float dc = buf[0].x;
        buf[0] = make_float2(dc + ny, -(dc - ny));
    }
    SmemPasses<LOG2M, 0, F::PASSES - 2, PT>::run(buf, tw, t, gb);
    {
        constexpr int P = F::PASSES - 1;
        float2 v[F::E];
        smem_load<LOG2M, P>(buf, v, t);
        pass_compute<LOG2M, P, PT>(v, tw, t);
        if (active) {
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int i = F::template store_index<P>(t, e);
                if (i >= M / 2) emit(i - M / 2, v[e].x, -v[e].y);
            }
        }
    }
    group_sync<LOG2M>(gb);
}

// The input channels FDL row `s` sums (c[0] is always one; -1 = no more): read once per frame, so that the operand loads below do
// not each wait for a table lookup.
struct RowSources { int c[kKpMaxRowSources]; };
__device__ __forceinline__ RowSources row_sources(const BlockGeom &g, int s)
{
    RowSources r;
    if (g.rows) {
        const char4 v = *reinterpret_cast<const char4 *>(g.rows->src[s]);
        r.c[0] = v.x; r.c[1] = v.y; r.c[2] = v.z; r.c[3] = v.w;
    } else {
        r.c[0] = s; r.c[1] = r.c[2] = r.c[3] = -1;
    }
    return r;
}
static_assert(kKpMaxRowSources == 4, "row_sources reads the four sources of a row as one char4");

// Operand i of the packed overlap-save frame [previous block | current block] of FDL row `s` of one stream (K2, KF): a row sums
// the input channels that share its filter pair; the overlap buffer holds the previous block already summed (prev_is_rows), the
// caller's input does not.
template <int M>
__device__ __forceinline__ float2 frame_operand(const BlockGeom &g, const RowSources &rs, const StridedIn &prev, const StridedIn &cur, int stream,
                                                int s, int i)
{
    float2 v;
    if (i < M / 2) {
        v = ld_once2(prev.ptr + stream * prev.ss + (g.prev_is_rows ? s : rs.c[0]) * prev.cs + 2 * i);
        if (!g.prev_is_rows) {
#pragma unroll
            for (int q = 1; q < kKpMaxRowSources; ++q)
                if (rs.c[q] >= 0) {
                    const float2 x = ld_once2(prev.ptr + stream * prev.ss + rs.c[q] * prev.cs + 2 * i);
                    v.x += x.x; v.y += x.y;
                }
        }
        return v;
    }
    const int j = 2 * (i - M / 2);
    v = ld_once2(cur.ptr + stream * cur.ss + rs.c[0] * cur.cs + j);
#pragma unroll
    for (int q = 1; q < kKpMaxRowSources; ++q)
        if (rs.c[q] >= 0) {
            const float2 x = ld_once2(cur.ptr + stream * cur.ss + rs.c[q] * cur.cs + j);
            v.x += x.x; v.y += x.y;
        }
    return v;
}

// The pass-0 operands of thread t for that frame (v[e] = z[load_index<0>(t, e)]), all loads first; then inputOverlapBuffer <- the
// (summed) current block (ConvolutionEngine.swift:243) when `ov` is given: the overlap buffer may be `prev` itself, and a store
// between the loads would make each of them wait for the one before.
template <int LOG2M>
__device__ __forceinline__ void fetch_frame(const BlockGeom &g, const StridedIn &prev, const StridedIn &cur, int stream, int s, float *ov, int t,
                                            bool active, float2 (&v)[RegFft<LOG2M>::E])
{
    using F = RegFft<LOG2M>;
    constexpr int M = F::M;
    const RowSources rs = row_sources(g, s);
#pragma unroll
    for (int e = 0; e < F::E; ++e)
        v[e] = active ? frame_operand<M>(g, rs, prev, cur, stream, s, F::template load_index<0>(t, e)) : make_float2(0.f, 0.f);
    if (active && ov) {
#pragma unroll
        for (int e = 0; e < F::E; ++e) {
            const int i = F::template load_index<0>(t, e);
            if (i >= M / 2) *reinterpret_cast<float2 *>(ov + 2 * (i - M / 2)) = v[e];
        }
    }
}

__device__ __forceinline__ float group_sum(float v, int width)   // deterministic butterfly sum over `width` (<= 32) lanes
{
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Nyquist product sum for one (stream, ear): sum_{s,p} fdl_ny[stream][s][(head+p)%P] * bank_ny[s][p][ear], computed by
// the G threads of a transform.  part_s: G floats of scratch for this transform.  Result valid in every thread after
// the two barriers inside.
// Thread t's share of that sum (partitions t, t + G, ...).
template <int G>
__device__ __forceinline__ float nyquist_partial(const BlockGeom &g, const float *fdl_ny, const float *bank_ny, int stream, int ear, bool active,
                                                 int t)
{
    float sum = 0.f;
    const int Pm = g.Pm > 0 ? g.Pm : g.P;
    if (active) {
        for (int s = 0; s < g.S; ++s) {
            const float *xrow = fdl_ny + ((size_t)stream * g.Se + s) * g.P_cap;
            const float *hrow = bank_ny + (size_t)s * g.P * 2 + ear;
            for (int p0 = t; p0 < g.P; p0 += 8 * G) {     // 8 independent load pairs in flight per thread
                float xa[8], ha[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int p = p0 + k * G;
                    xa[k] = 0.f; ha[k] = 0.f;
                    if (p < g.P) {
                        int slot = g.head + p;
                        if (slot >= Pm) slot -= Pm;
                        xa[k] = xrow[slot];
                        ha[k] = hrow[2 * p];
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) sum = fmaf(xa[k], ha[k], sum);
            }
        }
    }
    return sum;
}
// The sum of the G partial sums in part_s[0..G), in the order nyquist_sum adds them; call with the threads t < 32 of the transform
// (G > 32) after a barrier that follows the stores to part_s.  Result in each of them.
template <int G>
__device__ __forceinline__ float nyquist_reduce_warp0(const float *part_s, int t)
{
    float acc = 0.f;
    for (int i = t; i < G; i += 32) acc += part_s[i];
    return group_sum(acc, 32);
}
template <int G>
__device__ __forceinline__ float nyquist_sum(const BlockGeom &g, const float *fdl_ny, const float *bank_ny, int stream, int ear,
                                             bool active, int t, float *part_s, GroupBar gb = GroupBar{0, 0})
{
    const float sum = nyquist_partial<G>(g, fdl_ny, bank_ny, stream, ear, active, t);
    if constexpr (G <= 32) {
        return group_sum(sum, G);
    } else {
        auto sync = [&]() { if (gb.count == 0) __syncthreads(); else named_sync(gb.id, gb.count); };
        part_s[t] = sum;
        sync();
        float acc = 0.f;
        if (t < 32) acc = nyquist_reduce_warp0<G>(part_s, t);
        sync();               // everyone has read part_s[.] before slot 0 is overwritten
        if (t == 0) part_s[0] = acc;
        sync();
        acc = part_s[0];
        sync();               // part_s may be reused by the caller's next transform
        return acc;
    }
}

__device__ __forceinline__ void store_pair(const StridedOut &o, float *row, int i2, float x0, float x1)
{
    if (o.ring_cap > 0) {
        int p0 = o.ring_start + i2;
        if (p0 >= o.ring_cap) p0 -= o.ring_cap;
        int p1 = p0 + 1;
        if (p1 >= o.ring_cap) p1 -= o.ring_cap;
        row[p0] = x0;
        row[p1] = x1;
    } else if ((reinterpret_cast<uintptr_t>(row + i2) & 7u) == 0) {
        *reinterpret_cast<float2 *>(row + i2) = make_float2(x0, x1);
    } else {
        row[i2] = x0;
        row[i2 + 1] = x1;
    }
}

}  // namespace aw
