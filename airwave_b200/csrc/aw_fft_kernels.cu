// aw_fft_kernels.cu — FFT-bearing kernels built on the register-radix core (aw_fft_reg.cuh):
//
//   K1 k_bank_build<LOG2M>   partition + zero-pad + forward real FFT of every HRIR   (ConvolutionEngine.swift:143-182)
//   K2 k_input_rfft<LOG2M>   [prev | cur] frame -> forward real FFT -> FDL head slot  (ConvolutionEngine.swift:237-264)
//   K4 k_irfft_out<LOG2M>    Nyquist sum + inverse real FFT + overlap-save discard    (ConvolutionEngine.swift:353-366)
//   KF k_fused<LOG2M, T>     K2 + K3 + K4 for a tile of T streams in ONE kernel: the spectra of the tile are
//                            produced, the FDL is streamed exactly once with the accumulators in registers,
//                            and the two ear spectra go straight through the inverse transform — no
//                            intermediate (acc) ever reaches HBM.  Used for 64 <= B <= 512.
//
// Thread organisation: a transform of M = B complex points is done by G = M/16 threads holding 16 values each;
// a CTA runs NF transforms side by side.  In the fused kernel THREADS = B/2 (one bin pair per thread in the MAC
// phase) which makes NF = 8 for every supported B.
#include <stdint.h>

#include "aw_fft_reg.cuh"
#include "aw_kernels.h"

namespace aw {

using namespace awfft;

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
__device__ __forceinline__ void group_sync()
{
    if constexpr (RegFft<LOG2M>::G <= 32) __syncwarp();
    else __syncthreads();
}

// Passes [P, LAST] entirely in shared memory (in place); every thread of the transform's group must call it.
template <int LOG2M, int P, int LAST>
struct SmemPasses {
    __device__ __forceinline__ static void run(float2 *buf, const float2 *tw, int t)
    {
        if constexpr (P <= LAST) {
            float2 v[RegFft<LOG2M>::E];
            smem_load<LOG2M, P>(buf, v, t);
            group_sync<LOG2M>();
            RegFft<LOG2M>::template compute<P>(v, tw, t);
            smem_store<LOG2M, P>(buf, v, t);
            group_sync<LOG2M>();
            SmemPasses<LOG2M, P + 1, LAST>::run(buf, tw, t);
        }
    }
};

// Forward real FFT of one frame.  `load(i)` returns z[i] = x[2i] + i*x[2i+1] of the frame (i < M); on return the
// padded buffer holds Z and (after the trailing barrier) `emit(k, X)` has been called by the owning threads for
// k = 0..M-1 with the packed spectrum 2*X[k] (k = 0: (2*DC, 0)) and `emit_ny(2*X[M])` once.
// Body of the forward real FFT of one frame whose pass-0 operands are already in registers (v[e] = z[load_index<0>(t, e)]):
// lets a caller fetch the next frame from global memory while this one is being transformed.
template <int LOG2M, class Emit, class EmitNy>
__device__ __forceinline__ void forward_frame_regs(float2 *buf, const float2 *tw, int t, bool active, float2 (&v)[RegFft<LOG2M>::E],
                                                   Emit emit, EmitNy emit_ny)
{
    using F = RegFft<LOG2M>;
    constexpr int M = F::M;
    F::template compute<0>(v, tw, t);
    smem_store<LOG2M, 0>(buf, v, t);
    group_sync<LOG2M>();
    SmemPasses<LOG2M, 1, F::PASSES - 1>::run(buf, tw, t);
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
    group_sync<LOG2M>();
}

// Forward real FFT of one frame.  `load(i)` returns z[i] = x[2i] + i*x[2i+1] of the frame (i < M); `emit(k, X)` is called
// by the owning threads for k = 0..M-1 with the packed spectrum 2*X[k] (k = 0: (2*DC, 0)) and `emit_ny(2*X[M])` once.
template <int LOG2M, class Load, class Emit, class EmitNy>
__device__ __forceinline__ void forward_frame(float2 *buf, const float2 *tw, int t, bool active, Load load, Emit emit, EmitNy emit_ny)
{
    using F = RegFft<LOG2M>;
    float2 v[F::E];
#pragma unroll
    for (int e = 0; e < F::E; ++e) v[e] = active ? load(F::template load_index<0>(t, e), e) : make_float2(0.f, 0.f);
    forward_frame_regs<LOG2M>(buf, tw, t, active, v, emit, emit_ny);
}

// Inverse real FFT: the padded buffer holds the packed spectrum acc[0..M) (acc[0].x = DC) and ny the Nyquist value;
// `emit(i, x0, x1)` receives the time samples x[2i], x[2i+1] for i in [M/2, M) — the overlap-save "second half".
// All threads of the CTA must call it (barriers); `active` masks the stores only.
template <int LOG2M, class Emit>
__device__ __forceinline__ void inverse_frame(float2 *buf, float ny, const float2 *tw, int t, bool active, Emit emit)
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
    group_sync<LOG2M>();
    SmemPasses<LOG2M, 0, F::PASSES - 2>::run(buf, tw, t);
    {
        constexpr int P = F::PASSES - 1;
        float2 v[F::E];
        smem_load<LOG2M, P>(buf, v, t);
        F::template compute<P>(v, tw, t);
        if (active) {
#pragma unroll
            for (int e = 0; e < F::E; ++e) {
                const int i = F::template store_index<P>(t, e);
                if (i >= M / 2) emit(i - M / 2, v[e].x, -v[e].y);
            }
        }
    }
    group_sync<LOG2M>();
}

__device__ __forceinline__ float group_sum(float v, int width)   // deterministic butterfly sum over `width` (<= 32) lanes
{
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Nyquist product sum for one (stream, ear): sum_{s,p} fdl_ny[stream][s][(head+p)%P] * bank_ny[s][p][ear], computed by
// the G threads of a transform.  part_s: G floats of scratch for this transform.  Result valid in every thread after
// the two barriers inside.
template <int G>
__device__ __forceinline__ float nyquist_sum(const BlockGeom &g, const float *fdl_ny, const float *bank_ny, int stream, int ear,
                                             bool active, int t, float *part_s)
{
    float sum = 0.f;
    if (active) {
        const int terms = g.S * g.P;
        for (int i0 = t; i0 < terms; i0 += 8 * G) {     // 8 independent load pairs in flight per thread
            float xa[8], ha[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = i0 + k * G;
                xa[k] = 0.f; ha[k] = 0.f;
                if (i < terms) {
                    const int s = i / g.P, p = i - s * g.P;
                    int slot = g.head + p;
                    if (slot >= g.P) slot -= g.P;
                    xa[k] = fdl_ny[((size_t)stream * g.Se + s) * g.P_cap + slot];
                    ha[k] = bank_ny[(size_t)i * 2 + ear];
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) sum = fmaf(xa[k], ha[k], sum);
        }
    }
    if constexpr (G <= 32) {
        return group_sum(sum, G);
    } else {
        part_s[t] = sum;
        __syncthreads();
        float acc = 0.f;
        if (t < 32) {
            for (int i = t; i < G; i += 32) acc += part_s[i];
        }
        acc = group_sum(acc, 32);
        if (t == 0) part_s[0] = acc;
        __syncthreads();
        return part_s[0];
    }
}

template <int LOG2M>
struct Geo {
    using F = RegFft<LOG2M>;
    static constexpr int M = F::M, G = F::G;
    static constexpr int SA_THREADS = G > 128 ? G : 128;      // stand-alone K1/K2/K4
    static constexpr int SA_NF = SA_THREADS / G;
    static constexpr int PS = PaddedSize<LOG2M>::value;
    static constexpr size_t sa_smem = (size_t)(M + SA_NF * PS) * sizeof(float2) + (size_t)SA_THREADS * sizeof(float);
};

// ------------------------------------------------------------------------------------------------
// K2  input_rfft
// ------------------------------------------------------------------------------------------------
struct InputRfftArgs {
    BlockGeom g;
    StridedIn cur, prev;
    float *overlap_save;
    float2 *fdl;
    float *fdl_ny;
    const float2 *tw;
};

template <int LOG2M>
__global__ void __launch_bounds__(Geo<LOG2M>::SA_THREADS) k_input_rfft(const InputRfftArgs a)
{
    using Gm = Geo<LOG2M>;
    constexpr int M = Gm::M, G = Gm::G, NF = Gm::SA_NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufs = tw + M;
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    for (int k = tid; k < M; k += Gm::SA_THREADS) tw[k] = a.tw[k];
    __syncthreads();
    const int job = blockIdx.x * NF + f;
    const bool active = job < a.g.n_streams * a.g.S;
    const int ls = active ? job / a.g.S : 0, s = active ? job - ls * a.g.S : 0;
    const int stream = a.g.first_stream + ls;
    const float *prev = a.prev.ptr + stream * a.prev.ss + s * a.prev.cs;
    const float *cur = a.cur.ptr + stream * a.cur.ss + s * a.cur.cs;
    float *ov = a.overlap_save ? a.overlap_save + ((size_t)stream * a.g.Se + s) * M : nullptr;
    const size_t row = ((size_t)stream * a.g.Se + s) * a.g.P_cap + a.g.head;
    float2 *dst = a.fdl + row * M;
    float *dst_ny = a.fdl_ny + row;
    forward_frame<LOG2M>(
        bufs + (size_t)f * Gm::PS, tw, t, active,
        [&](int i, int) -> float2 {   // frame = [previous block | current block]  (:237-248)
            if (i < M / 2) return *reinterpret_cast<const float2 *>(prev + 2 * i);
            const float2 v = *reinterpret_cast<const float2 *>(cur + 2 * (i - M / 2));
            if (ov) *reinterpret_cast<float2 *>(ov + 2 * (i - M / 2)) = v;   // inputOverlapBuffer <- current block (:243);
            return v;                                                          // same thread read this address as `prev`
        },
        [&](int k, float2 x) { dst[k] = x; },       // FDL[head] <- spectrum (:256-264)
        [&](float ny) { *dst_ny = ny; });
}

// ------------------------------------------------------------------------------------------------
// K1  bank_build
// ------------------------------------------------------------------------------------------------
struct BankArgs {
    const float *ir;   // [S][2][taps]
    int S, taps, B, P;
    float4 *bank;
    float *bank_ny;
    const float2 *tw;
};

template <int LOG2M>
__global__ void __launch_bounds__(Geo<LOG2M>::SA_THREADS) k_bank_build(const BankArgs a)
{
    using Gm = Geo<LOG2M>;
    constexpr int M = Gm::M, G = Gm::G, NF = Gm::SA_NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufs = tw + M;
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    for (int k = tid; k < M; k += Gm::SA_THREADS) tw[k] = a.tw[k];
    __syncthreads();
    const int job = blockIdx.x * NF + f;   // (s*2 + ear)*P + p
    const bool active = job < a.S * 2 * a.P;
    const int se = active ? job / a.P : 0, p = active ? job - se * a.P : 0;
    const float *h = a.ir + (size_t)se * a.taps;
    const float scale = 0.25f / (float)(2 * M);   // ConvolutionEngine.swift:356, folded into the bank (exact: power of two)
    float *dst = reinterpret_cast<float *>(a.bank + ((size_t)(se >> 1) * a.P + p) * M) + 2 * (se & 1);
    float *dst_ny = a.bank_ny + ((size_t)(se >> 1) * a.P + p) * 2 + (se & 1);
    forward_frame<LOG2M>(
        bufs + (size_t)f * Gm::PS, tw, t, active,
        [&](int i, int) -> float2 {   // h[pB .. (p+1)B) || 0_B  (:145-155)
            float2 v = make_float2(0.f, 0.f);
            if (i < M / 2) {
                const int t0 = p * M + 2 * i;
                if (t0 < a.taps) v.x = h[t0];
                if (t0 + 1 < a.taps) v.y = h[t0 + 1];
            }
            return v;
        },
        [&](int k, float2 x) { dst[4 * k] = x.x * scale; dst[4 * k + 1] = x.y * scale; },
        [&](float ny) { *dst_ny = ny * scale; });
}

// ------------------------------------------------------------------------------------------------
// K4  irfft_out
// ------------------------------------------------------------------------------------------------
struct IrfftArgs {
    BlockGeom g;
    const float2 *acc;
    const float *fdl_ny;
    const float *bank_ny;
    StridedOut out;
    const float2 *tw;
};

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

template <int LOG2M>
__global__ void __launch_bounds__(Geo<LOG2M>::SA_THREADS) k_irfft_out(const IrfftArgs a)
{
    using Gm = Geo<LOG2M>;
    constexpr int M = Gm::M, G = Gm::G, NF = Gm::SA_NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufs = tw + M;
    float *part = reinterpret_cast<float *>(bufs + (size_t)NF * Gm::PS);
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    for (int k = tid; k < M; k += Gm::SA_THREADS) tw[k] = a.tw[k];
    const int job = blockIdx.x * NF + f;   // (local stream, ear)
    const bool active = job < a.g.n_streams * 2;
    const int stream = a.g.first_stream + (active ? job >> 1 : 0), ear = job & 1;
    float2 *buf = bufs + (size_t)f * Gm::PS;
    if (active) {
        const float2 *src = a.acc + ((size_t)stream * 2 + ear) * M;
        for (int k = t; k < M; k += G) buf[pad16(k)] = src[k];
    }
    __syncthreads();
    const float ny = nyquist_sum<G>(a.g, a.fdl_ny, a.bank_ny, stream, ear, active, t, part + (size_t)f * G);
    float *row = a.out.ptr + stream * a.out.ss + ear * a.out.cs;
    inverse_frame<LOG2M>(buf, ny, tw, t, active, [&](int i, float x0, float x1) { store_pair(a.out, row, 2 * i, x0, x1); });
}

// ------------------------------------------------------------------------------------------------
// FDL multiply-accumulate phase shared by KF (fused) and K3 (stand-alone).
// One thread owns one bin pair (two adjacent complex bins) of T streams.  The FDL rows are streamed with
// cp.async (LDGSTS, 16 B per copy, L1 bypass) into a PRIVATE shared-memory ring: every thread consumes only
// the bytes it copied itself, so the pipeline needs no barrier at all — cp.async.wait_group is the only
// synchronisation — and the number of bytes in flight (STAGES-1 iterations x T x 16 B per thread) no longer
// costs registers.  Filter values (both ears of both bins = 2 x float4) are prefetched one iteration ahead
// in registers and reused for the T streams of the tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cmac2f(float4 &acc, const float4 x, const float hr0, const float hi0, const float hr1, const float hi1)
{
    acc.x = fmaf(x.x, hr0, acc.x); acc.x = fmaf(-x.y, hi0, acc.x);
    acc.y = fmaf(x.x, hi0, acc.y); acc.y = fmaf(x.y, hr0, acc.y);
    acc.z = fmaf(x.z, hr1, acc.z); acc.z = fmaf(-x.w, hi1, acc.z);
    acc.w = fmaf(x.z, hi1, acc.w); acc.w = fmaf(x.w, hr1, acc.w);
}

constexpr int kMacStages = 3;

// ring: kMacStages * T * THREADS float4.  fp[u]: FDL base of stream u at this thread's bin pair (float4 units);
// bank_jp: bank + 2*jp.  Accumulates partitions p in [p0, p1) of every speaker into aL/aR (left/right ear, 2 bins each).
template <int T, int THREADS>
__device__ __forceinline__ void mac_phase(const BlockGeom &g, const float4 *const (&fp)[T], const float4 *bank_jp, float4 *ring,
                                          float4 (&aL)[T], float4 (&aR)[T], int p0, int p1)
{
    const int tid = threadIdx.x;
    const int halfB = g.B >> 1;
    const int np = p1 - p0;
    const int total = g.S * np;
    if (total <= 0) return;
    int slot0 = g.head + p0;                          // ring slot of partition p0 (modulus is partitionCount, Q4)
    if (slot0 >= g.P) slot0 -= g.P;
    float4 *my = ring + tid;
    // issue state (runs kMacStages-1 iterations ahead of the consume state)
    int is = 0, ip = 0, islot = slot0, istage = 0;
    auto issue = [&]() {
        if (is < g.S) {
            const size_t off = ((size_t)is * g.P_cap + islot) * halfB;
#pragma unroll
            for (int u = 0; u < T; ++u) cp_async16(my + (istage * T + u) * THREADS, fp[u] + off);
            if (++ip == np) { ip = 0; ++is; islot = slot0; }
            else islot = (islot + 1 == g.P) ? 0 : islot + 1;
            istage = (istage + 1 == kMacStages) ? 0 : istage + 1;
        }
        cp_async_commit();   // always commit so the group count stays in step with the iteration count
    };
#pragma unroll
    for (int k = 0; k < kMacStages - 1; ++k) issue();
    // filter walk: (s*P + p) * B float4; consecutive p are contiguous, a speaker change skips the partitions outside [p0, p1)
    const float4 *bk = bank_jp + (size_t)p0 * g.B;
    const size_t skip = (size_t)(g.P - np) * g.B;
    float4 h0 = __ldg(bk), h1 = __ldg(bk + 1);
    int cstage = 0, cp = 0;
    for (int it = 0; it < total; ++it) {
        issue();
        bk += g.B;
        if (++cp == np) { cp = 0; bk += skip; }
        float4 n0 = h0, n1 = h1;
        if (it + 1 < total) { n0 = __ldg(bk); n1 = __ldg(bk + 1); }
        cp_async_wait<kMacStages - 1>();
#pragma unroll
        for (int u = 0; u < T; ++u) {
            const float4 x = my[(cstage * T + u) * THREADS];
            cmac2f(aL[u], x, h0.x, h0.y, h1.x, h1.y);
            cmac2f(aR[u], x, h0.z, h0.w, h1.z, h1.w);
        }
        h0 = n0; h1 = n1;
        cstage = (cstage + 1 == kMacStages) ? 0 : cstage + 1;
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// K3  fdl_cmac (stand-alone): acc[stream][ear][bin] = sum_{s,p} FDL[stream][s][(head+p)%P][bin] * bank[s][p][bin].ear
// (ConvolutionEngine.swift:270-350 summed over speakers as RealtimeAudioProcessor.swift:146-163 does in the time
// domain).  Used when the fused kernel does not apply (B < 64 or B > 512).
// ------------------------------------------------------------------------------------------------
constexpr int kMacThreads = 128;

template <int T>
__global__ void __launch_bounds__(kMacThreads) k_fdl_cmac(const BlockGeom g, const float4 *__restrict__ fdl,
                                                           const float4 *__restrict__ bank, float4 *__restrict__ acc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int halfB = g.B >> 1;
    const int chunks = (halfB + kMacThreads - 1) / kMacThreads;
    const int tile = blockIdx.x / chunks, chunk = blockIdx.x - tile * chunks;
    const int jp = chunk * kMacThreads + threadIdx.x;
    if (jp >= halfB) return;   // no barrier below: the cp.async ring is private to each thread
    const int s0 = g.first_stream + tile * T;
    const int last = g.first_stream + g.n_streams - 1;
    const size_t stream_stride = (size_t)g.Se * g.P_cap * halfB;
    const float4 *fp[T];
#pragma unroll
    for (int u = 0; u < T; ++u) fp[u] = fdl + (size_t)min(s0 + u, last) * stream_stride + jp;
    float4 aL[T], aR[T];
#pragma unroll
    for (int u = 0; u < T; ++u) { aL[u] = make_float4(0.f, 0.f, 0.f, 0.f); aR[u] = aL[u]; }
    mac_phase<T, kMacThreads>(g, fp, bank + 2 * jp, reinterpret_cast<float4 *>(smem_raw), aL, aR, 0, g.P);
#pragma unroll
    for (int u = 0; u < T; ++u) {
        if (s0 + u <= last) {
            acc[((size_t)(s0 + u) * 2 + 0) * halfB + jp] = aL[u];
            acc[((size_t)(s0 + u) * 2 + 1) * halfB + jp] = aR[u];
        }
    }
}

cudaError_t launch_fdl_cmac(const BlockGeom &g, const float2 *fdl, const float4 *bank, float2 *acc, int tile, cudaStream_t st)
{
    const int halfB = g.B >> 1;
    const int chunks = (halfB + kMacThreads - 1) / kMacThreads;
    const int tiles = (g.n_streams + tile - 1) / tile;
    const int grid = tiles * chunks;
    if (grid <= 0) return cudaSuccess;
    const float4 *f4 = reinterpret_cast<const float4 *>(fdl);
    float4 *a4 = reinterpret_cast<float4 *>(acc);
    const size_t smem = (size_t)kMacStages * tile * kMacThreads * sizeof(float4);
    switch (tile) {
    case 1: k_fdl_cmac<1><<<grid, kMacThreads, smem, st>>>(g, f4, bank, a4); break;
    case 2: k_fdl_cmac<2><<<grid, kMacThreads, smem, st>>>(g, f4, bank, a4); break;
    case 4: k_fdl_cmac<4><<<grid, kMacThreads, smem, st>>>(g, f4, bank, a4); break;
    case 8: k_fdl_cmac<8><<<grid, kMacThreads, smem, st>>>(g, f4, bank, a4); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// KF  fused block kernel: K2 + K3 + K4 for T streams per CTA
// ------------------------------------------------------------------------------------------------
struct FusedArgs {
    BlockGeom g;
    StridedIn cur, prev;
    float *overlap_save;
    float2 *fdl;
    float *fdl_ny;
    const float4 *bank;
    const float *bank_ny;
    StridedOut out;
    const float2 *tw;
};

template <int LOG2M> struct FusedGeo {
    static constexpr int M = 1 << LOG2M;
    static constexpr int THREADS = M / 2;                // one bin pair per thread in the MAC phase
    static constexpr int G = RegFft<LOG2M>::G;
    static constexpr int NF = THREADS / G;               // = 8
    static constexpr int PS = PaddedSize<LOG2M>::value;
    static constexpr size_t fft_bytes = (size_t)NF * PS * sizeof(float2) + (size_t)THREADS * sizeof(float);
    template <int T> static constexpr size_t ring_bytes() { return (size_t)kMacStages * T * THREADS * sizeof(float4); }
    template <int T> static constexpr size_t smem() { return (size_t)M * sizeof(float2) + (fft_bytes > ring_bytes<T>() ? fft_bytes : ring_bytes<T>()); }
};

template <int LOG2M, int T, int MINB>
__global__ void __launch_bounds__(FusedGeo<LOG2M>::THREADS, MINB) k_fused(const FusedArgs a)
{
    using FG = FusedGeo<LOG2M>;
    constexpr int M = FG::M, G = FG::G, NF = FG::NF, THREADS = FG::THREADS, halfB = M / 2;
    static_assert(NF == 8 && 2 * T <= NF, "fused kernel geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *tw = reinterpret_cast<float2 *>(smem_raw);
    float2 *bufs = tw + M;
    float *part = reinterpret_cast<float *>(bufs + (size_t)NF * FG::PS);
    const BlockGeom &g = a.g;
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    const int s0 = g.first_stream + blockIdx.x * T;
    const int nvalid = min(T, g.first_stream + g.n_streams - s0);
    for (int k = tid; k < M; k += THREADS) tw[k] = a.tw[k];
    __syncthreads();

    // ---- phase A: forward FFT of the T*S input frames, 8 at a time; spectra go to the FDL head slot ----------
    const int nfft = T * g.S;
    for (int base = 0; base < nfft; base += NF) {
        const int idx = base + f;
        const int ls = idx / g.S, s = idx - ls * g.S;
        const bool active = idx < nfft && ls < nvalid;
        const int stream = s0 + (active ? ls : 0);
        const float *prev = a.prev.ptr + stream * a.prev.ss + s * a.prev.cs;
        const float *cur = a.cur.ptr + stream * a.cur.ss + s * a.cur.cs;
        float *ov = a.overlap_save ? a.overlap_save + ((size_t)stream * g.Se + s) * M : nullptr;
        const size_t row = ((size_t)stream * g.Se + s) * g.P_cap + g.head;
        float2 *dst = a.fdl + row * M;
        float *dst_ny = a.fdl_ny + row;
        forward_frame<LOG2M>(
            bufs + (size_t)f * FG::PS, tw, t, active,
            [&](int i, int) -> float2 {
                if (i < M / 2) return *reinterpret_cast<const float2 *>(prev + 2 * i);
                const float2 v = *reinterpret_cast<const float2 *>(cur + 2 * (i - M / 2));
                if (ov) *reinterpret_cast<float2 *>(ov + 2 * (i - M / 2)) = v;
                return v;
            },
            [&](int k, float2 x) { dst[k] = x; },
            [&](float ny) { *dst_ny = ny; });
    }
    // the head slot written above is re-read below by other threads of this CTA: make the global writes visible
    __threadfence_block();
    __syncthreads();

    // ---- phase B: FDL multiply-accumulate over speakers and partitions, accumulators in registers ------------
    const int jp = tid;   // bins 2*jp, 2*jp+1
    const size_t stream_stride = (size_t)g.Se * g.P_cap * halfB;
    const float4 *fdl4 = reinterpret_cast<const float4 *>(a.fdl);
    const float4 *fp[T];
#pragma unroll
    for (int u = 0; u < T; ++u) fp[u] = fdl4 + (size_t)(s0 + (u < nvalid ? u : 0)) * stream_stride + jp;
    float4 aL[T], aR[T];
#pragma unroll
    for (int u = 0; u < T; ++u) { aL[u] = make_float4(0.f, 0.f, 0.f, 0.f); aR[u] = aL[u]; }
    mac_phase<T, THREADS>(g, fp, a.bank + 2 * jp, reinterpret_cast<float4 *>(bufs), aL, aR, 0, g.P);
    __syncthreads();   // the ring aliases the FFT buffers

    // ---- phase C: accumulators -> shared, Nyquist sums, inverse FFT, overlap-save discard, output -------------
#pragma unroll
    for (int u = 0; u < T; ++u) {
        float2 *bl = bufs + (size_t)(2 * u) * FG::PS, *br = bl + FG::PS;
        bl[pad16(2 * jp)] = make_float2(aL[u].x, aL[u].y);
        bl[pad16(2 * jp + 1)] = make_float2(aL[u].z, aL[u].w);
        br[pad16(2 * jp)] = make_float2(aR[u].x, aR[u].y);
        br[pad16(2 * jp + 1)] = make_float2(aR[u].z, aR[u].w);
    }
    __syncthreads();
    {
        const int ls = f >> 1, ear = f & 1;
        const bool active = f < 2 * T && ls < nvalid;
        const int stream = s0 + (active ? ls : 0);
        const float ny = nyquist_sum<G>(g, a.fdl_ny, a.bank_ny, stream, ear, active, t, part + (size_t)f * G);
        float *row = a.out.ptr + stream * a.out.ss + ear * a.out.cs;
        inverse_frame<LOG2M>(bufs + (size_t)f * FG::PS, ny, tw, t, active,
                             [&](int i, float x0, float x1) { store_pair(a.out, row, 2 * i, x0, x1); });
    }
}

// ------------------------------------------------------------------------------------------------
// KP  persistent warp-specialised block kernel (64 <= B <= 512): one CTA per SM walks over tiles of 4 streams.
//
//   warp 0            producer: one elected thread streams the FDL *history* rows (partitions p >= 1) of tile after
//                     tile, plus the matching filter rows, into a deep shared-memory ring with TMA-class bulk copies
//                     (cp.async.bulk ... mbarrier::complete_tx).  It never waits for anything but a free ring stage, so
//                     HBM stays busy across tile boundaries and while other warps transform.
//   MAC warps (B/2 threads)   one bin pair of the 4 streams per thread; consume ring stages (full/empty mbarriers), then
//                     add the head partition (p = 0, just written by the FFT warps, read with plain loads).
//   FFT warps (B/2 threads)   while the MAC warps stream the history of tile i they run the inverse transforms of tile i-1
//                     (accumulators handed over through shared memory) and the forward transforms of tile i.
//
// The two latency-bound phases of KF (input FFT, inverse FFT) thereby run in the shadow of the bandwidth-bound phase.
// Named barriers: HEAD_READY (FFT -> MAC), ACC_READY (MAC -> FFT), ACC_FREE (FFT -> MAC).
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int LOG2M> struct PersistGeo {
    static constexpr int M = 1 << LOG2M, halfB = M / 2, T = 4;
    static constexpr int G = RegFft<LOG2M>::G;
    static constexpr int PRODUCERS = 4;                                     // producer warps, one issuing lane each
    static constexpr int MAC_GROUPS = 2, TG = T / MAC_GROUPS;               // two groups of MAC warps, 2 streams each
    static constexpr int FFT_GROUPS = 1, NFT = 8 * FFT_GROUPS;             // transforms the FFT warps run side by side
    static constexpr int MAC_THREADS = MAC_GROUPS * halfB, FFT_THREADS = FFT_GROUPS * halfB;
    static constexpr int THREADS = 32 * PRODUCERS + MAC_THREADS + FFT_THREADS;
    static constexpr int PS = PaddedSize<LOG2M>::value;
    static constexpr size_t stage_bytes = (size_t)T * halfB * sizeof(float4) + (size_t)M * sizeof(float4);   // T FDL rows + one filter row
    static constexpr size_t fixed_bytes = (size_t)M * sizeof(float2)               // twiddles
                                          + (size_t)(NFT + 8) * PS * sizeof(float2) // FFT buffers + accumulator hand-over buffers
                                          + 512;                                   // mbarriers (2 x STAGES x 8 B)
    static constexpr int max_stages = (int)((220 * 1024 - fixed_bytes) / stage_bytes);
    static constexpr int STAGES = max_stages > 24 ? 24 : max_stages;
    static constexpr size_t smem = fixed_bytes + (size_t)STAGES * stage_bytes;
};

struct PersistArgs {
    BlockGeom g;
    StridedIn cur, prev;
    float *overlap_save;
    float2 *fdl;
    float *fdl_ny;
    const float4 *bank;
    const float *bank_ny;
    StridedOut out;
    const float2 *tw;
    int debug;   // timing experiments only: bit 0 skips the forward transforms, bit 1 the inverse transforms
};

template <int LOG2M>
__global__ void __launch_bounds__(PersistGeo<LOG2M>::THREADS, 1) k_persistent(const PersistArgs a)
{
    using PG = PersistGeo<LOG2M>;
    constexpr int M = PG::M, halfB = PG::halfB, T = PG::T, TG = PG::TG, G = PG::G, STAGES = PG::STAGES, PS = PG::PS;
    constexpr int BAR_HEAD_READY = 1, BAR_ACC_READY = 2, BAR_ACC_FREE = 3, PAIR = PG::MAC_THREADS + PG::FFT_THREADS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);                                   // STAGES x (T*halfB + M) float4
    float2 *tw = reinterpret_cast<float2 *>(smem_raw + (size_t)STAGES * PG::stage_bytes);
    float2 *fftbuf = tw + M;
    float2 *accbuf = fftbuf + (size_t)PG::NFT * PS;
    uint64_t *full = reinterpret_cast<uint64_t *>(accbuf + (size_t)8 * PS);
    uint64_t *empty = full + STAGES;
    const BlockGeom &g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (g.n_streams + T - 1) / T;
    const int last = g.first_stream + g.n_streams - 1;
    const int old_iters = g.S * (g.P - 1);
    constexpr int stage_f4 = T * halfB + M;

    for (int k = tid; k < M; k += PG::THREADS) tw[k] = a.tw[k];
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], PG::MAC_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < PG::PRODUCERS) {
        // ===== producers: warp w issues the iterations k = w, w + PRODUCERS, ... of this CTA's iteration sequence =====
        if (lane == 0 && old_iters > 0) {
            const size_t stream_stride = (size_t)g.Se * g.P_cap * halfB;
            const float4 *fdl4 = reinterpret_cast<const float4 *>(a.fdl);
            const int pm1 = g.P - 1;
            long long k = warp;                       // global iteration index over (tile_local, s, p-1)
            int tile = blockIdx.x;
            int rem = warp;                           // iteration within the tile
            while (rem >= old_iters) { rem -= old_iters; tile += gridDim.x; }
            while (tile < n_tiles) {
                const int s = rem / pm1, p = 1 + (rem - s * pm1);
                int slot = g.head + p;
                if (slot >= g.P) slot -= g.P;         // modulus is partitionCount (Q4)
                const int stage = (int)(k % STAGES);
                const unsigned phase = (unsigned)((k / STAGES) & 1);
                const int s0 = g.first_stream + tile * T;
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_expect_tx(&full[stage], (unsigned)PG::stage_bytes);
                float4 *dst = ring + (size_t)stage * stage_f4;
#pragma unroll
                for (int u = 0; u < T; ++u)
                    bulk_g2s(dst + u * halfB, fdl4 + (size_t)min(s0 + u, last) * stream_stride + ((size_t)s * g.P_cap + slot) * halfB,
                             (unsigned)(halfB * sizeof(float4)), &full[stage]);
                bulk_g2s(dst + T * halfB, a.bank + ((size_t)s * g.P + p) * M, (unsigned)(M * sizeof(float4)), &full[stage]);
                k += PG::PRODUCERS;
                rem += PG::PRODUCERS;
                while (rem >= old_iters) { rem -= old_iters; tile += gridDim.x; }
            }
        }
    } else if (warp < PG::PRODUCERS + PG::MAC_THREADS / 32) {
        // ===== MAC warps: group gi owns streams 2*gi, 2*gi+1 of the tile; one bin pair per thread =====
        const int mt = tid - 32 * PG::PRODUCERS;
        const int gi = mt / halfB, jp = mt - gi * halfB;   // bins 2*jp, 2*jp+1
        int stage = 0;
        unsigned phase = 0;
        const size_t stream_stride = (size_t)g.Se * g.P_cap * halfB;
        const float4 *fdl4 = reinterpret_cast<const float4 *>(a.fdl);
        int local = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
            const int s0 = g.first_stream + tile * T + gi * TG;
            float4 aL[TG], aR[TG];
#pragma unroll
            for (int u = 0; u < TG; ++u) { aL[u] = make_float4(0.f, 0.f, 0.f, 0.f); aR[u] = aL[u]; }
            for (int it = 0; it < old_iters; ++it) {
                mbar_wait(&full[stage], phase);
                const float4 *src = ring + (size_t)stage * stage_f4;
                const float4 h0 = src[T * halfB + 2 * jp], h1 = src[T * halfB + 2 * jp + 1];
                float4 x[TG];
#pragma unroll
                for (int u = 0; u < TG; ++u) x[u] = src[(gi * TG + u) * halfB + jp];
#pragma unroll
                for (int u = 0; u < TG; ++u) {
                    cmac2f(aL[u], x[u], h0.x, h0.y, h1.x, h1.y);
                    cmac2f(aR[u], x[u], h0.z, h0.w, h1.z, h1.w);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            // head partition (p = 0): written by this CTA's FFT warps for this very block
            named_sync(BAR_HEAD_READY, PAIR);
            for (int s = 0; s < g.S; ++s) {
                const float4 h0 = __ldg(a.bank + (size_t)s * g.P * M + 2 * jp), h1 = __ldg(a.bank + (size_t)s * g.P * M + 2 * jp + 1);
                const size_t off = ((size_t)s * g.P_cap + g.head) * halfB + jp;
                float4 x[TG];
#pragma unroll
                for (int u = 0; u < TG; ++u) x[u] = *(fdl4 + (size_t)min(s0 + u, last) * stream_stride + off);
#pragma unroll
                for (int u = 0; u < TG; ++u) {
                    cmac2f(aL[u], x[u], h0.x, h0.y, h1.x, h1.y);
                    cmac2f(aR[u], x[u], h0.z, h0.w, h1.z, h1.w);
                }
            }
            if (local > 0) named_sync(BAR_ACC_FREE, PAIR);   // the FFT warps are done with the previous accumulators
#pragma unroll
            for (int u = 0; u < TG; ++u) {
                float2 *bl = accbuf + (size_t)(2 * (gi * TG + u)) * PS, *br = bl + PS;
                bl[pad16(2 * jp)] = make_float2(aL[u].x, aL[u].y);
                bl[pad16(2 * jp + 1)] = make_float2(aL[u].z, aL[u].w);
                br[pad16(2 * jp)] = make_float2(aR[u].x, aR[u].y);
                br[pad16(2 * jp + 1)] = make_float2(aR[u].z, aR[u].w);
            }
            __threadfence_block();
            named_arrive(BAR_ACC_READY, PAIR);
        }
    } else {
        // ===== FFT warps =====
        const int ft = tid - 32 * PG::PRODUCERS - PG::MAC_THREADS;
        const int f = ft / G, t = ft % G;
        auto inverse_tile = [&](int tile) {
            const int s0 = g.first_stream + tile * T;
            const int nvalid = min(T, g.first_stream + g.n_streams - s0);
            if (f >= 8) return;                 // 2*T = 8 inverse transforms; transforms never share a warp with f < 8 (8*G >= 32)
            const int ls = f >> 1, ear = f & 1;
            const bool active = ls < nvalid;
            const int stream = s0 + (active ? ls : 0);
            const float ny = nyquist_sum<G>(g, a.fdl_ny, a.bank_ny, stream, ear, active, t, nullptr);
            float *row = a.out.ptr + stream * a.out.ss + ear * a.out.cs;
            inverse_frame<LOG2M>(accbuf + (size_t)f * PS, ny, tw, t, active,
                                 [&](int i, float x0, float x1) { store_pair(a.out, row, 2 * i, x0, x1); });
        };
        int local = 0, prev_tile = -1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local) {
            if (local > 0) {
                named_sync(BAR_ACC_READY, PAIR);
                if (!(a.debug & 2)) inverse_tile(prev_tile);
                __syncwarp();
                named_arrive(BAR_ACC_FREE, PAIR);
            }
            const int s0 = g.first_stream + tile * T;
            const int nvalid = min(T, g.first_stream + g.n_streams - s0);
            const int nfft = T * g.S;
            using F = RegFft<LOG2M>;
            // round r transforms frames r*8 + f; the operands of round r+1 are fetched before round r is transformed
            auto fetch = [&](int base, float2 (&v)[F::E], bool &active, int &stream, int &s) {
                const int idx = base + f;
                const int ls = idx / g.S;
                s = idx - ls * g.S;
                active = idx < nfft && ls < nvalid;
                stream = s0 + (active ? ls : 0);
                const float *prev = a.prev.ptr + stream * a.prev.ss + s * a.prev.cs;
                const float *cur = a.cur.ptr + stream * a.cur.ss + s * a.cur.cs;
#pragma unroll
                for (int e = 0; e < F::E; ++e) {
                    const int i = F::template load_index<0>(t, e);
                    v[e] = !active ? make_float2(0.f, 0.f)
                                   : (i < M / 2 ? *reinterpret_cast<const float2 *>(prev + 2 * i) : *reinterpret_cast<const float2 *>(cur + 2 * (i - M / 2)));
                }
            };
            float2 v[F::E], vn[F::E];
            bool active = false, active_n = false;
            int stream = 0, sp = 0, stream_n = 0, sp_n = 0;
            const int base0 = (a.debug & 1) ? nfft : 0;
            if (base0 < nfft) fetch(base0, v, active, stream, sp);
            for (int base = base0; base < nfft; base += PG::NFT) {
                if (base + PG::NFT < nfft) fetch(base + PG::NFT, vn, active_n, stream_n, sp_n);
                if (active && a.overlap_save) {   // inputOverlapBuffer <- current block (the same thread read these addresses as `prev`)
                    float *ov = a.overlap_save + ((size_t)stream * g.Se + sp) * M;
#pragma unroll
                    for (int e = 0; e < F::E; ++e) {
                        const int i = F::template load_index<0>(t, e);
                        if (i >= M / 2) *reinterpret_cast<float2 *>(ov + 2 * (i - M / 2)) = v[e];
                    }
                }
                const size_t row = ((size_t)stream * g.Se + sp) * g.P_cap + g.head;
                float2 *dst = a.fdl + row * M;
                float *dst_ny = a.fdl_ny + row;
                forward_frame_regs<LOG2M>(fftbuf + (size_t)f * PS, tw, t, active, v,
                                          [&](int k, float2 x) { dst[k] = x; }, [&](float ny) { *dst_ny = ny; });
#pragma unroll
                for (int e = 0; e < F::E; ++e) v[e] = vn[e];
                active = active_n; stream = stream_n; sp = sp_n;
            }
            __threadfence_block();
            named_arrive(BAR_HEAD_READY, PAIR);
            prev_tile = tile;
        }
        if (prev_tile >= 0) {
            named_sync(BAR_ACC_READY, PAIR);
            if (!(a.debug & 2)) inverse_tile(prev_tile);
        }
    }
}

template <int LOG2M>
cudaError_t launch_persistent_l(const PersistArgs &a, int ctas, cudaStream_t st)
{
    static bool configured = false;   // opt in to the large dynamic shared memory once per process
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_persistent<LOG2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PersistGeo<LOG2M>::smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_persistent<LOG2M><<<ctas, PersistGeo<LOG2M>::THREADS, PersistGeo<LOG2M>::smem, st>>>(a);
    return cudaGetLastError();
}

bool persistent_supported(int log2m, int P) { return log2m >= 6 && log2m <= 9 && P >= 1; }

cudaError_t launch_persistent(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl, float *fdl_ny,
                              const float4 *bank, const float *bank_ny, StridedOut out, const float2 *tw, int num_sms, int debug,
                              cudaStream_t st)
{
    if (g.n_streams <= 0) return cudaSuccess;
    PersistArgs a{g, cur, prev, overlap_save, fdl, fdl_ny, bank, bank_ny, out, tw, debug};
    const int tiles = (g.n_streams + 3) / 4;
    const int ctas = tiles < num_sms ? tiles : num_sms;
    switch (g.log2m) {
    case 6: return launch_persistent_l<6>(a, ctas, st);
    case 7: return launch_persistent_l<7>(a, ctas, st);
    case 8: return launch_persistent_l<8>(a, ctas, st);
    case 9: return launch_persistent_l<9>(a, ctas, st);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
#define AW_LOG2M_SWITCH(log2m, CALL)                                            \
    switch (log2m) {                                                            \
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   \
    case 5: CALL(5); break;   case 6: CALL(6); break;   case 7: CALL(7); break;   \
    case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break; \
    case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break; \
    default: return cudaErrorInvalidValue;                                      \
    }

cudaError_t launch_input_rfft(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl,
                              float *fdl_ny, const float2 *tw, cudaStream_t st)
{
    const int jobs = g.n_streams * g.S;
    if (jobs <= 0) return cudaSuccess;
    InputRfftArgs a{g, cur, prev, overlap_save, fdl, fdl_ny, tw};
#define CALL(L) k_input_rfft<L><<<(jobs + Geo<L>::SA_NF - 1) / Geo<L>::SA_NF, Geo<L>::SA_THREADS, Geo<L>::sa_smem, st>>>(a)
    AW_LOG2M_SWITCH(g.log2m, CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_irfft_out(const BlockGeom &g, const float2 *acc, const float *fdl_ny, const float *bank_ny,
                             StridedOut out, const float2 *tw, cudaStream_t st)
{
    const int jobs = g.n_streams * 2;
    if (jobs <= 0) return cudaSuccess;
    IrfftArgs a{g, acc, fdl_ny, bank_ny, out, tw};
#define CALL(L) k_irfft_out<L><<<(jobs + Geo<L>::SA_NF - 1) / Geo<L>::SA_NF, Geo<L>::SA_THREADS, Geo<L>::sa_smem, st>>>(a)
    AW_LOG2M_SWITCH(g.log2m, CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_bank_build(const float *ir, int S, int taps, int B, int log2m, int P, float4 *bank, float *bank_ny,
                              const float2 *tw, cudaStream_t st)
{
    const int jobs = S * 2 * P;
    BankArgs a{ir, S, taps, B, P, bank, bank_ny, tw};
#define CALL(L) k_bank_build<L><<<(jobs + Geo<L>::SA_NF - 1) / Geo<L>::SA_NF, Geo<L>::SA_THREADS, Geo<L>::sa_smem, st>>>(a)
    AW_LOG2M_SWITCH(log2m, CALL)
#undef CALL
    return cudaGetLastError();
}

// ---- fused kernel: variants and planner ------------------------------------------------------------
namespace {

template <int LOG2M, int T, int MINB>
cudaError_t launch_fused_t(const FusedArgs &a, int tiles, cudaStream_t st)
{
    k_fused<LOG2M, T, MINB><<<tiles, FusedGeo<LOG2M>::THREADS, FusedGeo<LOG2M>::template smem<T>(), st>>>(a);
    return cudaGetLastError();
}

// occupancy targets (CTAs/SM) the variants are compiled for: threads/CTA = B/2
template <int LOG2M> struct FusedOcc;
template <> struct FusedOcc<6> { static constexpr int t4 = 16, t2 = 24, t1 = 32; };   // 32 threads
template <> struct FusedOcc<7> { static constexpr int t4 = 12, t2 = 16, t1 = 16; };   // 64 threads
template <> struct FusedOcc<8> { static constexpr int t4 = 7, t2 = 9, t1 = 9; };      // 128 threads
template <> struct FusedOcc<9> { static constexpr int t4 = 3, t2 = 4, t1 = 4; };      // 256 threads

template <int LOG2M>
cudaError_t launch_fused_l(const FusedArgs &a, int tile, cudaStream_t st)
{
    const int tiles = (a.g.n_streams + tile - 1) / tile;
    switch (tile) {
    case 4: return launch_fused_t<LOG2M, 4, FusedOcc<LOG2M>::t4>(a, tiles, st);
    case 2: return launch_fused_t<LOG2M, 2, FusedOcc<LOG2M>::t2>(a, tiles, st);
    case 1: return launch_fused_t<LOG2M, 1, FusedOcc<LOG2M>::t1>(a, tiles, st);
    default: return cudaErrorInvalidValue;
    }
}

template <int LOG2M, int T, int MINB>
int fused_blocks_per_sm()
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fused<LOG2M, T, MINB>, FusedGeo<LOG2M>::THREADS, FusedGeo<LOG2M>::template smem<T>()) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

template <int LOG2M>
int fused_blocks_per_sm_l(int tile)
{
    switch (tile) {
    case 4: return fused_blocks_per_sm<LOG2M, 4, FusedOcc<LOG2M>::t4>();
    case 2: return fused_blocks_per_sm<LOG2M, 2, FusedOcc<LOG2M>::t2>();
    case 1: return fused_blocks_per_sm<LOG2M, 1, FusedOcc<LOG2M>::t1>();
    default: return 0;
    }
}

}  // namespace

bool fused_supported(int log2m) { return log2m >= 6 && log2m <= 9; }

int fused_blocks_per_sm(int log2m, int tile)
{
    switch (log2m) {
    case 6: return fused_blocks_per_sm_l<6>(tile);
    case 7: return fused_blocks_per_sm_l<7>(tile);
    case 8: return fused_blocks_per_sm_l<8>(tile);
    case 9: return fused_blocks_per_sm_l<9>(tile);
    default: return 0;
    }
}

cudaError_t launch_fused(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl, float *fdl_ny,
                         const float4 *bank, const float *bank_ny, StridedOut out, const float2 *tw, int tile, cudaStream_t st)
{
    if (g.n_streams <= 0) return cudaSuccess;
    FusedArgs a{g, cur, prev, overlap_save, fdl, fdl_ny, bank, bank_ny, out, tw};
    switch (g.log2m) {
    case 6: return launch_fused_l<6>(a, tile, st);
    case 7: return launch_fused_l<7>(a, tile, st);
    case 8: return launch_fused_l<8>(a, tile, st);
    case 9: return launch_fused_l<9>(a, tile, st);
    default: return cudaErrorInvalidValue;
    }
}

size_t fft_smem_bytes(int log2m)
{
    size_t b = 0;
#define CALL(L) b = Geo<L>::sa_smem
    switch (log2m) {
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   case 5: CALL(5); break;
    case 6: CALL(6); break;   case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;
    case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break;
    default: break;
    }
#undef CALL
    return b;
}

cudaError_t configure_kernels(int log2m)
{
    cudaError_t e = cudaSuccess;
#define CALL(L)                                                                                                              \
    do {                                                                                                                     \
        if (Geo<L>::sa_smem > 48 * 1024) {                                                                                   \
            if ((e = cudaFuncSetAttribute(k_input_rfft<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<L>::sa_smem)) != cudaSuccess) return e; \
            if ((e = cudaFuncSetAttribute(k_irfft_out<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<L>::sa_smem)) != cudaSuccess) return e;  \
            if ((e = cudaFuncSetAttribute(k_bank_build<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<L>::sa_smem)) != cudaSuccess) return e; \
        }                                                                                                                    \
    } while (0)
    AW_LOG2M_SWITCH(log2m, CALL)
#undef CALL
    return e;
}

}  // namespace aw
