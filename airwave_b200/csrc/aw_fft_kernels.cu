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
#include "aw_fft_blocks.cuh"
#include <cstdlib>

// How K2 / K4 overlap the fetch of their next frame with the transform of the current one (block sizes where a CTA loops, B >= 2048):
// 0 = not at all; 1 = prefetch.global.L2 of the next frame's lines; 2 = K2 only: pass-0 operands of the next frame in registers
// (measured slower: the registers cost a resident CTA).
#ifndef AW_K2_PIPE
#define AW_K2_PIPE 1
#endif
#ifndef AW_SA_RESIDENT_THREADS
#define AW_SA_RESIDENT_THREADS 768
#endif
#ifndef AW_K4_PIPE
#define AW_K4_PIPE 1
#endif

namespace aw {


template <int LOG2M>
struct Geo {
    using F = RegFft<LOG2M>;
    static constexpr int M = F::M, G = F::G;
    static constexpr int SA_THREADS = G > 128 ? G : 128;      // stand-alone K1/K2/K4
    static constexpr int SA_NF = SA_THREADS / G;
    static constexpr int PS = PaddedSize<LOG2M>::value;
    // K1 keeps the half-circle twiddle table in shared memory; K2/K4 read the per-pass tables (PT layout, conflict-free and
    // coalesced: k contiguous) straight from the plan in global memory — L1 serves them — so that a CTA needs nothing but its
    // transform buffers and 3x as many CTAs fit on an SM at B = 4096
    static constexpr size_t k1_smem = (size_t)(M + SA_NF * PS) * sizeof(float2) + (size_t)SA_THREADS * sizeof(float);
    static constexpr size_t sa_smem = (size_t)(SA_NF * PS) * sizeof(float2) + (size_t)SA_THREADS * sizeof(float);
    // from B = 2048 a CTA of K2/K4 loops over several frames (sa_grid): it fetches the next frame while it transforms this one
    static constexpr bool LOOPS = LOG2M >= 11;
    // resident CTAs of K2 / K4 the register budget is held to (80 registers: 768 threads)
    static constexpr int MIN_CTAS = AW_SA_RESIDENT_THREADS / SA_THREADS > 0 ? AW_SA_RESIDENT_THREADS / SA_THREADS : 1;
    static constexpr size_t k4_smem = sa_smem;
};

// the lines of the operands of `job` of K2 -> L2 (the loads of the next round then wait for L2, not DRAM)
template <int LOG2M>
__device__ __forceinline__ void k2_prefetch(const BlockGeom &g, const StridedIn &prev, const StridedIn &cur, int job, int jobs, int t)
{
    constexpr int M = 1 << LOG2M, G = RegFft<LOG2M>::G, LINES = M / 32;   // 128-byte lines in one block of M floats
    if (job >= jobs) return;
    const int ls = job / g.S, s = job - ls * g.S, stream = g.first_stream + ls;
    const RowSources rs = row_sources(g, s);
    for (int line = t; line < 2 * LINES; line += G) {
        const bool second = line >= LINES;
        const int off = (second ? line - LINES : line) * 32;
        if (!second && g.prev_is_rows) { prefetch_l2(prev.ptr + stream * prev.ss + s * prev.cs + off); continue; }
        const StridedIn &in = second ? cur : prev;
#pragma unroll
        for (int q = 0; q < kKpMaxRowSources; ++q)
            if (rs.c[q] >= 0) prefetch_l2(in.ptr + stream * in.ss + rs.c[q] * in.cs + off);
    }
}

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
__global__ void __launch_bounds__(Geo<LOG2M>::SA_THREADS, Geo<LOG2M>::MIN_CTAS) k_input_rfft(const InputRfftArgs a)
{
    using Gm = Geo<LOG2M>;
    constexpr int M = Gm::M, G = Gm::G, NF = Gm::SA_NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *bufs = reinterpret_cast<float2 *>(smem_raw);
    const float2 *tw = a.tw + M;          // the plan's per-pass twiddle tables (global memory, read through L1)
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    const int jobs = a.g.n_streams * a.g.S;
    using F = RegFft<LOG2M>;
    // pass-0 operands of `job`: frame = [previous block | current block] (:237-248); inputOverlapBuffer <- current block (:243)
    auto fetch = [&](int job, float2 (&v)[F::E]) {
        const bool active = job < jobs;
        const int ls = active ? job / a.g.S : 0, s = active ? job - ls * a.g.S : 0;
        const int stream = a.g.first_stream + ls;
        float *ov = a.overlap_save ? a.overlap_save + ((size_t)stream * a.g.Se + s) * M : nullptr;
        fetch_frame<LOG2M>(a.g, a.prev, a.cur, stream, s, ov, t, active, v);
    };
    float2 v[F::E];
    if constexpr (Gm::LOOPS && AW_K2_PIPE == 2) fetch(blockIdx.x * NF + f, v);
    for (int base = blockIdx.x * NF; base < jobs; base += gridDim.x * NF) {   // uniform trip count: barriers inside
        const int job = base + f;
        const bool active = job < jobs;
        const int ls = active ? job / a.g.S : 0, s = active ? job - ls * a.g.S : 0;
        const int stream = a.g.first_stream + ls;
        const size_t row = ((size_t)stream * a.g.Se + s) * a.g.P_cap + a.g.head;
        float2 *dst = a.fdl + row * M;
        float *dst_ny = a.fdl_ny + row;
        float2 nv[F::E];
        if constexpr (Gm::LOOPS && AW_K2_PIPE == 2) {
            fetch(job + gridDim.x * NF, nv);               // in flight while this frame is transformed
        } else {
            if constexpr (Gm::LOOPS && AW_K2_PIPE == 1) k2_prefetch<LOG2M>(a.g, a.prev, a.cur, job + gridDim.x * NF, jobs, t);
            fetch(job, v);
        }
        forward_frame_regs<LOG2M, 2>(
            bufs + (size_t)f * Gm::PS, tw, t, active, v,
            [&](int k, float2 x) { dst[k] = x; },       // FDL[head] <- spectrum (:256-264)
            [&](float ny) { *dst_ny = ny; });
        if constexpr (Gm::LOOPS && AW_K2_PIPE == 2) {
#pragma unroll
            for (int e = 0; e < F::E; ++e) v[e] = nv[e];
        }
    }
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
    // bank row (s, p): two planes of B/2 float4 {L.re, L.im, R.re, R.im} — even bins, then odd bins — so that the thread owning
    // bins (2j, 2j+1) reads plane0[j] and plane1[j]: coalesced in HBM and conflict-free in shared memory
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
        [&](int k, float2 x) {
            float *d = dst + 4 * ((k & 1) * (M / 2) + (k >> 1));
            d[0] = x.x * scale; d[1] = x.y * scale;
        },
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


template <int LOG2M>
__global__ void __launch_bounds__(Geo<LOG2M>::SA_THREADS, Geo<LOG2M>::MIN_CTAS) k_irfft_out(const IrfftArgs a)
{
    using Gm = Geo<LOG2M>;
    constexpr int M = Gm::M, G = Gm::G, NF = Gm::SA_NF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *bufs = reinterpret_cast<float2 *>(smem_raw);
    const float2 *tw = a.tw + M;          // the plan's per-pass twiddle tables (global memory, read through L1)
    float *part = reinterpret_cast<float *>(bufs + (size_t)NF * Gm::PS);
    const int tid = threadIdx.x, f = tid / G, t = tid % G;
    float2 *buf = bufs + (size_t)f * Gm::PS;
    const int jobs = a.g.n_streams * 2;   // (local stream, ear)
    for (int base = blockIdx.x * NF; base < jobs; base += gridDim.x * NF) {   // uniform trip count: barriers inside
        const int job = base + f;
        const bool active = job < jobs;
        const int stream = a.g.first_stream + (active ? job >> 1 : 0), ear = job & 1;
        const float ny_part = nyquist_partial<G>(a.g, a.fdl_ny, a.bank_ny, stream, ear, active, t);   // in flight with the copy below
        __syncthreads();                  // the previous round is done with buf and part
        if constexpr (Gm::LOOPS && AW_K4_PIPE == 1) {
            const int next = job + gridDim.x * NF;
            if (next < jobs) {
                const float2 *src = a.acc + ((size_t)(a.g.first_stream + (next >> 1)) * 2 + (next & 1)) * M;
                for (int line = t; line < M / 16; line += G) prefetch_l2(src + line * 16);
            }
        }
        if constexpr (G > 32) part[(size_t)f * G + t] = ny_part;
        float *row = a.out.ptr + stream * a.out.ss + ear * a.out.cs;
        inverse_frame_global<LOG2M, 2>(
            buf, a.acc + ((size_t)stream * 2 + ear) * M, tw, t, active,
            // the Nyquist sum, added up in nyquist_sum's order; only thread 0 of a transform uses it (bin 0 of the inverse split)
            [&]() -> float {
                if constexpr (G <= 32) return group_sum(ny_part, G);
                else return t < 32 ? nyquist_reduce_warp0<G>(part + (size_t)f * G, t) : 0.f;
            },
            [&](int i, float x0, float x1) { store_pair(a.out, row, 2 * i, x0, x1); });
    }
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

constexpr int kMacStages = 3;

// ring: kMacStages * T * THREADS float4.  fp[u]: FDL base of stream u at this thread's bin pair (float4 units);
// bank_jp: bank + jp.  Accumulates partitions p in [p0, p1) of every speaker into aL/aR (left/right ear, 2 bins each).
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
    float4 h0 = __ldg(bk), h1 = __ldg(bk + halfB);   // even-bin plane, odd-bin plane
    int cstage = 0, cp = 0;
    for (int it = 0; it < total; ++it) {
        issue();
        bk += g.B;
        if (++cp == np) { cp = 0; bk += skip; }
        float4 n0 = h0, n1 = h1;
        if (it + 1 < total) { n0 = __ldg(bk); n1 = __ldg(bk + halfB); }
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
    mac_phase<T, kMacThreads>(g, fp, bank + jp, reinterpret_cast<float4 *>(smem_raw), aL, aR, 0, g.P);
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
        float *ov = a.overlap_save ? a.overlap_save + ((size_t)stream * g.Se + s) * M : nullptr;
        const size_t row = ((size_t)stream * g.Se + s) * g.P_cap + g.head;
        float2 *dst = a.fdl + row * M;
        float *dst_ny = a.fdl_ny + row;
        float2 v[RegFft<LOG2M>::E];
        fetch_frame<LOG2M>(g, a.prev, a.cur, stream, s, ov, t, active, v);
        forward_frame_regs<LOG2M>(
            bufs + (size_t)f * FG::PS, tw, t, active, v,
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
    mac_phase<T, THREADS>(g, fp, a.bank + jp, reinterpret_cast<float4 *>(bufs), aL, aR, 0, g.P);
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

// grid of the stand-alone transform kernels: every CTA loops over its share of the frames, so no more CTAs than can be resident
// (227 KB of shared memory per SM, 148 SMs) — the twiddle tables are then built once per resident CTA instead of once per frame
template <int LOG2M>
static int sa_grid(int jobs, size_t smem, const void *kernel, int which)
{
    const int ctas = (jobs + Geo<LOG2M>::SA_NF - 1) / Geo<LOG2M>::SA_NF;
    if (LOG2M <= 10) return ctas;   // measured: up to B = 1024 one frame group per CTA is faster (the CTA scheduler overlaps them)
    static int resident[2] = {0, 0};   // [K2, K4] of this block size: SMs x the CTAs of the kernel that fit on one
    int &r = resident[which];
    if (r == 0) {
        int per_sm = 0, dev = 0, sms = 148;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, Geo<LOG2M>::SA_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int waves = 2;   // measured at B = 4096: two CTAs per resident slot, each looping over half as many frames, hide the tail better
        if (const char *v = getenv("AW_SA_WAVES")) waves = atoi(v) > 0 ? atoi(v) : 1;
        r = sms * per_sm * waves;
    }
    return ctas < r ? ctas : r;
}

cudaError_t launch_input_rfft(const BlockGeom &g, StridedIn cur, StridedIn prev, float *overlap_save, float2 *fdl,
                              float *fdl_ny, const float2 *tw, cudaStream_t st)
{
    const int jobs = g.n_streams * g.S;
    if (jobs <= 0) return cudaSuccess;
    InputRfftArgs a{g, cur, prev, overlap_save, fdl, fdl_ny, tw};
#define CALL(L) k_input_rfft<L><<<sa_grid<L>(jobs, Geo<L>::sa_smem, (const void *)k_input_rfft<L>, 0), Geo<L>::SA_THREADS, Geo<L>::sa_smem, st>>>(a)
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
#define CALL(L) k_irfft_out<L><<<sa_grid<L>(jobs, Geo<L>::k4_smem, (const void *)k_irfft_out<L>, 1), Geo<L>::SA_THREADS, Geo<L>::k4_smem, st>>>(a)
    AW_LOG2M_SWITCH(g.log2m, CALL)
#undef CALL
    return cudaGetLastError();
}

cudaError_t launch_bank_build(const float *ir, int S, int taps, int B, int log2m, int P, float4 *bank, float *bank_ny,
                              const float2 *tw, cudaStream_t st)
{
    const int jobs = S * 2 * P;
    BankArgs a{ir, S, taps, B, P, bank, bank_ny, tw};
#define CALL(L) k_bank_build<L><<<(jobs + Geo<L>::SA_NF - 1) / Geo<L>::SA_NF, Geo<L>::SA_THREADS, Geo<L>::k1_smem, st>>>(a)
    AW_LOG2M_SWITCH(log2m, CALL)
#undef CALL
    return cudaGetLastError();
}

// ---- plan: per-pass twiddle tables appended to the half-circle table ---------------------------------
template <int LOG2M>
__global__ void k_build_pt(const float2 *hc, float2 *pt)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pt_total_entries(LOG2M); i += gridDim.x * blockDim.x)
        pt[i] = RegFft<LOG2M>::pt_entry(hc, i);
}

int plan_pt_entries(int log2m)
{
    int n = 0;
#define CALL(L) n = pt_total_entries(L)
    switch (log2m) {
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;   case 5: CALL(5); break;
    case 6: CALL(6); break;   case 7: CALL(7); break;   case 8: CALL(8); break;   case 9: CALL(9); break;
    case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; case 13: CALL(13); break;
    default: break;
    }
#undef CALL
    return n;
}

cudaError_t launch_build_pt(int log2m, const float2 *hc, float2 *pt, cudaStream_t st)
{
#define CALL(L) k_build_pt<L><<<8, 256, 0, st>>>(hc, pt)
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
        }                                                                                                                    \
        if (Geo<L>::k4_smem > 48 * 1024) {                                                                                   \
            if ((e = cudaFuncSetAttribute(k_irfft_out<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<L>::k4_smem)) != cudaSuccess) return e;  \
        }                                                                                                                    \
        if (Geo<L>::k1_smem > 48 * 1024) {                                                                                   \
            if ((e = cudaFuncSetAttribute(k_bank_build<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Geo<L>::k1_smem)) != cudaSuccess) return e; \
        }                                                                                                                    \
    } while (0)
    AW_LOG2M_SWITCH(log2m, CALL)
#undef CALL
    return e;
}

}  // namespace aw
