// aw_kernels.cu — hand-written sm_100a kernels of the batched binaural renderer.
//
//   K1 k_bank_build      ConvolutionEngine.init partition FFTs            (ConvolutionEngine.swift:143-182)
//   K2 k_input_rfft      overlap-save frame + forward FFT + FDL write     (ConvolutionEngine.swift:237-264)
//   K3 k_fdl_cmac        frequency-domain delay-line multiply-accumulate  (ConvolutionEngine.swift:270-350,
//                        summed over speakers as RealtimeAudioProcessor.swift:146-163 does in time domain)
//   K4 k_irfft_out       inverse FFT, scale, keep second half             (ConvolutionEngine.swift:353-366)
//   K5 k_eq              float64 biquad cascade + 20 ms crossfade         (ParametricEqualizerProcessor.swift:58-91, 254-314)
//   K6 k_resample_vgenp  Resampler.resampleHighQuality                    (Resampler.swift:31-68)
//   K7 k_gather_pending / k_drain_fifo  frame adapter                     (RealtimeAudioProcessor.swift:88-116, 166-190)
//
// Layouts in HBM (B = block = bins per spectrum, P_cap = FDL slots per (stream, speaker)):
//   FDL     float2 [stream][S][P_cap][B]      bin 0 = (2*DC, 0); Nyquist kept aside so the MAC is uniformly complex
//   FDL_ny  float  [stream][S][P_cap]
//   bank    float4 [S][P][B] = {L.re, L.im, R.re, R.im}; bank_ny float [S][P][2]
//   acc     float2 [stream][2][B]
#include "aw_kernels.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>


namespace aw {

// ------------------------------------------------------------------------------------------------
// K6  resample (vDSP_vramp + vDSP_vgenp semantics).  Bit-exact with the oracle:
// explicit round-to-nearest mul/add so nvcc cannot contract them into FMAs.
// ------------------------------------------------------------------------------------------------
__global__ void k_resample_vgenp(const float *__restrict__ in, int rows, int count, float step, float *__restrict__ out, int out_count)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (n >= out_count || row >= rows) return;
    const float *x = in + (size_t)row * count;
    const int M = count;
    const float fn = (float)n;
    const float bLast = __fmul_rn((float)(M - 1), step);
    float r;
    if (fn <= 0.0f) r = x[0];                       // n <= trunc(B[0]) = 0
    else if (fn > truncf(bLast)) r = x[M - 1];      // beyond the last breakpoint: hold
    else {
        int m = (int)(fn / step);
        if (m > M - 2) m = M - 2;
        if (m < 0) m = 0;
        while (m + 1 < M && truncf(__fmul_rn((float)(m + 1), step)) < fn) ++m;   // largest m with trunc(B[m]) < n
        while (m > 0 && !(truncf(__fmul_rn((float)m, step)) < fn)) --m;
        const float bm = __fmul_rn((float)m, step), bm1 = __fmul_rn((float)(m + 1), step);
        // reference order: A[m] + (A[m+1]-A[m]) * (n - B[m]) / (B[m+1]-B[m]); the oracle evaluates (d*(n-bm))/(bm1-bm)
        const float d = __fsub_rn(x[m + 1], x[m]);
        r = __fadd_rn(x[m], __fdiv_rn(__fmul_rn(d, __fsub_rn(fn, bm)), __fsub_rn(bm1, bm)));
    }
    out[(size_t)row * out_count + n] = r;
}

cudaError_t launch_resample_vgenp(const float *in, int rows, int count, float step, float *out, int out_count, cudaStream_t st)
{
    if (rows <= 0 || out_count <= 0) return cudaSuccess;
    dim3 grid((out_count + 255) / 256, rows);
    k_resample_vgenp<<<grid, 256, 0, st>>>(in, rows, count, step, out, out_count);
    return cudaGetLastError();
}

// K6b  AW_RESAMPLE_CORRECT: what Resampler.swift:16-30 documents ("linear interpolation ... at the target rate") rather than
// what its vgenp call computes.  Source position n * fromRate / toRate in float64, explicit round-to-nearest operations so
// that the numpy float64 restatement in oracle/binding.py (resample_linear_f64) matches bit for bit.
__global__ void k_resample_linear(const float *__restrict__ in, int rows, int count, double step, float *__restrict__ out, int out_count)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (n >= out_count || row >= rows) return;
    const float *x = in + (size_t)row * count;
    const double pos = __dmul_rn((double)n, step);
    float r;
    if (pos >= (double)(count - 1)) r = x[count - 1];          // past the last sample: hold it
    else {
        const int m = (int)pos;                                  // pos >= 0: truncation = floor
        const double frac = __dsub_rn(pos, (double)m);
        const double a = (double)x[m], d = __dsub_rn((double)x[m + 1], a);
        r = (float)__dadd_rn(a, __dmul_rn(d, frac));
    }
    out[(size_t)row * out_count + n] = r;
}

cudaError_t launch_resample_linear(const float *in, int rows, int count, double step, float *out, int out_count, cudaStream_t st)
{
    if (rows <= 0 || out_count <= 0) return cudaSuccess;
    dim3 grid((out_count + 255) / 256, rows);
    k_resample_linear<<<grid, 256, 0, st>>>(in, rows, count, step, out, out_count);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K7  frame adapter
// ------------------------------------------------------------------------------------------------
__global__ void k_gather_pending(StridedIn in, int in_offset, int copy_count, float *pending, int pending_count, int n_streams,
                                 int S, int B, int dup_mono)
{
    const long long total = (long long)n_streams * S * copy_count;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % copy_count);
        const long long row = idx / copy_count;
        const int s = (int)(row % S);
        const long long stream = row / S;
        const int src_ch = (dup_mono && s == 1) ? 0 : s;   // nil right input: duplicate left (:101-107)
        pending[(stream * S + s) * B + pending_count + i] = in.ptr[stream * in.ss + src_ch * in.cs + in_offset + i];
    }
}

cudaError_t launch_gather_pending(StridedIn in, int in_offset, int copy_count, float *pending, int pending_count, int n_streams,
                                  int S, int B, int dup_mono, cudaStream_t st)
{
    const long long total = (long long)n_streams * S * copy_count;
    if (total <= 0) return cudaSuccess;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_gather_pending<<<grid, 256, 0, st>>>(in, in_offset, copy_count, pending, pending_count, n_streams, S, B, dup_mono);
    return cudaGetLastError();
}

__global__ void k_drain_fifo(const float *__restrict__ fifo, int fifo_cap, int fifo_read, int fifo_count, StridedOut out,
                             int out_offset, int frames, int n_streams)
{
    const long long total = (long long)n_streams * 2 * frames;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % frames);
        const long long row = idx / frames;     // stream*2 + ear
        float v = 0.f;                          // underflow: silence (:185-188)
        if (i < fifo_count) {
            int pos = fifo_read + i;
            if (pos >= fifo_cap) pos -= fifo_cap;
            v = fifo[row * fifo_cap + pos];
        }
        out.ptr[(row >> 1) * out.ss + (row & 1) * out.cs + out_offset + i] = v;
    }
}

cudaError_t launch_drain_fifo(const float *fifo, int fifo_cap, int fifo_read, int fifo_count, StridedOut out, int out_offset,
                              int frames, int n_streams, cudaStream_t st)
{
    const long long total = (long long)n_streams * 2 * frames;
    if (total <= 0) return cudaSuccess;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_drain_fifo<<<grid, 256, 0, st>>>(fifo, fifo_cap, fifo_read, fifo_count, out, out_offset, frames, n_streams);
    return cudaGetLastError();
}

__global__ void k_passthrough(StridedIn in, StridedOut out, int first_stream, int n_streams, int S, int frames)
{
    const long long total = (long long)n_streams * 2 * frames;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % frames);
        const long long row = idx / frames;
        const long long stream = first_stream + (row >> 1);
        const int ear = (int)(row & 1);
        const int ch = (ear == 1 && S > 1) ? 1 : 0;
        out.ptr[stream * out.ss + ear * out.cs + i] = in.ptr[stream * in.ss + ch * in.cs + i];
    }
}

cudaError_t launch_passthrough(StridedIn in, StridedOut out, int first_stream, int n_streams, int S, int frames, cudaStream_t st)
{
    const long long total = (long long)n_streams * 2 * frames;
    if (total <= 0) return cudaSuccess;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_passthrough<<<grid, 256, 0, st>>>(in, out, first_stream, n_streams, S, frames);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Synthetic input: counter-based hash keyed by (seed, stream, speaker, frame), SURVEY.md 8(d)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void k_synth_fill(float *out, int first_stream, int n_streams, int S, long long frame0, int frames, uint32_t seed)
{
    const long long total = (long long)n_streams * S * frames;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % frames);
        const long long row = idx / frames;
        const uint32_t s = (uint32_t)(row % S), stream = (uint32_t)(first_stream + row / S);
        uint32_t h = mix32(seed ^ mix32(stream * 0x9E3779B9U + 0x85EBCA6BU));
        h = mix32(h ^ (s * 0xC2B2AE35U + 0x27D4EB2FU));
        h = mix32(h ^ ((uint32_t)(frame0 + i) * 0x165667B1U + 0x9E3779B9U));
        out[idx] = ((float)(h >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.5f;
    }
}

cudaError_t launch_synth_fill(float *out, int first_stream, int n_streams, int S, long long frame0, int frames, uint32_t seed,
                              cudaStream_t st)
{
    const long long total = (long long)n_streams * S * frames;
    if (total <= 0) return cudaSuccess;
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    k_synth_fill<<<grid, 256, 0, st>>>(out, first_stream, n_streams, S, frame0, frames, seed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K5  eq: one thread per (stream, ear) channel; filter state in registers (FMAX bucket), float64.
// ------------------------------------------------------------------------------------------------
// abs(v) < 1e-30 ? 0 : v (:94-97), with the comparison done on the bit pattern (|v| as an unsigned integer orders like |v| for
// non-NaN doubles, and NaN compares "not below" either way) so that it stays off the float64 pipe, the scarce resource of K5.
__device__ __forceinline__ double flush_subnormal(double v)
{
    const unsigned long long mag = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffULL;
    return mag < 0x39B4484BFEEBC2A0ULL ? 0.0 : v;       // 0x39B4484BFEEBC2A0 = 1e-30
}

template <int FMAX>
__device__ __forceinline__ double biquad_cascade(double x, const EqProgram *__restrict__ prog, int nf, double (&z1)[FMAX], double (&z2)[FMAX])
{
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        if (f < nf) {
            const double b0 = prog->coef[f][0], b1 = prog->coef[f][1], b2 = prog->coef[f][2], a1 = prog->coef[f][3], a2 = prog->coef[f][4];
            // no FMA contraction: the reference (and the oracle) round every product and sum (:73-75)
            const double y = __dadd_rn(__dmul_rn(b0, x), z1[f]);
            const double n1 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x), __dmul_rn(a1, y)), z2[f]);
            const double n2 = __dsub_rn(__dmul_rn(b2, x), __dmul_rn(a2, y));
            z1[f] = flush_subnormal(n1);
            z2[f] = flush_subnormal(n2);
            x = y;
        }
    }
    return x;
}

template <int FMAX>
__global__ void __launch_bounds__(64) k_eq(const EqLaunch l, double *__restrict__ zstate, StridedOut io)
{
    __shared__ EqProgram progs[2];
    {
        const int words = (int)(sizeof(EqProgram) / sizeof(double));
        const double *src0 = reinterpret_cast<const double *>(l.from);
        double *dst0 = reinterpret_cast<double *>(&progs[0]);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst0[i] = src0[i];
        if (l.to != nullptr) {
            const double *src1 = reinterpret_cast<const double *>(l.to);
            double *dst1 = reinterpret_cast<double *>(&progs[1]);
            for (int i = threadIdx.x; i < words; i += blockDim.x) dst1[i] = src1[i];
        }
    }
    __syncthreads();
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;   // channel = stream*2 + ear within the launch
    if (ch >= l.n_streams * 2) return;
    const int stream = l.first_stream + (ch >> 1), ear = ch & 1;
    const bool fading = l.to != nullptr;
    const int nfA = progs[0].n_filters, nfB = fading ? progs[1].n_filters : 0;
    double zA1[FMAX], zA2[FMAX], zB1[FMAX], zB2[FMAX];
    double *zA = zstate + ((((size_t)stream * 2 + l.from_voice) * 2 + ear) * 64) * 2;
    double *zB = zstate + ((((size_t)stream * 2 + l.to_voice) * 2 + ear) * 64) * 2;
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        zA1[f] = f < nfA ? zA[2 * f] : 0.0; zA2[f] = f < nfA ? zA[2 * f + 1] : 0.0;
        zB1[f] = (fading && f < nfB) ? zB[2 * f] : 0.0; zB2[f] = (fading && f < nfB) ? zB[2 * f + 1] : 0.0;
    }
    float *p = io.ptr + stream * io.ss + ear * io.cs + l.seg_start;
    const double preA = progs[0].preamp_linear, preB = fading ? progs[1].preamp_linear : 1.0;
    for (int i = 0; i < l.seg_len; ++i) {
        const double x = (double)p[i];
        const float yo = (float)biquad_cascade<FMAX>(__dmul_rn(x, preA), &progs[0], nfA, zA1, zA2);   // :66, :88
        float r = yo;
        if (fading) {
            const float yn = (float)biquad_cascade<FMAX>(__dmul_rn(x, preB), &progs[1], nfB, zB1, zB2);
            const double progress = (double)(l.transition_frame + i + 1) / (double)l.transition_length;   // :298
            const double inverse = 1.0 - progress;
            r = (float)__dadd_rn(__dmul_rn((double)yo, inverse), __dmul_rn((double)yn, progress));        // :300-302
        }
        p[i] = r;
    }
#pragma unroll
    for (int f = 0; f < FMAX; ++f) {
        if (f < nfA) { zA[2 * f] = zA1[f]; zA[2 * f + 1] = zA2[f]; }
        if (fading && f < nfB) { zB[2 * f] = zB1[f]; zB[2 * f + 1] = zB2[f]; }
    }
}

// Steady state (no crossfade), 1..32 filters: the cascade as a systolic array across the lanes of a warp.  Lane f of a group of
// `gw` = n_filters lanes owns biquad f of one (stream, ear) channel — coefficients and both state words in registers — and at
// step t filters sample t - f, taking its input from lane f-1's output of the previous step (two 32-bit shuffles).  The
// per-sample critical path is one biquad instead of n_filters, every lane does useful float64 work, and a warp carries
// floor(32/gw) channels.  Same operations in the same order per biquad as ParametricEqualizerState.process (:65-90): bit-exact.
struct EqSteadyArgs {
    int n_segs, seg_start, seg_len, total_warps;
    EqSegment seg[kEqMaxSegments];
};

__global__ void __launch_bounds__(128) k_eq_systolic(const __grid_constant__ EqSteadyArgs a, double *__restrict__ zstate, StridedOut io)
{
    constexpr int kPerWarp = 1024;                          // doubles of staging per warp (>= 32 frames for each of up to 32 channels)
    __shared__ double stage_s[4][kPerWarp];                 // per warp: `groups` channels x `chunk` frames, staged coalesced
    // programmatic dependent launch (see k_persistent): the block kernel of the next call may be scheduled as this grid retires;
    // nothing is read before the convolution that feeds this equalizer has completed
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gwarp = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (gwarp >= a.total_warps) return;
    int si = 0;                                             // the stream range (equalizer) this warp serves
    while (si + 1 < a.n_segs && gwarp >= a.seg[si + 1].warp0) ++si;
    const EqSegment &l = a.seg[si];
    const int gw = l.n_filters;
    const int groups = 32 / gw;
    const int g = lane / gw, f = lane - g * gw;
    const long long warp = gwarp - l.warp0;
    const long long ch0 = warp * groups;                    // first channel (= stream*2 + ear within the range) of this warp
    const long long total = (long long)l.n_streams * 2;
    if (ch0 >= total) return;                               // whole warp idle
    const long long ch = ch0 + g;
    const bool live = g < groups && ch < total;
    const int stream = l.first_stream + (int)(live ? ch >> 1 : 0), ear = (int)(ch & 1);
    const EqProgram *prog = l.prog;
    const double pre = prog->preamp_linear;
    double b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0, z1 = 0, z2 = 0;
    double *zp = zstate + ((((size_t)stream * 2 + l.voice) * 2 + ear) * 64 + (live ? f : 0)) * 2;
    if (live) {
        b0 = prog->coef[f][0]; b1 = prog->coef[f][1]; b2 = prog->coef[f][2]; a1 = prog->coef[f][3]; a2 = prog->coef[f][4];
        z1 = zp[0]; z2 = zp[1];
    }
    const int chunk = (kPerWarp / groups) & ~31;            // frames per channel staged at a time (a multiple of 32)
    double *mine = stage_s[wid] + (g < groups ? g : 0) * chunk;
    const bool first = f == 0, last_f = f == gw - 1;
    for (int c0 = 0; c0 < a.seg_len; c0 += chunk) {
        const int cl = min(chunk, a.seg_len - c0);
        // coalesced load, one channel at a time; the float -> double conversion and the preamp (:66) happen here, in parallel,
        // instead of on the serial path below
        for (int q = 0; q < groups && ch0 + q < total; ++q) {
            const long long cq = ch0 + q;
            const float *src = io.ptr + (l.first_stream + (cq >> 1)) * io.ss + (cq & 1) * io.cs + a.seg_start + c0;
            for (int i = lane; i < cl; i += 32) stage_s[wid][q * chunk + i] = __dmul_rn((double)src[i], pre);
        }
        __syncwarp();
        double y = 0.0;                                     // my output of the previous step
        auto step = [&](int t, bool active) {
            const double up = __shfl_up_sync(0xffffffffu, y, 1);
            if (active) {
                const int i = t - f;                        // the sample this lane filters now
                const double x = first ? mine[i] : up;
                // no FMA contraction: the reference (and the oracle) round every product and sum (:73-75)
                y = __dadd_rn(__dmul_rn(b0, x), z1);
                const double n1 = __dadd_rn(__dsub_rn(__dmul_rn(b1, x), __dmul_rn(a1, y)), z2);
                const double n2 = __dsub_rn(__dmul_rn(b2, x), __dmul_rn(a2, y));
                z1 = flush_subnormal(n1);
                z2 = flush_subnormal(n2);
                if (last_f) mine[i] = y;                    // in place: sample i was consumed gw-1 steps ago
            }
        };
        const int ramp = min(gw - 1, cl);
        for (int t = 0; t < ramp; ++t) step(t, live && f <= t);                    // pipeline fills
        for (int t = ramp; t < cl; ++t) step(t, live);                             // every lane has a sample
        for (int t = cl; t < cl + gw - 1; ++t) step(t, live && t - f >= 0 && t - f < cl);   // pipeline drains
        __syncwarp();
        for (int q = 0; q < groups && ch0 + q < total; ++q) {                       // :88-89, coalesced
            const long long cq = ch0 + q;
            float *dst = io.ptr + (l.first_stream + (cq >> 1)) * io.ss + (cq & 1) * io.cs + a.seg_start + c0;
            for (int i = lane; i < cl; i += 32) dst[i] = (float)stage_s[wid][q * chunk + i];
        }
        __syncwarp();
    }
    if (live) { zp[0] = z1; zp[1] = z2; }
}

template <int FMAX>
static cudaError_t launch_eq_t(const EqLaunch &l, double *z, StridedOut io, cudaStream_t st)
{
    const int channels = l.n_streams * 2;
    const int grid = (channels + 63) / 64;
    k_eq<FMAX><<<grid, 64, 0, st>>>(l, z, io);
    return cudaGetLastError();
}

cudaError_t launch_eq_steady(const EqSegment *segs, int n_segs, int seg_start, int seg_len, double *z, StridedOut io, cudaStream_t st)
{
    if (n_segs <= 0 || seg_len <= 0) return cudaSuccess;
    if (n_segs > kEqMaxSegments) return cudaErrorInvalidValue;
    EqSteadyArgs a;
    memset(&a, 0, sizeof(a));
    int warps = 0;
    for (int i = 0; i < n_segs; ++i) {
        if (segs[i].n_filters < 1 || segs[i].n_filters > 32) return cudaErrorInvalidValue;
        a.seg[i] = segs[i];
        a.seg[i].warp0 = warps;
        const int groups = 32 / segs[i].n_filters;
        warps += (segs[i].n_streams * 2 + groups - 1) / groups;
    }
    if (warps == 0) return cudaSuccess;
    a.n_segs = n_segs; a.seg_start = seg_start; a.seg_len = seg_len; a.total_warps = warps;
    static const bool pdl = !(getenv("AW_PDL") && atoi(getenv("AW_PDL")) == 0);
    static bool carveout_set[64] = {};   // same carve-out as the block kernel, next to which this kernel's CTAs run (AW_ENGINE_OVERLAP_EQ)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !carveout_set[dev]) {
        cudaFuncSetAttribute(k_eq_systolic, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        carveout_set[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((warps + 3) / 4));
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_eq_systolic, a, z, io);
}

cudaError_t launch_eq(const EqLaunch &l, int max_filters, double *z, StridedOut io, cudaStream_t st)
{
    if (l.n_streams <= 0 || l.seg_len <= 0) return cudaSuccess;
    if (l.to == nullptr && max_filters >= 1 && max_filters <= 32) {   // max_filters = the active state's filter count here
        EqSegment seg{l.first_stream, l.n_streams, max_filters, l.from_voice, 0, 0, l.from};
        return launch_eq_steady(&seg, 1, l.seg_start, l.seg_len, z, io, st);
    }
    if (max_filters <= 4) return launch_eq_t<4>(l, z, io, st);
    if (max_filters <= 8) return launch_eq_t<8>(l, z, io, st);
    if (max_filters <= 16) return launch_eq_t<16>(l, z, io, st);
    if (max_filters <= 32) return launch_eq_t<32>(l, z, io, st);
    return launch_eq_t<64>(l, z, io, st);
}

__global__ void k_eq_reset(double *z, int first_stream, int n_streams, int voice_mask)
{
    const long long per_voice = 2 * 64 * 2;
    const long long total = (long long)n_streams * 2 * per_voice;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long sv = idx / per_voice;   // stream*2 + voice
        const int voice = (int)(sv & 1);
        if ((voice_mask >> voice) & 1) z[(long long)first_stream * 2 * per_voice + idx] = 0.0;
    }
}

cudaError_t launch_eq_reset(double *z, int first_stream, int n_streams, int voice_mask, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    const long long total = (long long)n_streams * 2 * 2 * 64 * 2;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    k_eq_reset<<<grid, 256, 0, st>>>(z, first_stream, n_streams, voice_mask);
    return cudaGetLastError();
}

}  // namespace aw
