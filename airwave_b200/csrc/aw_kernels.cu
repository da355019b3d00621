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
__device__ __forceinline__ double flush_subnormal(double v) { return fabs(v) < 1e-30 ? 0.0 : v; }   // :94-97

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

template <int FMAX>
static cudaError_t launch_eq_t(const EqLaunch &l, double *z, StridedOut io, cudaStream_t st)
{
    const int channels = l.n_streams * 2;
    const int grid = (channels + 63) / 64;
    k_eq<FMAX><<<grid, 64, 0, st>>>(l, z, io);
    return cudaGetLastError();
}

cudaError_t launch_eq(const EqLaunch &l, int max_filters, double *z, StridedOut io, cudaStream_t st)
{
    if (l.n_streams <= 0 || l.seg_len <= 0) return cudaSuccess;
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
