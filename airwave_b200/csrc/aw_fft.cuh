// aw_fft.cuh — shared-memory real FFT / inverse real FFT building blocks.
//
// Replaces vDSP_ctoz + vDSP_fft_zrip (forward, ConvolutionEngine.swift:169-174, 248-252) and
// vDSP_fft_zrip (inverse) + vDSP_vsmul + vDSP_ztoc (ConvolutionEngine.swift:353-363).
//
// A real FFT of N = 2B samples is one M = B point complex FFT over z[n] = x[2n] + i*x[2n+1]
// followed by a split step.  The complex FFT is a Stockham autosort (natural order in and out,
// ping-pong between two shared-memory buffers): one radix-2 stage when log2(M) is odd, then
// radix-4 stages.  Every stage is written as a per-butterfly function of a flat work index so
// that (a) a CTA runs `nf` transforms side by side with one butterfly per thread per stage and
// (b) the same code is exercised on the CPU by tests/cpu/fft_harness.cpp (AW_HD).
//
// Conventions (chosen so no scaling pass is ever needed on the hot path):
//   forward  : spec[k] = 2*X[k] (k = 0..M-1, spec[0] = (2*DC, 0)),  ny = 2*X[M]   — vDSP's x2
//   bank     : H'[k]   = (2*H[k]) * 0.25/N                                         — aw_bank_build
//   inverse  : unnormalised, so irfft(spec .* H') = x (*) h exactly as ConvolutionEngine.swift:356.
// Twiddles come from the plan table tw[k] = exp(-2*pi*i*k/N), k < N/2 (double precision on the
// host, rounded once to float) — the device twin of FFTSetupManager's cached FFTSetup.
#pragma once

#ifdef __CUDACC__
#define AW_HD __host__ __device__ __forceinline__
#else
#define AW_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace awfft {

AW_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AW_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
AW_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// w_M^t = exp(-2*pi*i*t/M) for t in [0, M), from the half-circle table of N = 2M; conjugated when INV.
template <bool INV>
AW_HD float2 twiddle_m(const float2 *tw, int t, int M)
{
    int i = 2 * t;
    float2 w;
    if (i >= M) { w = tw[i - M]; w.x = -w.x; w.y = -w.y; }
    else w = tw[i];
    if (INV) w.y = -w.y;
    return w;
}

// Radix-2 Stockham stage with Ns = 1 (twiddle-free).  i in [0, nf*M/2).
AW_HD void stage_r2_first(const float2 *x, float2 *y, int log2m, int i)
{
    const int M = 1 << log2m, half = M >> 1;
    const int f = i >> (log2m - 1), j = i & (half - 1);
    const float2 *xf = x + (size_t)f * M;
    float2 *yf = y + (size_t)f * M;
    const float2 v0 = xf[j], v1 = xf[j + half];
    yf[2 * j] = cadd(v0, v1);
    yf[2 * j + 1] = csub(v0, v1);
}

// Radix-4 Stockham stage.  i in [0, nf*M/4); Ns = product of the radices already applied.
template <bool INV>
AW_HD void stage_r4(const float2 *x, float2 *y, const float2 *tw, int log2m, int Ns, int i)
{
    const int M = 1 << log2m, Q = M >> 2;
    const int f = i >> (log2m - 2), j = i & (Q - 1);
    const int k = j & (Ns - 1);
    const float2 *xf = x + (size_t)f * M;
    float2 *yf = y + (size_t)f * M;
    float2 v0 = xf[j], v1 = xf[j + Q], v2 = xf[j + 2 * Q], v3 = xf[j + 3 * Q];
    if (Ns > 1) {
        const int t = k * (Q / Ns);                 // angle k*r/(4*Ns) of a turn = w_M^(r*t)
        v1 = cmul(v1, twiddle_m<INV>(tw, t, M));
        v2 = cmul(v2, twiddle_m<INV>(tw, 2 * t, M));
        v3 = cmul(v3, twiddle_m<INV>(tw, 3 * t, M));
    }
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = csub(v1, v3);
    // forward: -i*a3 ; inverse: +i*a3
    const float2 r3 = INV ? make_float2(-a3.y, a3.x) : make_float2(a3.y, -a3.x);
    const int j0 = ((j - k) << 2) + k;
    yf[j0] = cadd(a0, a2);
    yf[j0 + Ns] = cadd(a1, r3);
    yf[j0 + 2 * Ns] = csub(a0, a2);
    yf[j0 + 3 * Ns] = csub(a1, r3);
}

// Forward split step: Z (M-point FFT of the packed real frame) -> spec[0..M) and the Nyquist value.
// i in [0, nf*(M/2+1)): pair index k = 0..M/2 of transform f.
AW_HD void split_forward(const float2 *z, float2 *spec, float *ny, const float2 *tw, int log2m, int i)
{
    const int M = 1 << log2m, per = (M >> 1) + 1;
    const int f = i / per, k = i - f * per;
    const float2 *zf = z + (size_t)f * M;
    float2 *sf = spec + (size_t)f * M;
    if (k == 0) {
        const float2 z0 = zf[0];
        sf[0] = make_float2(2.0f * (z0.x + z0.y), 0.0f);
        ny[f] = 2.0f * (z0.x - z0.y);
        return;
    }
    const int j = M - k;
    const float2 a = zf[k], b = zf[j];
    const float er = a.x + b.x, ei = a.y - b.y;     // E = Z[k] + conj(Z[M-k])
    const float dr = a.x - b.x, di = a.y + b.y;     // D = Z[k] - conj(Z[M-k])
    const float2 w = tw[k];                         // exp(-2*pi*i*k/N)
    const float tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;   // T = w*D
    sf[k] = make_float2(er + ti, ei - tr);          // 2X[k]   = E - i*T
    if (j != k) sf[j] = make_float2(er - ti, -ei - tr);   // 2X[M-k] = conj(E) - i*conj(T)
}

// Inverse split step: spectrum acc[0..M) (+ Nyquist) -> Z whose inverse M-point FFT is the real signal.
AW_HD void split_inverse(const float2 *acc, const float *ny, float2 *z, const float2 *tw, int log2m, int i)
{
    const int M = 1 << log2m, per = (M >> 1) + 1;
    const int f = i / per, k = i - f * per;
    const float2 *af = acc + (size_t)f * M;
    float2 *zf = z + (size_t)f * M;
    if (k == 0) {
        const float dc = af[0].x, nq = ny[f];
        zf[0] = make_float2(dc + nq, dc - nq);
        return;
    }
    const int j = M - k;
    const float2 a = af[k], b = af[j];
    const float er = a.x + b.x, ei = a.y - b.y;     // E = X[k] + conj(X[M-k])
    const float dr = a.x - b.x, di = a.y + b.y;     // D = X[k] - conj(X[M-k])
    const float2 w = make_float2(tw[k].x, -tw[k].y);   // conj(w^k)
    const float tr = w.x * dr - w.y * di, ti = w.x * di + w.y * dr;   // T = conj(w)*D
    zf[k] = make_float2(er - ti, ei + tr);          // Z[k]   = E + i*T
    if (j != k) zf[j] = make_float2(er + ti, -ei + tr);   // Z[M-k] = conj(E) + i*conj(T)
}

#ifdef __CUDACC__
// CTA-wide batched complex FFT over nf transforms of M = 2^log2m points held in `a` (scratch `b`).
// Ends with a __syncthreads(); returns the buffer that holds the result.
template <bool INV>
__device__ __forceinline__ float2 *cfft_batched(float2 *a, float2 *b, const float2 *tw, int log2m, int nf)
{
    const int tid = threadIdx.x, nth = blockDim.x;
    float2 *x = a, *y = b;
    int Ns = 1;
    if (log2m & 1) {
        const int total = nf << (log2m - 1);
        for (int i = tid; i < total; i += nth) stage_r2_first(x, y, log2m, i);
        __syncthreads();
        float2 *t = x; x = y; y = t;
        Ns = 2;
    }
    const int M = 1 << log2m;
    const int total4 = nf << (log2m - 2);
    while (Ns < M) {
        for (int i = tid; i < total4; i += nth) stage_r4<INV>(x, y, tw, log2m, Ns, i);
        __syncthreads();
        float2 *t = x; x = y; y = t;
        Ns <<= 2;
    }
    return x;
}
#endif

}  // namespace awfft
