/*
 * airwave_oracle.c — CPU restatement of the reference binaural render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker or as the timed CPU baseline.  The product path (airwave_b200/)
 * never links, imports or falls back to this file.
 *
 * The reference (sallliisa/Airwave) is Swift + Apple Accelerate/vDSP + AVFoundation and can
 * neither be compiled nor executed on Linux (no Swift toolchain, closed-source frameworks).
 * This file restates its algorithm line by line in plain C; every function cites the
 * reference file:line it follows (paths relative to the reference checkout).
 *
 * FFT substitution (documented): vDSP_fft_zrip is replaced by an own radix-2 FFT that
 * produces the same packed layout and the same scale factors (real forward = 2 x DFT,
 * inverse = unnormalised), so the reference's `0.25 / fftSize` constant
 * (ConvolutionEngine.swift:356) is kept verbatim.  Results differ from vDSP by rounding only.
 *
 * Parity pinning: pinned against the reference's own known-answer tests
 * (ConvolutionEngineTests, RealtimeAudioProcessorTests, ParametricEqualizerProcessorTests —
 * see tests/test_oracle_*.py).  Resampler (vDSP_vgenp semantics) has no reference test:
 * "parity unpinned" for or_resample_vgenp.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

#define OR_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * vDSP_create_fftsetup / vDSP_fft_zrip substitute
 * (call sites: ConvolutionEngine.swift:82,174,252,353; FFTSetupManager.swift:51)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int log2n;      /* log2 of the REAL length N */
    int n;          /* N */
    int m;          /* N/2 complex points */
    float *cw_re;   /* twiddles of the M-point complex FFT: exp(-2*pi*i*k/M), k < M/2 */
    float *cw_im;
    float *rw_re;   /* split-step twiddles exp(-2*pi*i*k/N), k < M */
    float *rw_im;
    int *bitrev;    /* M entries */
} or_fft_setup;

OR_API or_fft_setup *or_fft_create(int log2n)
{
    if (log2n < 2 || log2n > 24) return NULL;
    or_fft_setup *s = (or_fft_setup *)calloc(1, sizeof(*s));
    s->log2n = log2n;
    s->n = 1 << log2n;
    s->m = s->n / 2;
    int m = s->m;
    s->cw_re = (float *)malloc(sizeof(float) * (m / 2 + 1));
    s->cw_im = (float *)malloc(sizeof(float) * (m / 2 + 1));
    s->rw_re = (float *)malloc(sizeof(float) * m);
    s->rw_im = (float *)malloc(sizeof(float) * m);
    s->bitrev = (int *)malloc(sizeof(int) * m);
    for (int k = 0; k < m / 2; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)m;
        s->cw_re[k] = (float)cos(a);
        s->cw_im[k] = (float)sin(a);
    }
    for (int k = 0; k < m; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)s->n;
        s->rw_re[k] = (float)cos(a);
        s->rw_im[k] = (float)sin(a);
    }
    int bits = log2n - 1;
    for (int i = 0; i < m; ++i) {
        int r = 0;
        for (int b = 0; b < bits; ++b)
            if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        s->bitrev[i] = r;
    }
    return s;
}

OR_API void or_fft_destroy(or_fft_setup *s)
{
    if (!s) return;
    free(s->cw_re); free(s->cw_im); free(s->rw_re); free(s->rw_im); free(s->bitrev);
    free(s);
}

/* In-place M-point complex FFT on split arrays; sign = -1 forward, +1 inverse (unnormalised). */
static void or_cfft(const or_fft_setup *s, float *re, float *im, int sign)
{
    const int m = s->m;
    for (int i = 0; i < m; ++i) {
        int j = s->bitrev[i];
        if (j > i) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int half = 1; half < m; half <<= 1) {
        const int step = m / (2 * half);
        for (int base = 0; base < m; base += 2 * half) {
            for (int k = 0; k < half; ++k) {
                const float wr = s->cw_re[k * step];
                const float wi = sign < 0 ? s->cw_im[k * step] : -s->cw_im[k * step];
                const int a = base + k, b = a + half;
                const float tr = re[b] * wr - im[b] * wi;
                const float ti = re[b] * wi + im[b] * wr;
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] = re[a] + tr; im[a] = im[a] + ti;
            }
        }
    }
}

/* vDSP_fft_zrip semantics on a split-complex buffer of N/2 elements.
 * forward (direction > 0): in  = even samples in re[], odd samples in im[] (after vDSP_ctoz);
 *                          out = packed spectrum, re[0] = 2*X[0], im[0] = 2*X[N/2],
 *                                re[k]+i*im[k] = 2*X[k], k = 1..N/2-1.
 * inverse (direction < 0): in  = packed spectrum S; out = unnormalised inverse DFT of S,
 *                                even samples in re[], odd samples in im[].
 * (ConvolutionEngine.swift:174,252 forward; :353 inverse) */
OR_API void or_fft_zrip(const or_fft_setup *s, float *re, float *im, int direction)
{
    const int m = s->m;
    if (direction > 0) {
        or_cfft(s, re, im, -1);
        const float z0r = re[0], z0i = im[0];
        re[0] = 2.0f * (z0r + z0i);
        im[0] = 2.0f * (z0r - z0i);
        for (int k = 1; k <= m / 2; ++k) {
            const int j = m - k;
            const float ar = re[k], ai = im[k], br = re[j], bi = im[j];
            /* E = Z[k] + conj(Z[j]); D = Z[k] - conj(Z[j]) */
            const float er = ar + br, ei = ai - bi;
            const float dr = ar - br, di = ai + bi;
            /* 2X[k] = E - i*T, T = w^k * D, w = exp(-2*pi*i/N);  2X[M-k] = conj(E) - i*conj(T) */
            const float wr = s->rw_re[k], wi = s->rw_im[k];
            const float tr = wr * dr - wi * di, ti = wr * di + wi * dr;
            re[k] = er + ti; im[k] = ei - tr;
            if (j != k) { re[j] = er - ti; im[j] = -ei - tr; }
        }
    } else {
        const float dc = re[0], ny = im[0];
        /* Z[0] = (X0 + XM) + i*(X0 - XM) */
        re[0] = dc + ny;
        im[0] = dc - ny;
        for (int k = 1; k <= m / 2; ++k) {
            const int j = m - k;
            const float ar = re[k], ai = im[k], br = re[j], bi = im[j];
            /* E = X[k] + conj(X[j]); D = X[k] - conj(X[j]); Z[k] = E + i*conj(w^k)*D */
            const float er = ar + br, ei = ai - bi;
            const float dr = ar - br, di = ai + bi;
            /* Z[k] = E + i*T, T = conj(w^k) * D;  Z[M-k] = conj(E) + i*conj(T) */
            const float wr = s->rw_re[k], wi = -s->rw_im[k];
            const float tr = wr * dr - wi * di, ti = wr * di + wi * dr;
            re[k] = er - ti; im[k] = ei + tr;
            if (j != k) { re[j] = er + ti; im[j] = -ei + tr; }
        }
        or_cfft(s, re, im, +1);
    }
}

/* ------------------------------------------------------------------------------------------
 * ConvolutionEngine  (ConvolutionEngine.swift:14-408)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int log2n, fftSize, fftSizeHalf, blockSize, partitionCount, partitionCountPow2;
    or_fft_setup *fftSetup;
    int ownsFFTSetup;
    float *inputBuffer;          /* fftSize   (:32,102) */
    float *inputOverlapBuffer;   /* blockSize (:33,105) */
    float *fdlReal, *fdlImag;    /* partitionCountPow2 * fftSizeHalf (:37-38,129-130) */
    int fdlIndex;                /* :39 */
    float *hrirReal, *hrirImag;  /* :42-43 */
    float *inRe, *inIm;          /* splitComplexInput :46-48 */
    float *accRe, *accIm;        /* accumulator :50-52 */
    float *tmpRe, *tmpIm;        /* tempMul :54-56 */
    float *tempOutputBuffer;     /* :59 */
} or_conv_engine;

static float *or_zalloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

/* ConvolutionEngine.init?(hrirSamples:blockSize:sharedFFTSetup:)  (:68-197) */
OR_API or_conv_engine *or_conv_create(const float *hrirSamples, int hrirCount, int blockSize,
                                      or_fft_setup *sharedFFTSetup)
{
    if (blockSize <= 0 || hrirCount < 0) return NULL;
    or_conv_engine *e = (or_conv_engine *)calloc(1, sizeof(*e));
    e->blockSize = blockSize;
    e->fftSize = blockSize * 2;                                   /* :72 */
    e->fftSizeHalf = e->fftSize / 2;                              /* :73 */
    e->log2n = (int)log2((double)e->fftSize);                     /* :74 (truncating) */
    if (sharedFFTSetup) { e->fftSetup = sharedFFTSetup; e->ownsFFTSetup = 0; }
    else {
        e->fftSetup = or_fft_create(e->log2n);                    /* :82 */
        if (!e->fftSetup || (1 << e->log2n) != e->fftSize) {      /* non power of two: invalid (Q10) */
            or_fft_destroy(e->fftSetup); free(e); return NULL;
        }
        e->ownsFFTSetup = 1;
    }
    e->partitionCount = (int)ceil((double)hrirCount / (double)blockSize);           /* :93 */
    if (e->partitionCount < 1) e->partitionCount = 1;
    e->partitionCountPow2 = 1 << (int)ceil(log2((double)e->partitionCount));        /* :96 */
    const int H = e->fftSizeHalf;
    e->inputBuffer = or_zalloc(e->fftSize);
    e->inputOverlapBuffer = or_zalloc(blockSize);
    e->inRe = or_zalloc(H); e->inIm = or_zalloc(H);
    e->accRe = or_zalloc(H); e->accIm = or_zalloc(H);
    e->tmpRe = or_zalloc(H); e->tmpIm = or_zalloc(H);
    e->tempOutputBuffer = or_zalloc(blockSize);
    const size_t total = (size_t)e->partitionCountPow2 * H;       /* :126-127 */
    e->fdlReal = or_zalloc(total); e->fdlImag = or_zalloc(total);
    e->hrirReal = or_zalloc(total); e->hrirImag = or_zalloc(total);
    float *tempPad = or_zalloc(e->fftSize);
    for (int p = 0; p < e->partitionCount; ++p) {                 /* :143-182 */
        memset(tempPad, 0, sizeof(float) * e->fftSize);
        const int startIdx = p * blockSize;
        int endIdx = startIdx + blockSize; if (endIdx > hrirCount) endIdx = hrirCount;
        for (int i = 0; i < endIdx - startIdx; ++i) tempPad[i] = hrirSamples[startIdx + i];
        float *hr = e->hrirReal + (size_t)p * H, *hi = e->hrirImag + (size_t)p * H;
        for (int i = 0; i < H; ++i) { hr[i] = tempPad[2 * i]; hi[i] = tempPad[2 * i + 1]; }  /* vDSP_ctoz :169 */
        or_fft_zrip(e->fftSetup, hr, hi, +1);                     /* :174 */
    }
    free(tempPad);
    return e;
}

OR_API void or_conv_destroy(or_conv_engine *e)
{
    if (!e) return;
    if (e->ownsFFTSetup) or_fft_destroy(e->fftSetup);
    free(e->inputBuffer); free(e->inputOverlapBuffer); free(e->inRe); free(e->inIm);
    free(e->accRe); free(e->accIm); free(e->tmpRe); free(e->tmpIm); free(e->tempOutputBuffer);
    free(e->fdlReal); free(e->fdlImag); free(e->hrirReal); free(e->hrirImag);
    free(e);
}

OR_API int or_conv_partition_count(const or_conv_engine *e) { return e->partitionCount; }
OR_API int or_conv_block_size(const or_conv_engine *e) { return e->blockSize; }

/* vDSP_zvmul(A,B,C,len,conj=1): C = A*B   (:311,346) */
static void or_zvmul(const float *ar, const float *ai, const float *br, const float *bi,
                     float *cr, float *ci, int len)
{
    for (int i = 0; i < len; ++i) {
        const float r = ar[i] * br[i] - ai[i] * bi[i];
        const float m = ar[i] * bi[i] + ai[i] * br[i];
        cr[i] = r; ci[i] = m;
    }
}

/* ConvolutionEngine.process(input:output:)  (:232-367) */
OR_API void or_conv_process(or_conv_engine *e, const float *input, float *output)
{
    const int B = e->blockSize, H = e->fftSizeHalf, P = e->partitionCount;
    memcpy(e->inputBuffer, e->inputOverlapBuffer, sizeof(float) * B);        /* :237 */
    memcpy(e->inputBuffer + B, input, sizeof(float) * B);                    /* :240 */
    memcpy(e->inputOverlapBuffer, input, sizeof(float) * B);                 /* :243 */
    for (int i = 0; i < H; ++i) { e->inRe[i] = e->inputBuffer[2 * i]; e->inIm[i] = e->inputBuffer[2 * i + 1]; } /* :248 */
    or_fft_zrip(e->fftSetup, e->inRe, e->inIm, +1);                          /* :252 */
    e->fdlIndex -= 1;                                                        /* :256-259 */
    if (e->fdlIndex < 0) e->fdlIndex += P;
    const size_t fdlOffset = (size_t)e->fdlIndex * H;
    memcpy(e->fdlReal + fdlOffset, e->inRe, sizeof(float) * H);              /* :263 */
    memcpy(e->fdlImag + fdlOffset, e->inIm, sizeof(float) * H);              /* :264 */
    memset(e->accRe, 0, sizeof(float) * H);                                  /* :270-271 */
    memset(e->accIm, 0, sizeof(float) * H);
    const int len = H - 1;                                                   /* :274 */
    {   /* p = 0  (:293-311) */
        const float *fr = e->fdlReal + fdlOffset, *fi = e->fdlImag + fdlOffset;
        e->accRe[0] = fr[0] * e->hrirReal[0];                                /* :304 */
        e->accIm[0] = fi[0] * e->hrirImag[0];                                /* :305 */
        or_zvmul(fr + 1, fi + 1, e->hrirReal + 1, e->hrirImag + 1, e->accRe + 1, e->accIm + 1, len);
    }
    for (int p = 1; p < P; ++p) {                                            /* :315-350 */
        int fdlIdx = e->fdlIndex + p;
        if (fdlIdx >= P) fdlIdx -= P;                                        /* :320-323 (modulus P, Q4) */
        const float *fr = e->fdlReal + (size_t)fdlIdx * H, *fi = e->fdlImag + (size_t)fdlIdx * H;
        const float *hr = e->hrirReal + (size_t)p * H, *hi = e->hrirImag + (size_t)p * H;
        e->accRe[0] += fr[0] * hr[0];                                        /* :336 */
        e->accIm[0] += fi[0] * hi[0];                                        /* :337 */
        or_zvmul(fr + 1, fi + 1, hr + 1, hi + 1, e->tmpRe + 1, e->tmpIm + 1, len);   /* :346 */
        for (int i = 1; i <= len; ++i) { e->accRe[i] += e->tmpRe[i]; e->accIm[i] += e->tmpIm[i]; } /* zvadd :347 */
    }
    or_fft_zrip(e->fftSetup, e->accRe, e->accIm, -1);                        /* :353 */
    const float scaleFactor = 0.25f / (float)e->fftSize;                     /* :356 */
    for (int i = 0; i < H; ++i) { e->accRe[i] *= scaleFactor; e->accIm[i] *= scaleFactor; }  /* :357-358 */
    for (int i = 0; i < H; ++i) { e->inputBuffer[2 * i] = e->accRe[i]; e->inputBuffer[2 * i + 1] = e->accIm[i]; } /* ztoc :362 */
    memcpy(output, e->inputBuffer + B, sizeof(float) * B);                   /* :366 */
}

/* process(input:[Float], output:, frameCount:) (:370-380): no-op unless frameCount == blockSize */
OR_API int or_conv_process_array(or_conv_engine *e, const float *input, float *output, int frameCount)
{
    if (frameCount != e->blockSize) return 0;
    or_conv_process(e, input, output);
    return 1;
}

/* processAndAccumulate  (:388-394) */
OR_API void or_conv_process_and_accumulate(or_conv_engine *e, const float *input, float *outputAccumulator)
{
    or_conv_process(e, input, e->tempOutputBuffer);
    for (int i = 0; i < e->blockSize; ++i) outputAccumulator[i] = outputAccumulator[i] + e->tempOutputBuffer[i];
}

/* reset()  (:397-407) */
OR_API void or_conv_reset(or_conv_engine *e)
{
    memset(e->inputBuffer, 0, sizeof(float) * e->fftSize);
    memset(e->inputOverlapBuffer, 0, sizeof(float) * e->blockSize);
    const size_t total = (size_t)e->partitionCountPow2 * e->fftSizeHalf;
    memset(e->fdlReal, 0, sizeof(float) * total);
    memset(e->fdlImag, 0, sizeof(float) * total);
    e->fdlIndex = 0;
}

/* ------------------------------------------------------------------------------------------
 * RealtimeAudioProcessor  (RealtimeAudioProcessor.swift:11-191) over VirtualSpeakerRenderer
 * (HRIRManager.swift:84-88).  `literal_stereo` = 1 reproduces the reference exactly
 * (min(renderers.count, 2), renderer 0 <- left, renderer 1 <- right, :145-147).
 * `literal_stereo` = 0 is the generalisation SURVEY.md Q1 defines for S > 2:
 * renderer i <- input channel i.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int blockSize, maxFramesPerCallback, fifoCapacity, rendererCount, inputCount, literalStereo;
    or_conv_engine **left, **right;     /* per renderer: convolverLeftEar / convolverRightEar */
    float **pending;                    /* inputCount x blockSize (pendingLeft / pendingRight generalised) */
    float *blockLeft, *blockRight;
    float **leftTemp, **rightTemp;
    float *fifoLeft, *fifoRight;
    int pendingCount, fifoReadIndex, fifoCount;
} or_rap;

static void or_rap_reset_storage(or_rap *r)                                  /* :129-139 */
{
    for (int c = 0; c < r->inputCount; ++c) memset(r->pending[c], 0, sizeof(float) * r->blockSize);
    memset(r->blockLeft, 0, sizeof(float) * r->blockSize);
    memset(r->blockRight, 0, sizeof(float) * r->blockSize);
    memset(r->fifoLeft, 0, sizeof(float) * r->fifoCapacity);
    memset(r->fifoRight, 0, sizeof(float) * r->fifoCapacity);
    r->pendingCount = 0; r->fifoReadIndex = 0; r->fifoCount = 0;
}

/* init(renderers:blockSize:maxFramesPerCallback:)  (:30-62).  The processor takes ownership
 * of nothing: engines are created/destroyed by the caller, as in the reference. */
OR_API or_rap *or_rap_create(or_conv_engine **left, or_conv_engine **right, int rendererCount,
                             int blockSize, int maxFramesPerCallback, int literalStereo)
{
    if (blockSize <= 0 || maxFramesPerCallback <= 0) return NULL;             /* preconditions :35-36 */
    or_rap *r = (or_rap *)calloc(1, sizeof(*r));
    r->blockSize = blockSize; r->maxFramesPerCallback = maxFramesPerCallback;
    r->fifoCapacity = maxFramesPerCallback + blockSize;                       /* :41 */
    r->rendererCount = rendererCount; r->literalStereo = literalStereo;
    r->inputCount = literalStereo ? 2 : (rendererCount > 0 ? rendererCount : 1);
    r->left = (or_conv_engine **)calloc(rendererCount ? rendererCount : 1, sizeof(void *));
    r->right = (or_conv_engine **)calloc(rendererCount ? rendererCount : 1, sizeof(void *));
    r->leftTemp = (float **)calloc(rendererCount ? rendererCount : 1, sizeof(void *));
    r->rightTemp = (float **)calloc(rendererCount ? rendererCount : 1, sizeof(void *));
    for (int i = 0; i < rendererCount; ++i) {
        r->left[i] = left[i]; r->right[i] = right[i];
        r->leftTemp[i] = or_zalloc(blockSize); r->rightTemp[i] = or_zalloc(blockSize);
    }
    r->pending = (float **)calloc(r->inputCount, sizeof(void *));
    for (int c = 0; c < r->inputCount; ++c) r->pending[c] = or_zalloc(blockSize);
    r->blockLeft = or_zalloc(blockSize); r->blockRight = or_zalloc(blockSize);
    r->fifoLeft = or_zalloc(r->fifoCapacity); r->fifoRight = or_zalloc(r->fifoCapacity);
    or_rap_reset_storage(r);
    return r;
}

OR_API void or_rap_destroy(or_rap *r)
{
    if (!r) return;
    for (int i = 0; i < r->rendererCount; ++i) { free(r->leftTemp[i]); free(r->rightTemp[i]); }
    for (int c = 0; c < r->inputCount; ++c) free(r->pending[c]);
    free(r->pending); free(r->left); free(r->right); free(r->leftTemp); free(r->rightTemp);
    free(r->blockLeft); free(r->blockRight); free(r->fifoLeft); free(r->fifoRight);
    free(r);
}

static void or_rap_process_pending_block(or_rap *r)                           /* :141-172 */
{
    const int B = r->blockSize;
    memset(r->blockLeft, 0, sizeof(float) * B);
    memset(r->blockRight, 0, sizeof(float) * B);
    int rendererCount = r->rendererCount;
    if (r->literalStereo && rendererCount > 2) rendererCount = 2;             /* :145 */
    for (int i = 0; i < rendererCount; ++i) {
        const float *input = r->pending[r->literalStereo ? (i == 0 ? 0 : 1) : i];   /* :147 */
        or_conv_process(r->left[i], input, r->leftTemp[i]);                   /* :149 */
        or_conv_process(r->right[i], input, r->rightTemp[i]);                 /* :150 */
        for (int k = 0; k < B; ++k) r->blockLeft[k] = r->blockLeft[k] + r->leftTemp[i][k];     /* vadd :152 */
        for (int k = 0; k < B; ++k) r->blockRight[k] = r->blockRight[k] + r->rightTemp[i][k];  /* vadd :158 */
    }
    for (int k = 0; k < B; ++k) {                                             /* :166-171 */
        const int writeIndex = (r->fifoReadIndex + r->fifoCount) % r->fifoCapacity;
        r->fifoLeft[writeIndex] = r->blockLeft[k];
        r->fifoRight[writeIndex] = r->blockRight[k];
        r->fifoCount += 1;
    }
}

/* process(inputLeft:inputRight:leftOutput:rightOutput:frameCount:)  (:77-119), generalised to
 * `inputCount` planar input pointers.  In literal-stereo mode inputs[1] may be NULL (mono
 * duplication, :95-107).  Returns 0 on success, -1 on a violated precondition (:85). */
OR_API int or_rap_process(or_rap *r, const float *const *inputs, float *leftOutput, float *rightOutput,
                          int frameCount)
{
    if (frameCount <= 0) return 0;                                            /* :84 */
    if (frameCount > r->maxFramesPerCallback) return -1;                      /* :85 */
    int inputOffset = 0;
    while (inputOffset < frameCount) {                                        /* :88-116 */
        int copyCount = r->blockSize - r->pendingCount;
        if (frameCount - inputOffset < copyCount) copyCount = frameCount - inputOffset;
        for (int c = 0; c < r->inputCount; ++c) {
            const float *src = inputs[c] ? inputs[c] : inputs[0];
            memcpy(r->pending[c] + r->pendingCount, src + inputOffset, sizeof(float) * copyCount);
        }
        r->pendingCount += copyCount;
        inputOffset += copyCount;
        if (r->pendingCount == r->blockSize) { or_rap_process_pending_block(r); r->pendingCount = 0; }
    }
    for (int i = 0; i < frameCount; ++i) {                                    /* drain :174-190 */
        if (r->fifoCount > 0) {
            const float l = r->fifoLeft[r->fifoReadIndex], rr = r->fifoRight[r->fifoReadIndex];
            leftOutput[i] = l;
            rightOutput[i] = rr;   /* written second: aliasing L==R keeps the right value, as in the reference */
            r->fifoReadIndex = (r->fifoReadIndex + 1) % r->fifoCapacity;
            r->fifoCount -= 1;
        } else { leftOutput[i] = 0; rightOutput[i] = 0; }
    }
    return 0;
}

OR_API void or_rap_reset(or_rap *r)                                           /* :121-127 */
{
    for (int i = 0; i < r->rendererCount; ++i) { or_conv_reset(r->left[i]); or_conv_reset(r->right[i]); }
    or_rap_reset_storage(r);
}

/* ------------------------------------------------------------------------------------------
 * Resampler.resampleHighQuality  (Resampler.swift:31-68) — vDSP_vramp + vDSP_vgenp, restated
 * from Apple's documented contract.  PARITY UNPINNED: no reference test.
 * Returns output count (0 when empty), or -1 if the rates are equal within 0.01 (caller keeps
 * the input, :33), or -2 for down-sampling, where the reference reads past its control vector.
 * ---------------------------------------------------------------------------------------- */
OR_API int or_resample_output_count(int count, double fromRate, double toRate)
{
    if (fabs(fromRate - toRate) < 0.01) return count;
    const double stride = fromRate / toRate;
    return (int)((double)count / stride);                                     /* :39 */
}

OR_API int or_resample_vgenp(const float *input, int count, double fromRate, double toRate, float *output)
{
    if (fabs(fromRate - toRate) < 0.01) return -1;                            /* :33 */
    const double stride = fromRate / toRate;                                  /* :38 */
    const int outputCount = (int)((double)count / stride);                    /* :39 */
    if (outputCount <= 0) return 0;                                           /* :41 */
    if (outputCount < count) return -2;          /* control[] shorter than M: out-of-bounds in the reference */
    const float start = 0.0f, step = (float)stride;                           /* :54-55 */
    /* control[m] = start + m*step  (vDSP_vramp :56); only m < count (= M) is consulted by vgenp */
    /* vDSP_vgenp(A=input, B=control, C=output, N=outputCount, M=count)  (:65) */
    const int M = count;
    int m = 0;                                   /* largest m with trunc(B[m]) < n, advanced monotonically */
    const float bLast = start + (float)(M - 1) * step;
    const float b0 = start;
    for (int n = 0; n < outputCount; ++n) {
        if ((float)n <= truncf(b0)) { output[n] = input[0]; continue; }
        if ((float)n > truncf(bLast)) { output[n] = input[M - 1]; continue; }
        while (m + 1 < M && truncf(start + (float)(m + 1) * step) < (float)n) ++m;
        const float bm = start + (float)m * step, bm1 = start + (float)(m + 1) * step;
        output[n] = input[m] + (input[m + 1] - input[m]) * ((float)n - bm) / (bm1 - bm);
    }
    return outputCount;
}

/* ------------------------------------------------------------------------------------------
 * BiquadCoefficientBuilder.make  (BiquadCoefficientBuilder.swift:30-107)
 * type: 0 peaking, 1 lowShelf, 2 highShelf.  Returns 0 or a BiquadCoefficientError code:
 * 1 invalidSampleRate, 2 invalidFrequency, 3 invalidQ, 4 nonFiniteInput, 5 nonFiniteCoefficients.
 * ---------------------------------------------------------------------------------------- */
OR_API int or_biquad_make(int type, double gainDB, double frequencyHz, double q, double sampleRate,
                          double *out5)
{
    if (!(isfinite(sampleRate) && sampleRate > 0)) return 1;                  /* :37 */
    if (!(isfinite(gainDB) && isfinite(frequencyHz) && isfinite(q))) return 4; /* :40 */
    if (!(frequencyHz > 0 && frequencyHz < sampleRate / 2)) return 2;         /* :43 */
    if (!(q > 0)) return 3;                                                   /* :46 */
    const double amplitude = pow(10, gainDB / 40);                            /* :50 */
    const double omega = 2 * M_PI * frequencyHz / sampleRate;
    const double sine = sin(omega), cosine = cos(omega);
    const double alpha = sine / (2 * q);
    const double beta = 2 * sqrt(amplitude) * alpha;                          /* :55 */
    double b0, b1, b2, a0, a1, a2;
    switch (type) {
    case 0:                                                                   /* :59-67 */
        b0 = 1 + alpha * amplitude; b1 = -2 * cosine; b2 = 1 - alpha * amplitude;
        a0 = 1 + alpha / amplitude; a1 = -2 * cosine; a2 = 1 - alpha / amplitude;
        break;
    case 1:                                                                   /* :68-76 */
        b0 = amplitude * ((amplitude + 1) - (amplitude - 1) * cosine + beta);
        b1 = 2 * amplitude * ((amplitude - 1) - (amplitude + 1) * cosine);
        b2 = amplitude * ((amplitude + 1) - (amplitude - 1) * cosine - beta);
        a0 = (amplitude + 1) + (amplitude - 1) * cosine + beta;
        a1 = -2 * ((amplitude - 1) + (amplitude + 1) * cosine);
        a2 = (amplitude + 1) + (amplitude - 1) * cosine - beta;
        break;
    case 2:                                                                   /* :77-85 */
        b0 = amplitude * ((amplitude + 1) + (amplitude - 1) * cosine + beta);
        b1 = -2 * amplitude * ((amplitude - 1) + (amplitude + 1) * cosine);
        b2 = amplitude * ((amplitude + 1) + (amplitude - 1) * cosine - beta);
        a0 = (amplitude + 1) - (amplitude - 1) * cosine + beta;
        a1 = 2 * ((amplitude - 1) - (amplitude + 1) * cosine);
        a2 = (amplitude + 1) - (amplitude - 1) * cosine - beta;
        break;
    default: return 4;
    }
    if (!(isfinite(a0) && a0 != 0)) return 5;                                 /* :88 */
    out5[0] = b0 / a0; out5[1] = b1 / a0; out5[2] = b2 / a0; out5[3] = a1 / a0; out5[4] = a2 / a0;
    for (int i = 0; i < 5; ++i) if (!isfinite(out5[i])) return 5;             /* :99-105 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * ParametricEqualizerState  (ParametricEqualizerProcessor.swift:16-98)
 * ---------------------------------------------------------------------------------------- */
#define OR_EQ_MAX_FILTERS 64                                                  /* :17 */
typedef struct {
    double sampleRate, preampLinear;
    int filterCount;
    double coef[OR_EQ_MAX_FILTERS][5];
    double lz1[OR_EQ_MAX_FILTERS], lz2[OR_EQ_MAX_FILTERS], rz1[OR_EQ_MAX_FILTERS], rz2[OR_EQ_MAX_FILTERS];
    int refcount;   /* the restatement's stand-in for Swift ARC on state objects */
} or_eq_state;

static double or_flush_subnormal(double v) { return fabs(v) < 1e-30 ? 0 : v; }   /* :94-97 */

/* ParametricEqualizerProcessor.prepare(definition:sampleRate:)  (:174-217).
 * filters: n x {type, enabled, frequencyHz, gainDB, q} as doubles; definition == nil <=> n = 0, preampDB = 0.
 * err: 0 ok, 1 invalidSampleRate, 2 nonFinitePreamp, 3 tooManyFilters, 4 invalidFilter
 * (*errIndex = index among ENABLED filters, *errCode = BiquadCoefficientError code). */
OR_API or_eq_state *or_eq_prepare(double preampDB, const double *filters, int n, double sampleRate,
                                  int *err, int *errIndex, int *errCode)
{
    *err = 0; *errIndex = -1; *errCode = 0;
    if (!(isfinite(sampleRate) && sampleRate > 0)) { *err = 1; return NULL; } /* :178 */
    if (!isfinite(preampDB)) { *err = 2; return NULL; }                       /* :183 */
    const double preampLinear = pow(10, preampDB / 20);                       /* :186 */
    if (!isfinite(preampLinear)) { *err = 2; return NULL; }
    int enabled = 0;
    for (int i = 0; i < n; ++i) if (filters[i * 5 + 1] != 0) ++enabled;       /* :191 */
    if (enabled > OR_EQ_MAX_FILTERS) { *err = 3; *errIndex = enabled; return NULL; }   /* :192 */
    or_eq_state *s = (or_eq_state *)calloc(1, sizeof(*s));
    s->sampleRate = sampleRate; s->preampLinear = preampLinear; s->refcount = 1;
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const double *f = filters + i * 5;
        if (f[1] == 0) continue;
        const int rc = or_biquad_make((int)f[0], f[3], f[2], f[4], sampleRate, s->coef[k]);   /* :200-206 */
        if (rc) { *err = 4; *errIndex = k; *errCode = rc; free(s); return NULL; }             /* :208 */
        ++k;
    }
    s->filterCount = k;
    return s;
}

OR_API void or_eq_state_release(or_eq_state *s) { if (s && --s->refcount == 0) free(s); }
static or_eq_state *or_eq_retain(or_eq_state *s) { if (s) ++s->refcount; return s; }
OR_API int or_eq_state_filter_count(const or_eq_state *s) { return s->filterCount; }
OR_API void or_eq_state_coefficients(const or_eq_state *s, int i, double *out5) { memcpy(out5, s->coef[i], sizeof(double) * 5); }
OR_API double or_eq_state_preamp_linear(const or_eq_state *s) { return s->preampLinear; }

OR_API void or_eq_state_reset(or_eq_state *s)                                 /* :47-54 */
{
    for (int i = 0; i < s->filterCount; ++i) s->lz1[i] = s->lz2[i] = s->rz1[i] = s->rz2[i] = 0;
}

/* ParametricEqualizerState.process  (:58-91); inputRight may be NULL (:67); in-place allowed. */
OR_API void or_eq_state_process(or_eq_state *s, const float *inputLeft, const float *inputRight,
                                float *leftOutput, float *rightOutput, int frameCount)
{
    for (int frame = 0; frame < frameCount; ++frame) {
        double left = (double)inputLeft[frame] * s->preampLinear;             /* :66 */
        double right = (double)(inputRight ? inputRight[frame] : inputLeft[frame]) * s->preampLinear;
        for (int f = 0; f < s->filterCount; ++f) {
            const double *c = s->coef[f];   /* b0 b1 b2 a1 a2 */
            const double lo = c[0] * left + s->lz1[f];                        /* :73 */
            const double lz1 = c[1] * left - c[3] * lo + s->lz2[f];           /* :74 */
            const double lz2 = c[2] * left - c[4] * lo;                       /* :75 */
            s->lz1[f] = or_flush_subnormal(lz1); s->lz2[f] = or_flush_subnormal(lz2);
            left = lo;
            const double ro = c[0] * right + s->rz1[f];                       /* :80 */
            const double rz1 = c[1] * right - c[3] * ro + s->rz2[f];
            const double rz2 = c[2] * right - c[4] * ro;
            s->rz1[f] = or_flush_subnormal(rz1); s->rz2[f] = or_flush_subnormal(rz2);
            right = ro;
        }
        leftOutput[frame] = (float)left;                                      /* :88-89 */
        rightOutput[frame] = (float)right;
    }
}

/* ------------------------------------------------------------------------------------------
 * ParametricEqualizerProcessor  (ParametricEqualizerProcessor.swift:121-408), single-threaded
 * model: the try-locks always succeed unless `lockHeld` simulates a contended publication lock
 * (withPublicationLockForTesting, :229-234).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double sampleRate;
    int maxFramesPerCallback, transitionLength, transitionFrame;
    or_eq_state *unityState, *publishedTarget, *retired;
    int resetRequested, lockHeld;
    or_eq_state *audioThreadTarget, *activeState, *transitionFrom, *transitionTo, *pendingTarget,
        *observedTarget, *pendingRetirement;
    float *oldScratch, *oldRightScratch, *newScratch, *newRightScratch;
} or_eq_processor;

OR_API or_eq_processor *or_eqp_create(double sampleRate, int maxFramesPerCallback, int *err)
{
    *err = 0;
    if (!(isfinite(sampleRate) && sampleRate > 0)) { *err = 1; return NULL; }              /* :149 */
    if (!(maxFramesPerCallback > 0 && maxFramesPerCallback <= 4096)) { *err = 3; return NULL; } /* :152 */
    or_eq_processor *p = (or_eq_processor *)calloc(1, sizeof(*p));
    int e, ei, ec;
    p->sampleRate = sampleRate; p->maxFramesPerCallback = maxFramesPerCallback;
    p->unityState = or_eq_prepare(0, NULL, 0, sampleRate, &e, &ei, &ec);                   /* :158 */
    p->activeState = or_eq_retain(p->unityState);                                          /* :159 */
    long tl = lround(sampleRate * 0.020);                                                  /* :160 */
    p->transitionLength = tl < 1 ? 1 : (int)tl;
    p->oldScratch = or_zalloc(maxFramesPerCallback); p->oldRightScratch = or_zalloc(maxFramesPerCallback);
    p->newScratch = or_zalloc(maxFramesPerCallback); p->newRightScratch = or_zalloc(maxFramesPerCallback);
    return p;
}

OR_API void or_eqp_destroy(or_eq_processor *p)
{
    if (!p) return;
    or_eq_state *all[] = {p->unityState, p->publishedTarget, p->retired, p->audioThreadTarget, p->activeState,
                          p->transitionFrom, p->transitionTo, p->pendingTarget, p->pendingRetirement};
    for (unsigned i = 0; i < sizeof(all) / sizeof(all[0]); ++i) or_eq_state_release(all[i]);
    /* observedTarget is compared by identity only and never retained */
    free(p->oldScratch); free(p->oldRightScratch); free(p->newScratch); free(p->newRightScratch);
    free(p);
}

static void or_assign(or_eq_state **slot, or_eq_state *v)
{
    or_eq_state *old = *slot;
    *slot = or_eq_retain(v);
    or_eq_state_release(old);
}

/* setTarget(definition:) = publish(prepare(...))  (:219-238).  Returns or_eq_prepare's err. */
OR_API int or_eqp_set_target(or_eq_processor *p, double preampDB, const double *filters, int n,
                             int *errIndex, int *errCode)
{
    int err;
    or_eq_state *s = or_eq_prepare(preampDB, filters, n, p->sampleRate, &err, errIndex, errCode);
    if (!s) return err;
    or_assign(&p->publishedTarget, s);                                                     /* :223-225 */
    or_eq_state_release(s);
    return 0;
}
OR_API void or_eqp_reset(or_eq_processor *p) { p->resetRequested = 1; }                    /* :240-244 */
OR_API void or_eqp_drain_retired_states(or_eq_processor *p) { or_assign(&p->retired, NULL); } /* :247-251 */
OR_API void or_eqp_hold_publication_lock(or_eq_processor *p, int held) { p->lockHeld = held; }

static void or_eqp_begin_transition(or_eq_processor *p, or_eq_state *target)               /* :354-359 */
{
    if (target == p->activeState) return;
    or_assign(&p->transitionFrom, p->activeState);
    or_assign(&p->transitionTo, target);
    p->transitionFrame = 0;
}

static int or_eqp_retire(or_eq_processor *p, or_eq_state *state)                           /* :377-389 */
{
    if (p->pendingRetirement != NULL) return 0;
    if (p->retired == NULL) { or_assign(&p->retired, state); return 1; }
    or_assign(&p->pendingRetirement, state);
    return 0;
}

static void or_eqp_start_pending(or_eq_processor *p)
{
    if (p->pendingTarget) {
        or_eq_state *pending = or_eq_retain(p->pendingTarget);
        or_assign(&p->pendingTarget, NULL);
        if (pending != p->activeState) or_eqp_begin_transition(p, pending);
        or_eq_state_release(pending);
    }
}

static void or_eqp_finish_transition(or_eq_processor *p)                                   /* :361-375 */
{
    if (!p->transitionFrom || !p->transitionTo) return;
    or_eq_state *from = or_eq_retain(p->transitionFrom);
    or_assign(&p->activeState, p->transitionTo);
    or_assign(&p->transitionFrom, NULL);
    or_assign(&p->transitionTo, NULL);
    p->transitionFrame = 0;
    const int ok = or_eqp_retire(p, from);
    or_eq_state_release(from);
    if (!ok) return;
    or_eqp_start_pending(p);
}

static void or_eqp_observe_published_target(or_eq_processor *p)                            /* :317-339 */
{
    if (!p->lockHeld && p->publishedTarget) or_assign(&p->audioThreadTarget, p->publishedTarget);
    or_eq_state *target = p->audioThreadTarget;
    if (!target || target == p->observedTarget) return;
    p->observedTarget = target;
    if (p->transitionTo != NULL) {
        if (target != p->transitionTo) or_assign(&p->pendingTarget, target);
    } else if (p->pendingRetirement != NULL) {
        or_assign(&p->pendingTarget, target);
    } else if (target != p->activeState) {
        or_eqp_begin_transition(p, target);
    }
}

static void or_eqp_flush_pending_retirement(or_eq_processor *p)                            /* :391-407 */
{
    if (!p->pendingRetirement) return;
    if (p->retired != NULL) return;
    or_assign(&p->retired, p->pendingRetirement);
    or_assign(&p->pendingRetirement, NULL);
    or_eqp_start_pending(p);
}

static void or_eqp_apply_pending_reset(or_eq_processor *p)                                 /* :341-352 */
{
    if (!p->resetRequested) return;
    p->resetRequested = 0;
    or_eq_state_reset(p->activeState);
    if (p->transitionFrom) or_eq_state_reset(p->transitionFrom);
    if (p->transitionTo) or_eq_state_reset(p->transitionTo);
}

/* process  (:254-314).  Returns -1 on frameCount > maxFramesPerCallback (precondition :262). */
OR_API int or_eqp_process(or_eq_processor *p, const float *inputLeft, const float *inputRight,
                          float *leftOutput, float *rightOutput, int frameCount)
{
    if (frameCount <= 0) return 0;
    if (frameCount > p->maxFramesPerCallback) return -1;
    or_eqp_observe_published_target(p);
    or_eqp_flush_pending_retirement(p);
    or_eqp_apply_pending_reset(p);
    int offset = 0;
    while (offset < frameCount) {
        if (!p->transitionFrom || !p->transitionTo) {                                      /* :269-278 */
            or_eq_state_process(p->activeState, inputLeft + offset, inputRight ? inputRight + offset : NULL,
                                leftOutput + offset, rightOutput + offset, frameCount - offset);
            return 0;
        }
        const int remaining = p->transitionLength - p->transitionFrame;
        const int segment = remaining < frameCount - offset ? remaining : frameCount - offset;
        or_eq_state_process(p->transitionFrom, inputLeft + offset, inputRight ? inputRight + offset : NULL,
                            p->oldScratch, p->oldRightScratch, segment);                  /* :282-288 */
        or_eq_state_process(p->transitionTo, inputLeft + offset, inputRight ? inputRight + offset : NULL,
                            p->newScratch, p->newRightScratch, segment);                  /* :289-295 */
        for (int i = 0; i < segment; ++i) {                                                /* :297-306 */
            const double progress = (double)(p->transitionFrame + i + 1) / (double)p->transitionLength;
            const double inverse = 1 - progress;
            leftOutput[offset + i] = (float)((double)p->oldScratch[i] * inverse + (double)p->newScratch[i] * progress);
            rightOutput[offset + i] = (float)((double)p->oldRightScratch[i] * inverse + (double)p->newRightScratch[i] * progress);
        }
        p->transitionFrame += segment;
        offset += segment;
        if (p->transitionFrame == p->transitionLength) or_eqp_finish_transition(p);        /* :310-312 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * float64 direct convolution — the numerical oracle the north_star names (max-abs 1e-5,
 * SNR >= 100 dB).  out[e][n] = sum_s sum_k x[s][n-k] * h[s][e][k], n in [0, frames).
 * x: [S][frames]; h: [S][2][taps] (already mapped speaker -> ear pair); out: [2][frames] double.
 * ---------------------------------------------------------------------------------------- */
typedef struct { const float *x; int S, frames; const float *h; int taps; double *out; int n0, n1; } or_dc_job;

static void *or_dc_worker(void *arg)
{
    const or_dc_job *j = (const or_dc_job *)arg;
    for (int n = j->n0; n < j->n1; ++n) {
        double accL = 0, accR = 0;
        for (int s = 0; s < j->S; ++s) {
            const float *xs = j->x + (size_t)s * j->frames;
            const float *hl = j->h + ((size_t)s * 2 + 0) * j->taps, *hr = j->h + ((size_t)s * 2 + 1) * j->taps;
            const int kmax = n < j->taps - 1 ? n : j->taps - 1;
            double l = 0, r = 0;
            for (int k = 0; k <= kmax; ++k) {
                const double xv = (double)xs[n - k];
                l += xv * (double)hl[k];
                r += xv * (double)hr[k];
            }
            accL += l; accR += r;
        }
        j->out[n] = accL;
        j->out[(size_t)j->frames + n] = accR;
    }
    return NULL;
}

OR_API int or_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

OR_API void or_direct_conv_f64(const float *x, int S, int frames, const float *h, int taps, double *out)
{
    int nt = or_max_threads(); if (nt > 64) nt = 64; if (nt > frames) nt = frames > 0 ? frames : 1;
    pthread_t th[64]; or_dc_job jobs[64];
    for (int t = 0; t < nt; ++t) {
        jobs[t] = (or_dc_job){x, S, frames, h, taps, out, (int)((long long)frames * t / nt), (int)((long long)frames * (t + 1) / nt)};
        pthread_create(&th[t], NULL, or_dc_worker, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------
 * Counter-based synthetic input: uniform in [-0.25, 0.25], keyed by
 * (seed, stream, speaker, frame) so the CPU and the GPU regenerate identical data.
 * The CUDA side carries its own copy of this hash (airwave_b200/csrc/aw_synth.cuh).
 * ---------------------------------------------------------------------------------------- */
static inline uint32_t or_mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
OR_API float or_synth_sample(uint32_t seed, uint32_t stream, uint32_t speaker, uint32_t frame)
{
    uint32_t h = or_mix32(seed ^ or_mix32(stream * 0x9E3779B9U + 0x85EBCA6BU));
    h = or_mix32(h ^ (speaker * 0xC2B2AE35U + 0x27D4EB2FU));
    h = or_mix32(h ^ (frame * 0x165667B1U + 0x9E3779B9U));
    return ((float)(h >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.5f;
}
OR_API void or_synth_fill(uint32_t seed, uint32_t stream, uint32_t speaker, uint32_t frame0, int frames, float *out)
{
    for (int i = 0; i < frames; ++i) out[i] = or_synth_sample(seed, stream, speaker, frame0 + (uint32_t)i);
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline driver: the reference algorithm exactly as structured in the reference
 * (one ConvolutionEngine per speaker x ear, per-ear forward FFT, zvmul + zvadd as separate
 * passes, vDSP_vadd mix; RealtimeAudioProcessor.swift:141-172 generalised to S renderers),
 * one stream per thread.  h: [S][2][taps].  Renders `blocks` blocks of `B` frames for
 * `n_streams` streams of synthetic input and returns wall seconds (engine construction
 * excluded).  checksum (optional) receives the sum of all output samples.
 * ---------------------------------------------------------------------------------------- */
static double or_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    or_rap **raps; const float *all_in; size_t per_stream; int S, B, blocks, n_streams;
    volatile int *next; double sum;
} or_bench_job;

static void *or_bench_worker(void *arg)
{
    or_bench_job *j = (or_bench_job *)arg;
    const int S = j->S, B = j->B;
    float *oL = (float *)malloc(sizeof(float) * B), *oR = (float *)malloc(sizeof(float) * B);
    const float **ptrs = (const float **)malloc(sizeof(void *) * S);
    double sum = 0;
    for (;;) {
        const int t = __sync_fetch_and_add(j->next, 1);        /* one stream per thread at a time */
        if (t >= j->n_streams) break;
        for (int b = 0; b < j->blocks; ++b) {
            const float *in = j->all_in + (size_t)t * j->per_stream + (size_t)b * S * B;
            for (int s = 0; s < S; ++s) ptrs[s] = in + (size_t)s * B;
            or_rap_process(j->raps[t], ptrs, oL, oR, B);
            for (int i = 0; i < B; ++i) sum += (double)oL[i] + (double)oR[i];
        }
    }
    free(oL); free(oR); free((void *)ptrs);
    j->sum = sum;
    return NULL;
}

OR_API double or_bench_render(int n_streams, int S, int B, const float *h, int taps, int blocks,
                              int threads, uint32_t seed, double *checksum)
{
    or_conv_engine **L = (or_conv_engine **)calloc((size_t)n_streams * S, sizeof(void *));
    or_conv_engine **R = (or_conv_engine **)calloc((size_t)n_streams * S, sizeof(void *));
    or_rap **raps = (or_rap **)calloc(n_streams, sizeof(void *));
    for (int t = 0; t < n_streams; ++t) {
        for (int s = 0; s < S; ++s) {
            L[(size_t)t * S + s] = or_conv_create(h + ((size_t)s * 2 + 0) * taps, taps, B, NULL);
            R[(size_t)t * S + s] = or_conv_create(h + ((size_t)s * 2 + 1) * taps, taps, B, NULL);
        }
        raps[t] = or_rap_create(L + (size_t)t * S, R + (size_t)t * S, S, B, B, 0);
    }
    /* synthetic input is generated before the timed region: [stream][block][S][B] */
    const size_t per_stream = (size_t)blocks * S * B;
    float *all_in = (float *)malloc(sizeof(float) * per_stream * (size_t)n_streams);
    for (int t = 0; t < n_streams; ++t)
        for (int b = 0; b < blocks; ++b)
            for (int s = 0; s < S; ++s)
                or_synth_fill(seed, (uint32_t)t, (uint32_t)s, (uint32_t)b * B, B,
                              all_in + (size_t)t * per_stream + ((size_t)b * S + s) * B);
    if (threads < 1) threads = or_max_threads();
    if (threads > 256) threads = 256;
    pthread_t th[256]; or_bench_job jobs[256];
    volatile int next = 0;
    const double t0 = or_now();
    for (int i = 0; i < threads; ++i) {
        jobs[i] = (or_bench_job){raps, all_in, per_stream, S, B, blocks, n_streams, &next, 0.0};
        pthread_create(&th[i], NULL, or_bench_worker, &jobs[i]);
    }
    double sum = 0;
    for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); sum += jobs[i].sum; }
    const double t1 = or_now();
    free(all_in);
    for (int t = 0; t < n_streams; ++t) {
        or_rap_destroy(raps[t]);
        for (int s = 0; s < S; ++s) { or_conv_destroy(L[(size_t)t * S + s]); or_conv_destroy(R[(size_t)t * S + s]); }
    }
    free(L); free(R); free(raps);
    if (checksum) *checksum = sum;
    return t1 - t0;
}

/* ------------------------------------------------------------------------------------------
 * Persistent CPU batch for bench.py --impl reference: engines are built once, every step
 * renders `blocks` blocks for each of the sample's streams (one stream per thread at a time,
 * all host threads).  Same structure as or_bench_render.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int n_streams, S, B, ring_blocks, next_block;
    or_conv_engine **L, **R;
    or_rap **raps;
    or_eq_state **eq;   /* optional: one ParametricEqualizerState per stream, applied after the spatial stage (AudioEffectGraph.swift:195-210) */
    float *input;   /* [stream][ring_blocks][S][B], synthesised once */
} or_batch;

OR_API or_batch *or_batch_create(int n_streams, int S, int B, const float *h, int taps, int ring_blocks, uint32_t seed)
{
    or_batch *b = (or_batch *)calloc(1, sizeof(*b));
    b->n_streams = n_streams; b->S = S; b->B = B; b->ring_blocks = ring_blocks;
    b->L = (or_conv_engine **)calloc((size_t)n_streams * S, sizeof(void *));
    b->R = (or_conv_engine **)calloc((size_t)n_streams * S, sizeof(void *));
    b->raps = (or_rap **)calloc(n_streams, sizeof(void *));
    for (int t = 0; t < n_streams; ++t) {
        for (int s = 0; s < S; ++s) {
            b->L[(size_t)t * S + s] = or_conv_create(h + ((size_t)s * 2 + 0) * taps, taps, B, NULL);
            b->R[(size_t)t * S + s] = or_conv_create(h + ((size_t)s * 2 + 1) * taps, taps, B, NULL);
        }
        b->raps[t] = or_rap_create(b->L + (size_t)t * S, b->R + (size_t)t * S, S, B, B, 0);
    }
    const size_t per_stream = (size_t)ring_blocks * S * B;
    b->input = (float *)malloc(sizeof(float) * per_stream * (size_t)n_streams);
    for (int t = 0; t < n_streams; ++t)
        for (int k = 0; k < ring_blocks; ++k)
            for (int s = 0; s < S; ++s)
                or_synth_fill(seed, (uint32_t)t, (uint32_t)s, (uint32_t)k * B, B, b->input + (size_t)t * per_stream + ((size_t)k * S + s) * B);
    return b;
}

/* Full chain for the CPU baseline of BASELINE.json configs[3]: every stream gets its own equalizer state. */
OR_API int or_batch_set_eq(or_batch *b, double preampDB, const double *filters, int n, double sampleRate)
{
    if (!b->eq) b->eq = (or_eq_state **)calloc(b->n_streams, sizeof(void *));
    for (int t = 0; t < b->n_streams; ++t) {
        int err, ei, ec;
        if (b->eq[t]) or_eq_state_release(b->eq[t]);
        b->eq[t] = or_eq_prepare(preampDB, filters, n, sampleRate, &err, &ei, &ec);
        if (!b->eq[t]) return err;
    }
    return 0;
}

typedef struct { or_batch *b; int blocks; volatile int *next; double sum; } or_batch_job;

static void *or_batch_worker(void *arg)
{
    or_batch_job *j = (or_batch_job *)arg;
    or_batch *b = j->b;
    const int S = b->S, B = b->B;
    float *oL = (float *)malloc(sizeof(float) * B), *oR = (float *)malloc(sizeof(float) * B);
    const float **ptrs = (const float **)malloc(sizeof(void *) * S);
    const size_t per_stream = (size_t)b->ring_blocks * S * B;
    double sum = 0;
    for (;;) {
        const int t = __sync_fetch_and_add(j->next, 1);
        if (t >= b->n_streams) break;
        for (int k = 0; k < j->blocks; ++k) {
            const int rb = (b->next_block + k) % b->ring_blocks;
            const float *in = b->input + (size_t)t * per_stream + (size_t)rb * S * B;
            for (int s = 0; s < S; ++s) ptrs[s] = in + (size_t)s * B;
            or_rap_process(b->raps[t], ptrs, oL, oR, B);
            if (b->eq) or_eq_state_process(b->eq[t], oL, oR, oL, oR, B);
            sum += (double)oL[B - 1] + (double)oR[0];
        }
    }
    free(oL); free(oR); free((void *)ptrs);
    j->sum = sum;
    return NULL;
}

/* Renders `blocks` blocks for every stream with `threads` host threads; returns wall seconds. */
OR_API double or_batch_step(or_batch *b, int blocks, int threads, double *checksum)
{
    if (threads < 1) threads = or_max_threads();
    if (threads > 256) threads = 256;
    pthread_t th[256]; or_batch_job jobs[256];
    volatile int next = 0;
    const double t0 = or_now();
    for (int i = 0; i < threads; ++i) {
        jobs[i] = (or_batch_job){b, blocks, &next, 0.0};
        pthread_create(&th[i], NULL, or_batch_worker, &jobs[i]);
    }
    double sum = 0;
    for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); sum += jobs[i].sum; }
    const double t1 = or_now();
    b->next_block = (b->next_block + blocks) % b->ring_blocks;
    if (checksum) *checksum = sum;
    return t1 - t0;
}

OR_API void or_batch_destroy(or_batch *b)
{
    if (!b) return;
    for (int t = 0; t < b->n_streams; ++t) {
        or_rap_destroy(b->raps[t]);
        for (int s = 0; s < b->S; ++s) { or_conv_destroy(b->L[(size_t)t * b->S + s]); or_conv_destroy(b->R[(size_t)t * b->S + s]); }
    }
    if (b->eq) { for (int t = 0; t < b->n_streams; ++t) if (b->eq[t]) or_eq_state_release(b->eq[t]); free(b->eq); }
    free(b->L); free(b->R); free(b->raps); free(b->input); free(b);
}
