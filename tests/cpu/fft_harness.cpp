// CPU harness for airwave_b200/csrc/aw_fft.cuh: runs the per-butterfly stage functions the CUDA
// kernels use, with plain loops in place of threads, so the indexing is checked without a GPU.
#include <cmath>
#include <cstddef>
#include <vector>
#include "../../airwave_b200/csrc/aw_fft.cuh"

using namespace awfft;

static std::vector<float2> make_tw(int log2m)
{
    const int M = 1 << log2m, N = 2 * M;
    std::vector<float2> tw(M);
    for (int k = 0; k < M; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    return tw;
}

template <bool INV>
static float2 *cfft(float2 *a, float2 *b, const float2 *tw, int log2m, int nf)
{
    float2 *x = a, *y = b;
    int Ns = 1;
    if (log2m & 1) {
        for (int i = 0; i < (nf << (log2m - 1)); ++i) stage_r2_first(x, y, log2m, i);
        std::swap(x, y);
        Ns = 2;
    }
    const int M = 1 << log2m;
    while (Ns < M) {
        for (int i = 0; i < (nf << (log2m - 2)); ++i) stage_r4<INV>(x, y, tw, log2m, Ns, i);
        std::swap(x, y);
        Ns <<= 2;
    }
    return x;
}

extern "C" {
// x: nf frames of N = 2M reals -> spec: nf x M complex (interleaved), ny: nf
void harness_rfft_forward(const float *x, int log2m, int nf, float *spec, float *ny)
{
    const int M = 1 << log2m;
    auto tw = make_tw(log2m);
    std::vector<float2> a((size_t)nf * M), b((size_t)nf * M);
    for (size_t i = 0; i < (size_t)nf * M; ++i) a[i] = make_float2(x[2 * i], x[2 * i + 1]);
    float2 *z = cfft<false>(a.data(), b.data(), tw.data(), log2m, nf);
    float2 *out = (z == a.data()) ? b.data() : a.data();
    for (int i = 0; i < nf * (M / 2 + 1); ++i) split_forward(z, out, ny, tw.data(), log2m, i);
    for (size_t i = 0; i < (size_t)nf * M; ++i) { spec[2 * i] = out[i].x; spec[2 * i + 1] = out[i].y; }
}

// spec: nf x M complex, ny: nf -> x: nf frames of N reals (unnormalised inverse)
void harness_irfft(const float *spec, const float *ny, int log2m, int nf, float *x)
{
    const int M = 1 << log2m;
    auto tw = make_tw(log2m);
    std::vector<float2> acc((size_t)nf * M), a((size_t)nf * M), b((size_t)nf * M);
    for (size_t i = 0; i < (size_t)nf * M; ++i) acc[i] = make_float2(spec[2 * i], spec[2 * i + 1]);
    for (int i = 0; i < nf * (M / 2 + 1); ++i) split_inverse(acc.data(), ny, a.data(), tw.data(), log2m, i);
    float2 *z = cfft<true>(a.data(), b.data(), tw.data(), log2m, nf);
    for (size_t i = 0; i < (size_t)nf * M; ++i) { x[2 * i] = z[i].x; x[2 * i + 1] = z[i].y; }
}
}
