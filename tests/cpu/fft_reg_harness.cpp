// CPU harness for airwave_b200/csrc/aw_fft_reg.cuh: emulates the G cooperating threads of one transform
// (load all -> compute -> store all per pass), so pass plans, index maps, twiddles and the straight-line
// DFT_2/4/8/16 are validated without a GPU.
#include <cmath>
#include <cstddef>
#include <vector>
#include "../../airwave_b200/csrc/aw_fft_reg.cuh"

using namespace awfft;

// `pt`: 1 = use the per-pass twiddle tables (RegFft::pt_entry layout, compute_pt) instead of the half-circle table;
// 2 = the same tables, radix-16 twiddles as six reads and nine products (K2 / K4)
static int g_use_pt = 0;

template <int LOG2M, int P>
struct RunPasses {
    static void run(std::vector<float2> &buf, const float2 *tw)
    {
        using F = RegFft<LOG2M>;
        if (P >= F::PASSES) return;
        std::vector<float2> regs((size_t)F::G * F::E);
        for (int t = 0; t < F::G; ++t) {
            float2(&v)[F::E] = *reinterpret_cast<float2(*)[F::E]>(&regs[(size_t)t * F::E]);
            PassRunner<LOG2M, (P < F::PASSES ? P : 0)>::load(buf.data(), v, t);
        }
        for (int t = 0; t < F::G; ++t) {
            float2(&v)[F::E] = *reinterpret_cast<float2(*)[F::E]>(&regs[(size_t)t * F::E]);
            if (g_use_pt == 2) F::template compute_pt<(P < F::PASSES ? P : 0), true>(v, tw, t);
            else if (g_use_pt) F::template compute_pt<(P < F::PASSES ? P : 0)>(v, tw, t);
            else F::template compute<(P < F::PASSES ? P : 0)>(v, tw, t);
            PassRunner<LOG2M, (P < F::PASSES ? P : 0)>::store(buf.data(), v, t);
        }
        RunPasses<LOG2M, (P + 1 < 4 ? P + 1 : 4)>::run(buf, tw);
    }
};
template <int LOG2M>
struct RunPasses<LOG2M, 4> { static void run(std::vector<float2> &, const float2 *) {} };

template <int LOG2M>
static void fft_one(const float *in, float *out)
{
    constexpr int M = 1 << LOG2M;
    std::vector<float2> tw(M);
    for (int k = 0; k < M; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)(2 * M);
        tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<float2> buf(PaddedSize<LOG2M>::value);
    for (int i = 0; i < M; ++i) buf[pad16(i)] = make_float2(in[2 * i], in[2 * i + 1]);
    std::vector<float2> pt((size_t)pt_total_entries(LOG2M));
    for (int i = 0; i < (int)pt.size(); ++i) pt[i] = RegFft<LOG2M>::pt_entry(tw.data(), i);
    RunPasses<LOG2M, 0>::run(buf, g_use_pt ? pt.data() : tw.data());
    for (int i = 0; i < M; ++i) { out[2 * i] = buf[pad16(i)].x; out[2 * i + 1] = buf[pad16(i)].y; }
}

extern "C" void harness_regfft_use_pt(int mode) { g_use_pt = mode; }

extern "C" int harness_regfft(const float *in, int log2m, float *out)
{
    switch (log2m) {
    case 2: fft_one<2>(in, out); break;   case 3: fft_one<3>(in, out); break;   case 4: fft_one<4>(in, out); break;
    case 5: fft_one<5>(in, out); break;   case 6: fft_one<6>(in, out); break;   case 7: fft_one<7>(in, out); break;
    case 8: fft_one<8>(in, out); break;   case 9: fft_one<9>(in, out); break;   case 10: fft_one<10>(in, out); break;
    case 11: fft_one<11>(in, out); break; case 12: fft_one<12>(in, out); break; case 13: fft_one<13>(in, out); break;
    default: return 1;
    }
    return 0;
}
