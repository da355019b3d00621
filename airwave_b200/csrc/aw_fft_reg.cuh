// aw_fft_reg.cuh — register-radix shared-memory FFT core (compile-time size).
//
// One M = 2^LOG2M point complex transform is computed by G = M/E cooperating threads
// (E = min(16, M) elements per thread).  Every pass is a Stockham step of radix R in {16, 8, 4, 2}:
//   load   v[q*R + r] = x[j + r*M/R],  j = t + q*G            (coalesced / conflict-free)
//   twiddle v[.] *= w_M^(r * (j mod Ns) * M/(Ns*R))           (table, skipped when Ns == 1)
//   DFT_R in registers (straight-line code, constants for W8 / W16)
//   store  y[(j div Ns)*Ns*R + (j mod Ns) + r*Ns] = V[r]      (padded shared memory)
// All E elements a thread owns are in registers between load and store, so a pass works in place
// on ONE shared buffer with a barrier after the loads and one after the stores.
// The first pass may load from anywhere (global memory) and the last pass may store anywhere,
// which is how the kernels fuse the overlap-save frame assembly and the "keep the second half"
// discard into the transform.
//
// Shared-memory padding: element i lives at i + (i >> 4) (one float2 of padding per 16), which
// makes the strided Stockham stores conflict-free for 64-bit accesses.
//
// The inverse transform is conj(FFT(conj(.))): callers conjugate on the way in and out.
// Replaces vDSP_fft_zrip (ConvolutionEngine.swift:174,252,353); twiddle table = plan cache entry.
#pragma once

#include "aw_fft.cuh"   // AW_HD, float2 helpers, awfft::cadd/csub/cmul

namespace awfft {

AW_HD int pad16(int i) { return i + (i >> 4); }
template <int LOG2M> struct PaddedSize { static constexpr int value = (1 << LOG2M) + ((1 << LOG2M) >> 4) + 1; };

// ---- straight-line DFTs, natural order in and out (forward sign) -----------------------------
AW_HD void dft2(float2 &a, float2 &b)
{
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

AW_HD void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3)
{
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = csub(v1, v3);
    const float2 r3 = make_float2(a3.y, -a3.x);   // -i * a3
    v0 = cadd(a0, a2);
    v1 = cadd(a1, r3);
    v2 = csub(a0, a2);
    v3 = csub(a1, r3);
}

AW_HD float2 mul_w8_1(float2 a) { const float c = 0.70710678118654752440f; return make_float2(c * (a.x + a.y), c * (a.y - a.x)); }    // * (1 - i)/sqrt2
AW_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }                                                                      // * (-i)
AW_HD float2 mul_w8_3(float2 a) { const float c = 0.70710678118654752440f; return make_float2(c * (a.y - a.x), -c * (a.x + a.y)); }   // * (-1 - i)/sqrt2

// 8-point: two 4-point transforms over even/odd samples + W8 twiddles
AW_HD void dft8(float2 (&v)[8])
{
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = mul_w8_1(o1);
    o2 = mul_mi(o2);
    o3 = mul_w8_3(o3);
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// 16-point as 4 x 4: X[k1 + 4*k2] = sum_n2 W4^(n2*k2) * W16^(n2*k1) * sum_n1 W4^(n1*k1) * x[4*n1 + n2]
AW_HD void dft16(float2 (&v)[16])
{
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin of pi/8
    const float c2 = 0.70710678118654752440f;
    float2 a[4][4];   // a[n2][k1]
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int n2 = 0; n2 < 4; ++n2) {
        a[n2][0] = v[n2]; a[n2][1] = v[4 + n2]; a[n2][2] = v[8 + n2]; a[n2][3] = v[12 + n2];
        dft4(a[n2][0], a[n2][1], a[n2][2], a[n2][3]);
    }
    // twiddles W16^(n2*k1), W16^m = (cos(m*pi/8), -sin(m*pi/8))
    a[1][1] = cmul(a[1][1], make_float2(c1, -s1));     // m = 1
    a[1][2] = cmul(a[1][2], make_float2(c2, -c2));     // m = 2
    a[1][3] = cmul(a[1][3], make_float2(s1, -c1));     // m = 3
    a[2][1] = cmul(a[2][1], make_float2(c2, -c2));     // m = 2
    a[2][2] = mul_mi(a[2][2]);                         // m = 4
    a[2][3] = cmul(a[2][3], make_float2(-c2, -c2));    // m = 6
    a[3][1] = cmul(a[3][1], make_float2(s1, -c1));     // m = 3
    a[3][2] = cmul(a[3][2], make_float2(-c2, -c2));    // m = 6
    a[3][3] = cmul(a[3][3], make_float2(-c1, s1));     // m = 9
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int k1 = 0; k1 < 4; ++k1) {
        float2 b0 = a[0][k1], b1 = a[1][k1], b2 = a[2][k1], b3 = a[3][k1];
        dft4(b0, b1, b2, b3);
        v[k1] = b0; v[k1 + 4] = b1; v[k1 + 8] = b2; v[k1 + 12] = b3;
    }
}

template <int R> struct Dft;
template <> struct Dft<2> { AW_HD static void run(float2 *v) { dft2(v[0], v[1]); } };
template <> struct Dft<4> { AW_HD static void run(float2 *v) { dft4(v[0], v[1], v[2], v[3]); } };
template <> struct Dft<8> { AW_HD static void run(float2 *v) { dft8(*reinterpret_cast<float2(*)[8]>(v)); } };
template <> struct Dft<16> { AW_HD static void run(float2 *v) { dft16(*reinterpret_cast<float2(*)[16]>(v)); } };

// ---- pass plan ------------------------------------------------------------------------------------
// log2 of the radix of pass `p` for a transform of 2^LOG2M points (0 = no such pass).
AW_HD constexpr int pass_log2r(int log2m, int p)
{
    // {16,...} greedy with a tail chosen to avoid tiny radices where possible
    // 2:{4} 3:{8} 4:{16} 5:{8,4} 6:{8,8} 7:{16,8} 8:{16,16} 9:{8,8,8} 10:{16,8,8} 11:{16,16,8} 12:{16,16,16} 13:{16,16,16,2}
    return log2m == 2 ? (p == 0 ? 2 : 0)
         : log2m == 3 ? (p == 0 ? 3 : 0)
         : log2m == 4 ? (p == 0 ? 4 : 0)
         : log2m == 5 ? (p == 0 ? 3 : p == 1 ? 2 : 0)
         : log2m == 6 ? (p < 2 ? 3 : 0)
         : log2m == 7 ? (p == 0 ? 4 : p == 1 ? 3 : 0)
         : log2m == 8 ? (p < 2 ? 4 : 0)
         : log2m == 9 ? (p < 3 ? 3 : 0)
         : log2m == 10 ? (p == 0 ? 4 : p < 3 ? 3 : 0)
         : log2m == 11 ? (p < 2 ? 4 : p == 2 ? 3 : 0)
         : log2m == 12 ? (p < 3 ? 4 : 0)
         : log2m == 13 ? (p < 3 ? 4 : p == 3 ? 1 : 0)
         : 0;
}
AW_HD constexpr int pass_count(int log2m) { return pass_log2r(log2m, 3) ? 4 : pass_log2r(log2m, 2) ? 3 : pass_log2r(log2m, 1) ? 2 : 1; }
AW_HD constexpr int pass_log2ns(int log2m, int p) { return p == 0 ? 0 : pass_log2ns(log2m, p - 1) + pass_log2r(log2m, p - 1); }

// Per-pass twiddle tables ("PT" layout).  The half-circle table is read with strides that are powers of two (index
// 2*r*k*M/(Ns*R) for the lanes' k), which serialises into up to 16 shared-memory wavefronts per load.  The PT layout stores, for
// every pass with Ns > 1, the values that pass needs as [r-1][k] (k contiguous, so the lanes of a warp read consecutive
// addresses): ptw[pass_tw_offset(P) + (r-1)*Ns + k] = w_M^(r*k*M/(Ns*R)).  A PT table starts with the M/2+1 entries the split
// step reads (k contiguous already), padded to PT_SPLIT entries.
AW_HD constexpr int pass_tw_entries(int log2m, int p) { return p == 0 ? 0 : ((1 << pass_log2r(log2m, p)) - 1) << pass_log2ns(log2m, p); }
AW_HD constexpr int pass_tw_offset(int log2m, int p) { return p <= 1 ? 0 : pass_tw_offset(log2m, p - 1) + pass_tw_entries(log2m, p - 1); }
AW_HD constexpr int pt_split_entries(int log2m) { return ((1 << log2m) / 2 + 1 + 15) & ~15; }
AW_HD constexpr int pt_total_entries(int log2m)
{
    return pt_split_entries(log2m) + pass_tw_offset(log2m, pass_count(log2m) - 1) + pass_tw_entries(log2m, pass_count(log2m) - 1);
}

template <int LOG2M>
struct RegFft {
    static constexpr int M = 1 << LOG2M;
    static constexpr int E = M < 16 ? M : 16;      // elements per thread
    static constexpr int G = M / E;                // threads per transform
    static constexpr int PASSES = pass_count(LOG2M);

    // element index (within the transform) that register slot `e` of thread `t` holds when LOADING pass P
    template <int P>
    AW_HD static int load_index(int t, int e)
    {
        constexpr int R = 1 << pass_log2r(LOG2M, P);
        const int q = e / R, r = e % R;
        return (t + q * G) + r * (M / R);
    }
    // element index register slot `e` (= q*R + output r) is STORED to after pass P
    template <int P>
    AW_HD static int store_index(int t, int e)
    {
        constexpr int R = 1 << pass_log2r(LOG2M, P);
        constexpr int Ns = 1 << pass_log2ns(LOG2M, P);
        const int q = e / R, r = e % R;
        const int j = t + q * G, k = j & (Ns - 1);
        return (j - k) * R + k + r * Ns;
    }
    // twiddles + in-register DFTs of pass P on the E values of thread t (tw = exp(-2*pi*i*k/(2M)), k < M)
    template <int P>
    AW_HD static void compute(float2 (&v)[E], const float2 *tw, int t)
    {
        constexpr int R = 1 << pass_log2r(LOG2M, P);
        constexpr int Ns = 1 << pass_log2ns(LOG2M, P);
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int q = 0; q < E / R; ++q) {
            if (Ns > 1) {
                const int j = t + q * G, k = j & (Ns - 1);
                const int step = k * (M / (Ns * R));   // w_M^(r*step)
#ifdef __CUDACC__
#pragma unroll
#endif
                for (int r = 1; r < R; ++r) {
                    int i2 = 2 * r * step;             // index into the half-circle table of N = 2M
                    float2 w = tw[i2 & (M - 1)];
                    if (i2 & M) { w.x = -w.x; w.y = -w.y; }
                    v[q * R + r] = cmul(v[q * R + r], w);
                }
            }
            Dft<R>::run(&v[q * R]);
        }
    }
    // same as compute<P>, twiddles from a PT table (pt = table base; the pass tables start at pt + pt_split_entries)
    // PRODUCTS: a radix-16 pass reads six twiddles and multiplies (w^r = w^(r & 3) * w^(r & 12)) instead of reading fifteen: for
    // callers whose table is in global memory (K2, K4: measured 9% faster at B = 4096, neutral where the table is in shared
    // memory); costs 0.5 dB of the transform's 138 dB.
    template <int P, bool PRODUCTS = false>
    AW_HD static void compute_pt(float2 (&v)[E], const float2 *pt, int t)
    {
        constexpr int R = 1 << pass_log2r(LOG2M, P);
        constexpr int Ns = 1 << pass_log2ns(LOG2M, P);
        const float2 *tbl = pt + pt_split_entries(LOG2M) + pass_tw_offset(LOG2M, P);
#ifdef __CUDACC__
#pragma unroll
#endif
        for (int q = 0; q < E / R; ++q) {
            if (Ns > 1) {
                const int k = (t + q * G) & (Ns - 1);
                if constexpr (PRODUCTS && R == 16) {
                    float2 lo[4], hi[4];
#ifdef __CUDACC__
#pragma unroll
#endif
                    for (int i = 1; i < 4; ++i) { lo[i] = tbl[(i - 1) * Ns + k]; hi[i] = tbl[(4 * i - 1) * Ns + k]; }
#ifdef __CUDACC__
#pragma unroll
#endif
                    for (int r = 1; r < R; ++r) {
                        const float2 w = (r & 3) == 0 ? hi[r >> 2] : (r >> 2) == 0 ? lo[r & 3] : cmul(lo[r & 3], hi[r >> 2]);
                        v[q * R + r] = cmul(v[q * R + r], w);
                    }
                } else {
#ifdef __CUDACC__
#pragma unroll
#endif
                    for (int r = 1; r < R; ++r) v[q * R + r] = cmul(v[q * R + r], tbl[(r - 1) * Ns + k]);
                }
            }
            Dft<R>::run(&v[q * R]);
        }
    }
    // entry i of the PT table from the half-circle table hc[k] = exp(-2*pi*i*k/(2M)), k < M
    AW_HD static float2 pt_entry(const float2 *hc, int i)
    {
        if (i < pt_split_entries(LOG2M)) return hc[i < M ? i : M - 1];
        i -= pt_split_entries(LOG2M);
        float2 w = make_float2(1.f, 0.f);
        for (int p = 1; p < PASSES; ++p) {
            const int n = pass_tw_entries(LOG2M, p);
            if (i < n) {
                const int log2ns = pass_log2ns(LOG2M, p), log2r = pass_log2r(LOG2M, p);
                const int r = 1 + (i >> log2ns), k = i & ((1 << log2ns) - 1);
                const int i2 = 2 * r * (k * (M >> (log2ns + log2r)));
                w = hc[i2 & (M - 1)];
                if (i2 & M) { w.x = -w.x; w.y = -w.y; }
                return w;
            }
            i -= n;
        }
        return w;
    }
};

// ---- CPU/GPU-neutral reference driver: whole transform on one padded buffer (used by the CPU harness
// and as documentation of the calling sequence; kernels inline the same steps with barriers) -----
template <int LOG2M, int P>
struct PassRunner {
    // runs pass P for "thread" t on buffer buf (padded); tmp holds the thread's registers
    AW_HD static void load(const float2 *buf, float2 (&v)[RegFft<LOG2M>::E], int t)
    {
        for (int e = 0; e < RegFft<LOG2M>::E; ++e) v[e] = buf[pad16(RegFft<LOG2M>::template load_index<P>(t, e))];
    }
    AW_HD static void store(float2 *buf, const float2 (&v)[RegFft<LOG2M>::E], int t)
    {
        for (int e = 0; e < RegFft<LOG2M>::E; ++e) buf[pad16(RegFft<LOG2M>::template store_index<P>(t, e))] = v[e];
    }
};

}  // namespace awfft
