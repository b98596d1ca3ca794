// hb_fft_core.cuh -- per-thread building blocks of the shared-memory FFT, written as
// __host__ __device__ functions of (thread id, thread count) so that tests/host_emul can run the
// exact index arithmetic on the CPU (one loop per barrier-separated phase) before it ever meets a GPU.
//
// Conventions are those of HISSTools_FFT (reference: HISSTools_FFT/HISSTools_FFT_Core.h):
//   forward kernel exp(-j theta), unscaled                                  (Core:439-443, 669-675)
//   real forward of N = 2M points = complex FFT of M + split pass -> 2*DFT,
//   DC in re[0], Nyquist in im[0]                                           (Core:934-988, 1350-1360)
//   real inverse = reverse split pass + planes-exchanged complex FFT        (Core:1341-1346, 1364-1374)
// The algorithm is our own: a Stockham autosort radix-8/4/2 FFT done in place in shared memory
// (every thread pulls all its butterfly inputs into registers, barrier, scatters the outputs).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define HB_HD __host__ __device__ __forceinline__
#define HB_ALIGN(n) __align__(n)
#else
#define HB_HD inline
#define HB_ALIGN(n) alignas(n)
#endif

namespace hb
{

template <class T> struct Cx;
template <> struct HB_ALIGN(8) Cx<float> { float x, y; };
template <> struct HB_ALIGN(16) Cx<double> { double x, y; };

template <class T> HB_HD Cx<T> cx(T x, T y) { Cx<T> r; r.x = x; r.y = y; return r; }
template <class T> HB_HD Cx<T> cadd(Cx<T> a, Cx<T> b) { return cx<T>(a.x + b.x, a.y + b.y); }
template <class T> HB_HD Cx<T> csub(Cx<T> a, Cx<T> b) { return cx<T>(a.x - b.x, a.y - b.y); }
template <class T> HB_HD Cx<T> cmul(Cx<T> a, Cx<T> b) { return cx<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by -i / +i
template <class T> HB_HD Cx<T> mul_mi(Cx<T> a) { return cx<T>(a.y, -a.x); }
template <class T> HB_HD Cx<T> mul_pi(Cx<T> a) { return cx<T>(-a.y, a.x); }

// Twiddle table: tw[q] = exp(-2 pi i q / 2^tw_log2) for q in [0, 2^(tw_log2-1)) (half circle),
// computed in double on the host and rounded to T like the reference's tables (Core:414-448).
// root(a, l) = exp(-2 pi i a / 2^l), a < 2^l, l <= tw_log2.
template <class T>
HB_HD Cx<T> tw_root(const Cx<T> *tw, int tw_log2, uint32_t a, int l)
{
    uint32_t q = a << (tw_log2 - l);
    uint32_t half = 1u << (tw_log2 - 1);
    if (q >= half)
    {
        Cx<T> w = tw[q - half];
        return cx<T>(-w.x, -w.y);
    }
    return tw[q];
}

// shared-memory index with optional padding (one extra slot every 2^PADSH elements)
template <int PADSH> HB_HD uint32_t sidx(uint32_t i) { return PADSH >= 31 ? i : i + (i >> PADSH); }
template <int PADSH> HB_HD uint32_t padded_elems(uint32_t m) { return PADSH >= 31 ? m : m + (m >> PADSH) + 1; }

// ---- in-register forward DFTs of 2, 4, 8 points (natural order in and out) -------------------------
template <class T> HB_HD void dft2(Cx<T> &a, Cx<T> &b)
{
    Cx<T> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

template <class T> HB_HD void dft4(Cx<T> &v0, Cx<T> &v1, Cx<T> &v2, Cx<T> &v3)
{
    Cx<T> a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi(csub(v1, v3));
    v0 = cadd(a0, a2);
    v2 = csub(a0, a2);
    v1 = cadd(a1, a3);
    v3 = csub(a1, a3);
}

template <class T> HB_HD void dft8(Cx<T> *v)
{
    const T h = (T) 0.70710678118654752440084436210484903928;
    dft4(v[0], v[2], v[4], v[6]);          // E[0..3] in v0,v2,v4,v6
    dft4(v[1], v[3], v[5], v[7]);          // O[0..3] in v1,v3,v5,v7
    Cx<T> o1 = cx<T>((v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h);     // O1 * (1 - i)/sqrt2
    Cx<T> o2 = mul_mi(v[5]);                                            // O2 * (-i)
    Cx<T> o3 = cx<T>((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);    // O3 * (-1 - i)/sqrt2
    Cx<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

template <class T, int R> struct Dft;
template <class T> struct Dft<T, 2> { static HB_HD void run(Cx<T> *v) { dft2(v[0], v[1]); } };
template <class T> struct Dft<T, 4> { static HB_HD void run(Cx<T> *v) { dft4(v[0], v[1], v[2], v[3]); } };
template <class T> struct Dft<T, 8> { static HB_HD void run(Cx<T> *v) { dft8(v); } };

// ---- one Stockham pass, split at the barrier ---------------------------------------------------------
// M points, radix R, Ns = product of the radices of the earlier passes.  Each thread owns EPT points
// (EPT/R butterflies); butterfly j reads s[j + t*M/R] (t < R) and writes s[expand(j) + t*Ns].
// `nthr` threads take part; requires nthr * EPT == M (or, for M < EPT, nthr == 1 and the guard below).
template <class T, int EPT, int R, int PADSH>
HB_HD void pass_load(const Cx<T> *s, uint32_t M, uint32_t tid, uint32_t nthr, Cx<T> *v)
{
    const uint32_t nb = M / R;             // butterflies in this pass
#pragma unroll
    for (int b = 0; b < EPT / R; b++)
    {
        uint32_t j = tid + b * nthr;
        if (j < nb)
        {
#pragma unroll
            for (int t = 0; t < R; t++) v[b * R + t] = s[sidx<PADSH>(j + t * nb)];
        }
    }
}

template <class T, int EPT, int R, int PADSH>
HB_HD void pass_store(Cx<T> *s, uint32_t M, uint32_t Ns, int log2_nsr, uint32_t tid, uint32_t nthr, Cx<T> *v,
                      const Cx<T> *tw, int tw_log2)
{
    const uint32_t nb = M / R;
#pragma unroll
    for (int b = 0; b < EPT / R; b++)
    {
        uint32_t j = tid + b * nthr;
        if (j < nb)
        {
            uint32_t k = j & (Ns - 1);
            if (Ns > 1)
            {
#pragma unroll
                for (int t = 1; t < R; t++) v[b * R + t] = cmul(v[b * R + t], tw_root(tw, tw_log2, k * t, log2_nsr));
            }
            Dft<T, R>::run(v + b * R);
            uint32_t j0 = ((j - k) * R) + k;
#pragma unroll
            for (int t = 0; t < R; t++) s[sidx<PADSH>(j0 + t * Ns)] = v[b * R + t];
        }
    }
}

// radix schedule for log2(M) = m: as many radix-8 passes as possible, never ending on a lone 2 after
// an 8 when a 4,4 split is available.  Returns the radix of pass `idx` given remaining bits.
HB_HD int next_radix(int bits_left)
{
    if (bits_left >= 3 && bits_left != 4) return 8;
    if (bits_left >= 2) return 4;
    return 2;
}

// ---- real <-> half-complex split passes (in place on the M-point complex array in shared memory) -----
// forward: Z = FFT_M(z), z[k] = x[2k] + i x[2k+1]  ->  packed 2*DFT_N(x).  Pair index k in [1, M/2];
// k == 0 handled by pair 0.  inverse: the exact algebraic reverse without the DC doubling.
// (behaviour of pass_real_trig_table<ifft>: Core:934-988)
template <class T, int PADSH>
HB_HD void real_split_pair(Cx<T> *s, uint32_t M, int log2N, uint32_t k, bool inverse, const Cx<T> *tw, int tw_log2)
{
    if (k == 0)
    {
        Cx<T> z = s[sidx<PADSH>(0)];
        T t1 = z.x + z.y, t2 = z.x - z.y;
        s[sidx<PADSH>(0)] = inverse ? cx<T>(t1, t2) : cx<T>(t1 + t1, t2 + t2);
        return;
    }
    uint32_t q = M - k;
    Cx<T> w = tw_root(tw, tw_log2, k, log2N);
    Cx<T> a = s[sidx<PADSH>(k)], b = s[sidx<PADSH>(q)];
    T sr = a.x + b.x, si = a.y + b.y, dr = a.x - b.x, di = a.y - b.y;
    T wr = inverse ? -w.x : w.x, wi = w.y;
    T u = wr * si + wi * dr;
    T v = wi * si - wr * dr;
    s[sidx<PADSH>(k)] = cx<T>(sr + u, v + di);
    s[sidx<PADSH>(q)] = cx<T>(sr - u, v - di);
}

// ---- work decomposition of the multiply-accumulate kernel (stream-K over contiguous IR units) -------
// U units are dealt to G CTAs as contiguous ranges [unit_begin(g), unit_begin(g+1)).
HB_HD uint64_t unit_begin(uint64_t g, uint64_t U, uint64_t G) { return (g * U) / G; }
// the CTA that owns unit u (largest g with unit_begin(g) <= u)
HB_HD uint64_t unit_owner(uint64_t u, uint64_t U, uint64_t G) { return ((u + 1) * G - 1) / U; }

} // namespace hb
