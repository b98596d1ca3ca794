// tests/host_emul/emul_fft.cpp -- CPU emulation of the CUDA shared-memory FFT (test infrastructure).
//
// Compiles hisstools_library_b200/csrc/hb_fft_core.cuh as plain C++ and drives the very same
// per-thread functions the kernels use, one loop over "threads" per barrier-separated phase.  This
// checks the index arithmetic of the Stockham passes, the split passes and the stream-K unit
// decomposition in the CPU test-suite, where no GPU exists.  It is NOT a CPU fallback: nothing in
// the product links it.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../hisstools_library_b200/csrc/hb_fft_core.cuh"

using namespace hb;

template <class T>
static std::vector<Cx<T>> make_tw(int tw_log2)
{
    size_t half = tw_log2 > 0 ? (size_t(1) << (tw_log2 - 1)) : 1;
    std::vector<Cx<T>> tw(half);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (size_t q = 0; q < half; q++)
    {
        long double a = -2.0L * pi * (long double) q / (long double) (size_t(1) << tw_log2);
        tw[q].x = (T) cosl(a);
        tw[q].y = (T) sinl(a);
    }
    return tw;
}

template <class T, int EPT, int PADSH>
static void emul_block_fft(Cx<T> *s, int log2m, uint32_t nthr, const Cx<T> *tw, int tw_log2)
{
    const uint32_t M = 1u << log2m;
    std::vector<Cx<T>> regs(size_t(nthr) * EPT);
    uint32_t Ns = 1;
    int done = 0;
    while (done < log2m)
    {
        const int R = next_radix(log2m - done);
        for (uint32_t tid = 0; tid < nthr; tid++)
        {
            Cx<T> *v = &regs[size_t(tid) * EPT];
            if (R == 8) pass_load<T, EPT, 8, PADSH>(s, M, tid, nthr, v);
            else if (R == 4) pass_load<T, EPT, 4, PADSH>(s, M, tid, nthr, v);
            else pass_load<T, EPT, 2, PADSH>(s, M, tid, nthr, v);
        }
        for (uint32_t tid = 0; tid < nthr; tid++)
        {
            Cx<T> *v = &regs[size_t(tid) * EPT];
            if (R == 8) pass_store<T, EPT, 8, PADSH>(s, M, Ns, done + 3, tid, nthr, v, tw, tw_log2);
            else if (R == 4) pass_store<T, EPT, 4, PADSH>(s, M, Ns, done + 2, tid, nthr, v, tw, tw_log2);
            else pass_store<T, EPT, 2, PADSH>(s, M, Ns, done + 1, tid, nthr, v, tw, tw_log2);
        }
        done += (R == 8 ? 3 : R == 4 ? 2 : 1);
        Ns *= R;
    }
}

template <class T, int EPT, int PADSH>
static void emul_cfft(T *re, T *im, int log2m, uint32_t nthr)
{
    const uint32_t M = 1u << log2m;
    int tw_log2 = log2m > 1 ? log2m : 1;
    auto tw = make_tw<T>(tw_log2);
    std::vector<Cx<T>> s(padded_elems<PADSH>(M));
    for (uint32_t i = 0; i < M; i++) s[sidx<PADSH>(i)] = cx<T>(re[i], im[i]);
    emul_block_fft<T, EPT, PADSH>(s.data(), log2m, nthr, tw.data(), tw_log2);
    for (uint32_t i = 0; i < M; i++) { re[i] = s[sidx<PADSH>(i)].x; im[i] = s[sidx<PADSH>(i)].y; }
}

// in-place real transforms on split planes of M = 2^(log2n-1) points, as the kernels order them
template <class T, int EPT, int PADSH>
static void emul_rfft(T *re, T *im, int log2n, int inverse, uint32_t nthr)
{
    const int log2m = log2n - 1;
    const uint32_t M = 1u << log2m;
    int tw_log2 = log2n;
    auto tw = make_tw<T>(tw_log2);
    std::vector<Cx<T>> s(padded_elems<PADSH>(M));
    for (uint32_t i = 0; i < M; i++) s[sidx<PADSH>(i)] = cx<T>(re[i], im[i]);
    if (!inverse)
    {
        emul_block_fft<T, EPT, PADSH>(s.data(), log2m, nthr, tw.data(), tw_log2);
        for (uint32_t k = 0; k <= M / 2; k++) real_split_pair<T, PADSH>(s.data(), M, log2n, k, false, tw.data(), tw_log2);
    }
    else
    {
        for (uint32_t k = 0; k <= M / 2; k++) real_split_pair<T, PADSH>(s.data(), M, log2n, k, true, tw.data(), tw_log2);
        for (uint32_t i = 0; i < M; i++) { Cx<T> v = s[sidx<PADSH>(i)]; s[sidx<PADSH>(i)] = cx<T>(v.y, v.x); }
        emul_block_fft<T, EPT, PADSH>(s.data(), log2m, nthr, tw.data(), tw_log2);
        for (uint32_t i = 0; i < M; i++) { Cx<T> v = s[sidx<PADSH>(i)]; s[sidx<PADSH>(i)] = cx<T>(v.y, v.x); }
    }
    for (uint32_t i = 0; i < M; i++) { re[i] = s[sidx<PADSH>(i)].x; im[i] = s[sidx<PADSH>(i)].y; }
}

extern "C"
{
void emul_cfft_f32(float *re, float *im, int log2m, int ept, uint32_t nthr)
{ if (ept == 16) emul_cfft<float, 16, 5>(re, im, log2m, nthr); else emul_cfft<float, 8, 5>(re, im, log2m, nthr); }
void emul_cfft_f64(double *re, double *im, int log2m, int ept, uint32_t nthr)
{ if (ept == 16) emul_cfft<double, 16, 5>(re, im, log2m, nthr); else emul_cfft<double, 8, 5>(re, im, log2m, nthr); }
void emul_rfft_f32(float *re, float *im, int log2n, int inverse, int ept, uint32_t nthr)
{ if (ept == 16) emul_rfft<float, 16, 5>(re, im, log2n, inverse, nthr); else emul_rfft<float, 8, 5>(re, im, log2n, inverse, nthr); }
void emul_rfft_f64(double *re, double *im, int log2n, int inverse, int ept, uint32_t nthr)
{ if (ept == 16) emul_rfft<double, 16, 5>(re, im, log2n, inverse, nthr); else emul_rfft<double, 8, 5>(re, im, log2n, inverse, nthr); }

uint64_t emul_unit_begin(uint64_t g, uint64_t U, uint64_t G) { return unit_begin(g, U, G); }
uint64_t emul_unit_owner(uint64_t u, uint64_t U, uint64_t G) { return unit_owner(u, U, G); }
}
