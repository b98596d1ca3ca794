// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A thin extern "C" window onto the UNMODIFIED reference, compiled from the sources where
// they lie under $REF_ROOT (default /root/reference) by oracle/Makefile into
// oracle/_ref/libhisstools_ref.so.  Nothing from the reference is copied into this
// repository: this file only #includes the reference headers and forwards calls.
//
// Exposed (all reference symbols cited as file:line under /root/reference):
//   * FFT family            HISSTools_FFT/HISSTools_FFT.h:87-369
//   * PartitionedConvolve   HIRT_Multichannel_Convolution/PartitionedConvolve.h:23-41
//   * MonoConvolve          HIRT_Multichannel_Convolution/MonoConvolve.h:30-48
//   * NToMonoConvolve       HIRT_Multichannel_Convolution/NToMonoConvolve.h:18-24
//   * Convolver             HIRT_Multichannel_Convolution/Convolver.h:25-50
//   * a uniform-partition N x M matrix assembled from MonoConvolve(maxLen,false,A)
//     exactly as NToMonoConvolve.cpp:35-43 sums them (BASELINE configs 3/4)
//   * PConvRestated<T>: the double-precision oracle of SURVEY 8c -- the state machine of
//     PartitionedConvolve.cpp:173-426 restated over the reference's own FFT_SETUP_D calls
//     (the reference class is float-only, PartitionedConvolve.h:38-41); its float
//     instantiation is cross-checked against the real class in tests/.
//   * CPU timing helpers used for bench.py's cpu_baseline / --impl reference legs.

#include "HISSTools_FFT/HISSTools_FFT.h"
#include "HIRT_Multichannel_Convolution/PartitionedConvolve.h"
#include "HIRT_Multichannel_Convolution/MonoConvolve.h"
#include "HIRT_Multichannel_Convolution/NToMonoConvolve.h"
#include "HIRT_Multichannel_Convolution/Convolver.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#if defined(__SSE__)
#include <xmmintrin.h>
#endif

#define SHIM extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------
// FFT family
// ---------------------------------------------------------------------------------------------

SHIM void *ref_fft_setup_f32(uintptr_t max_log2) { FFT_SETUP_F s; hisstools_create_setup(&s, max_log2); return s; }
SHIM void *ref_fft_setup_f64(uintptr_t max_log2) { FFT_SETUP_D s; hisstools_create_setup(&s, max_log2); return s; }
SHIM void ref_fft_setup_free_f32(void *s) { hisstools_destroy_setup(static_cast<FFT_SETUP_F>(s)); }
SHIM void ref_fft_setup_free_f64(void *s) { hisstools_destroy_setup(static_cast<FFT_SETUP_D>(s)); }

#define SHIM_INPLACE(NAME, FN)                                                                          \
    SHIM void ref_##NAME##_f32(void *s, float *re, float *im, uintptr_t log2n)                        \
    { FFT_SPLIT_COMPLEX_F sp(re, im); FN(static_cast<FFT_SETUP_F>(s), &sp, log2n); }                  \
    SHIM void ref_##NAME##_f64(void *s, double *re, double *im, uintptr_t log2n)                      \
    { FFT_SPLIT_COMPLEX_D sp(re, im); FN(static_cast<FFT_SETUP_D>(s), &sp, log2n); }

SHIM_INPLACE(fft, hisstools_fft)
SHIM_INPLACE(ifft, hisstools_ifft)
SHIM_INPLACE(rfft, hisstools_rfft)
SHIM_INPLACE(rifft, hisstools_rifft)

SHIM void ref_rfft_real_f32(void *s, const float *in, float *re, float *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_F sp(re, im); hisstools_rfft(static_cast<FFT_SETUP_F>(s), in, &sp, in_length, log2n); }
SHIM void ref_rfft_real_f64(void *s, const double *in, double *re, double *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_rfft(static_cast<FFT_SETUP_D>(s), in, &sp, in_length, log2n); }
SHIM void ref_rfft_real_f32_f64(void *s, const float *in, double *re, double *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_rfft(static_cast<FFT_SETUP_D>(s), in, &sp, in_length, log2n); }
SHIM void ref_rifft_real_f32(void *s, float *re, float *im, float *out, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_F sp(re, im); hisstools_rifft(static_cast<FFT_SETUP_F>(s), &sp, out, log2n); }
SHIM void ref_rifft_real_f64(void *s, double *re, double *im, double *out, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_rifft(static_cast<FFT_SETUP_D>(s), &sp, out, log2n); }

SHIM void ref_unzip_f32(const float *in, float *re, float *im, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_F sp(re, im); hisstools_unzip(in, &sp, log2n); }
SHIM void ref_unzip_f64(const double *in, double *re, double *im, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_unzip(in, &sp, log2n); }
SHIM void ref_zip_f32(const float *re, const float *im, float *out, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_F sp(const_cast<float *>(re), const_cast<float *>(im)); hisstools_zip(&sp, out, log2n); }
SHIM void ref_zip_f64(const double *re, const double *im, double *out, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(const_cast<double *>(re), const_cast<double *>(im)); hisstools_zip(&sp, out, log2n); }
SHIM void ref_unzip_zero_f32(const float *in, float *re, float *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_F sp(re, im); hisstools_unzip_zero(in, &sp, in_length, log2n); }
SHIM void ref_unzip_zero_f64(const double *in, double *re, double *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_unzip_zero(in, &sp, in_length, log2n); }
SHIM void ref_unzip_zero_f32_f64(const float *in, double *re, double *im, uintptr_t in_length, uintptr_t log2n)
{ FFT_SPLIT_COMPLEX_D sp(re, im); hisstools_unzip_zero(in, &sp, in_length, log2n); }

// ---------------------------------------------------------------------------------------------
// PartitionedConvolve
// ---------------------------------------------------------------------------------------------

using HISSTools::PartitionedConvolve;
using HISSTools::MonoConvolve;
using HISSTools::NToMonoConvolve;
using HISSTools::Convolver;

SHIM void *ref_pconv_create(uintptr_t maxFFT, uintptr_t maxLen, uintptr_t offset, uintptr_t length)
{ return new PartitionedConvolve(maxFFT, maxLen, offset, length); }
SHIM void ref_pconv_destroy(void *p) { delete static_cast<PartitionedConvolve *>(p); }
SHIM int ref_pconv_set_fft_size(void *p, uintptr_t n) { return static_cast<PartitionedConvolve *>(p)->setFFTSize(n); }
SHIM int ref_pconv_set_length(void *p, uintptr_t n) { return static_cast<PartitionedConvolve *>(p)->setLength(n); }
SHIM void ref_pconv_set_offset(void *p, uintptr_t n) { static_cast<PartitionedConvolve *>(p)->setOffset(n); }
SHIM void ref_pconv_set_reset_offset(void *p, intptr_t n) { static_cast<PartitionedConvolve *>(p)->setResetOffset(n); }
SHIM int ref_pconv_set(void *p, const float *ir, uintptr_t len) { return static_cast<PartitionedConvolve *>(p)->set(ir, len); }
SHIM void ref_pconv_reset(void *p) { static_cast<PartitionedConvolve *>(p)->reset(); }
SHIM int ref_pconv_process(void *p, const float *in, float *out, uintptr_t n)
{ return static_cast<PartitionedConvolve *>(p)->process(in, out, n) ? 1 : 0; }

// ---------------------------------------------------------------------------------------------
// MonoConvolve
// ---------------------------------------------------------------------------------------------

SHIM void *ref_mono_create_latency(uintptr_t maxLen, int latency)
{
    try { return new MonoConvolve(maxLen, static_cast<LatencyMode>(latency)); } catch (...) { return nullptr; }
}
SHIM void *ref_mono_create_custom(uintptr_t maxLen, int zeroLatency, uint32_t A, uint32_t B, uint32_t C, uint32_t D)
{
    try { return new MonoConvolve(maxLen, zeroLatency != 0, A, B, C, D); } catch (...) { return nullptr; }
}
SHIM void ref_mono_destroy(void *p) { delete static_cast<MonoConvolve *>(p); }
SHIM void ref_mono_set_reset_offset(void *p, intptr_t off) { static_cast<MonoConvolve *>(p)->setResetOffset(off); }
SHIM int ref_mono_resize(void *p, uintptr_t len) { return static_cast<MonoConvolve *>(p)->resize(len); }
SHIM int ref_mono_set(void *p, const float *ir, uintptr_t len, int resize)
{ return static_cast<MonoConvolve *>(p)->set(ir, len, resize != 0); }
SHIM int ref_mono_reset(void *p) { return static_cast<MonoConvolve *>(p)->reset(); }
SHIM void ref_mono_process(void *p, const float *in, float *temp, float *out, uintptr_t n, int accumulate)
{ static_cast<MonoConvolve *>(p)->process(in, temp, out, n, accumulate != 0); }

// ---------------------------------------------------------------------------------------------
// NToMonoConvolve / Convolver (the shipped LatencyMode constructors)
// ---------------------------------------------------------------------------------------------

SHIM void *ref_n2m_create(uint32_t inChans, uintptr_t maxLen, int latency)
{ return new NToMonoConvolve(inChans, maxLen, static_cast<LatencyMode>(latency)); }
SHIM void ref_n2m_destroy(void *p) { delete static_cast<NToMonoConvolve *>(p); }
SHIM int ref_n2m_resize(void *p, uint32_t in, uintptr_t len) { return static_cast<NToMonoConvolve *>(p)->resize(in, len); }
SHIM int ref_n2m_set(void *p, uint32_t in, const float *ir, uintptr_t len, int resize)
{ return static_cast<NToMonoConvolve *>(p)->set(in, ir, len, resize != 0); }
SHIM int ref_n2m_reset(void *p, uint32_t in) { return static_cast<NToMonoConvolve *>(p)->reset(in); }
SHIM void ref_n2m_process(void *p, const float *const *ins, float *out, float *temp, size_t n, size_t active)
{ static_cast<NToMonoConvolve *>(p)->process(ins, out, temp, n, active); }

SHIM void *ref_conv_create(uint32_t nIn, uint32_t nOut, int latency)
{ return new Convolver(nIn, nOut, static_cast<LatencyMode>(latency)); }
SHIM void *ref_conv_create_parallel(uint32_t nIO, int latency)
{ return new Convolver(nIO, static_cast<LatencyMode>(latency)); }
SHIM void ref_conv_destroy(void *p) { delete static_cast<Convolver *>(p); }
SHIM void ref_conv_clear(void *p, int resize) { static_cast<Convolver *>(p)->clear(resize != 0); }
SHIM void ref_conv_clear_chan(void *p, uint32_t in, uint32_t out, int resize) { static_cast<Convolver *>(p)->clear(in, out, resize != 0); }
SHIM void ref_conv_reset(void *p) { static_cast<Convolver *>(p)->reset(); }
SHIM int ref_conv_reset_chan(void *p, uint32_t in, uint32_t out) { return static_cast<Convolver *>(p)->reset(in, out); }
SHIM int ref_conv_resize(void *p, uint32_t in, uint32_t out, uintptr_t len) { return static_cast<Convolver *>(p)->resize(in, out, len); }
SHIM int ref_conv_set_f32(void *p, uint32_t in, uint32_t out, const float *ir, uintptr_t len, int resize)
{ return static_cast<Convolver *>(p)->set(in, out, ir, len, resize != 0); }
SHIM int ref_conv_set_f64(void *p, uint32_t in, uint32_t out, const double *ir, uintptr_t len, int resize)
{ return static_cast<Convolver *>(p)->set(in, out, ir, len, resize != 0); }
SHIM void ref_conv_process_f32(void *p, const float *const *ins, float **outs, size_t nIn, size_t nOut, size_t n)
{ static_cast<Convolver *>(p)->process(ins, outs, nIn, nOut, n); }
SHIM void ref_conv_process_f64(void *p, const double *const *ins, double **outs, size_t nIn, size_t nOut, size_t n)
{ static_cast<Convolver *>(p)->process(ins, outs, nIn, nOut, n); }

// The shipped MonoConvolve staggers the reset phase of its parts from a random base
// (MonoConvolve.cpp:80-98); tests need the deterministic one.  NToMonoConvolve / Convolver do not
// expose setResetOffset, so determinism there comes only from results being phase-independent
// to rounding -- the tests use tolerances, not bit equality, for those classes.

// ---------------------------------------------------------------------------------------------
// Uniform-partition N x M matrix (BASELINE configs 3 and 4), assembled as NToMonoConvolve.cpp:35-43
// does it: zero the output row, then every input's MonoConvolve accumulates into it.
// ---------------------------------------------------------------------------------------------

struct RefMatrix
{
    uint32_t nIn, nOut;
    bool parallel;
    std::vector<std::unique_ptr<MonoConvolve>> pairs;   // [out][in] (parallel: [chan])
    std::vector<float> temp;
};

SHIM void *ref_matrix_create(uint32_t nIn, uint32_t nOut, uintptr_t maxLen, uint32_t fftSize, int parallel)
{
    try
    {
        auto *m = new RefMatrix{nIn, nOut, parallel != 0, {}, {}};
        size_t count = parallel ? nOut : size_t(nIn) * nOut;
        for (size_t i = 0; i < count; i++)
        {
            m->pairs.emplace_back(new MonoConvolve(maxLen, false, fftSize));
            m->pairs.back()->setResetOffset(0);
        }
        return m;
    }
    catch (...) { return nullptr; }
}
SHIM void ref_matrix_destroy(void *p) { delete static_cast<RefMatrix *>(p); }
SHIM int ref_matrix_set(void *p, uint32_t in, uint32_t out, const float *ir, uintptr_t len)
{
    auto *m = static_cast<RefMatrix *>(p);
    size_t idx = m->parallel ? out : size_t(out) * m->nIn + in;
    int err = m->pairs[idx]->set(ir, len, true);
    m->pairs[idx]->setResetOffset(0);
    return err;
}
// MonoConvolve::reset of one pair (what Convolver::reset(in, out) forwards to, Convolver.cpp:88-97)
SHIM int ref_matrix_reset_pair(void *p, uint32_t in, uint32_t out)
{
    auto *m = static_cast<RefMatrix *>(p);
    size_t idx = m->parallel ? out : size_t(out) * m->nIn + in;
    int err = m->pairs[idx]->reset();
    m->pairs[idx]->setResetOffset(0);
    return err;
}

// Load every pair of the matrix from a pool of `npool` impulse responses of `len` taps (pair (in, out) takes pool entry
// (in + out) % npool), the rows dealt to `threads` host threads: set-up of the full-size timing runs, where 4096 pairs
// set one after the other take longer than the timed steps.
SHIM int ref_matrix_set_pool_mt(void *p, const float *const *pool_irs, uint32_t npool, uintptr_t len, int threads)
{
    auto *m = static_cast<RefMatrix *>(p);
    if (threads < 1) threads = 1;
    if (uint32_t(threads) > m->nOut) threads = int(m->nOut);
    std::vector<int> errs(threads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        uint32_t r0 = uint32_t(uint64_t(m->nOut) * t / threads), r1 = uint32_t(uint64_t(m->nOut) * (t + 1) / threads);
        pool.emplace_back([=, &errs]()
        {
            for (uint32_t o = r0; o < r1; o++)
                for (uint32_t i = 0; i < (m->parallel ? 1u : m->nIn); i++)
                {
                    size_t idx = m->parallel ? o : size_t(o) * m->nIn + i;
                    int e = m->pairs[idx]->set(pool_irs[(i + o) % npool], len, true);
                    m->pairs[idx]->setResetOffset(0);
                    if (e) errs[t] = e;
                }
        });
    }
    for (auto &th : pool) th.join();
    for (int e : errs) if (e) return e;
    return 0;
}

// ins: nIn planar pointers, outs: nOut planar pointers; rows [row0,row1) only (for threading)
static void matrix_rows(RefMatrix *m, const float *const *ins, float *const *outs, size_t n, uint32_t row0, uint32_t row1, float *temp)
{
    for (uint32_t o = row0; o < row1; o++)
    {
        std::fill_n(outs[o], n, 0.f);
        if (m->parallel)
            m->pairs[o]->process(ins[o], temp, outs[o], n, true);
        else
            for (uint32_t i = 0; i < m->nIn; i++)
                m->pairs[size_t(o) * m->nIn + i]->process(ins[i], temp, outs[o], n, true);
    }
}
SHIM void ref_matrix_process(void *p, const float *const *ins, float *const *outs, size_t n)
{
    auto *m = static_cast<RefMatrix *>(p);
    m->temp.resize(n);
#if defined(__SSE__)
    unsigned int old = _mm_getcsr(); _mm_setcsr(old | 0x8040);   // as Convolver.cpp:197-203
#endif
    matrix_rows(m, ins, outs, n, 0, m->nOut, m->temp.data());
#if defined(__SSE__)
    _mm_setcsr(old);
#endif
}

// The same call with the output rows dealt to `threads` host threads (rows are independent objects, inputs are read-only):
// used by the full-size parity checks, where one thread would take minutes.  Results do not depend on the thread count.
SHIM void ref_matrix_process_mt(void *p, const float *const *ins, float *const *outs, size_t n, int threads)
{
    auto *m = static_cast<RefMatrix *>(p);
    if (threads < 1) threads = 1;
    if (uint32_t(threads) > m->nOut) threads = int(m->nOut);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
    {
        uint32_t r0 = uint32_t(uint64_t(m->nOut) * t / threads), r1 = uint32_t(uint64_t(m->nOut) * (t + 1) / threads);
        pool.emplace_back([=]()
        {
#if defined(__SSE__)
            _mm_setcsr(_mm_getcsr() | 0x8040);
#endif
            std::vector<float> temp(n);
            matrix_rows(m, ins, outs, n, r0, r1, temp.data());
        });
    }
    for (auto &th : pool) th.join();
}

// Time `hops` calls of `block` samples on all rows with `threads` host threads (one thread per
// contiguous band of output rows; the objects are independent, inputs are read-only).
// Returns seconds of wall time for the timed hops (after `warm` untimed hops).
SHIM double ref_matrix_time(void *p, const float *const *ins, float *const *outs, size_t block, int warm, int hops, int threads)
{
    auto *m = static_cast<RefMatrix *>(p);
    if (threads < 1) threads = 1;
    if (uint32_t(threads) > m->nOut) threads = int(m->nOut);
    std::vector<std::vector<float>> temps(threads, std::vector<float>(block));
    auto run = [&](int count)
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
        {
            uint32_t r0 = uint32_t(uint64_t(m->nOut) * t / threads), r1 = uint32_t(uint64_t(m->nOut) * (t + 1) / threads);
            pool.emplace_back([=, &temps]()
            {
#if defined(__SSE__)
                _mm_setcsr(_mm_getcsr() | 0x8040);
#endif
                for (int h = 0; h < count; h++)
                    matrix_rows(m, ins, outs, block, r0, r1, temps[t].data());
            });
        }
        for (auto &th : pool) th.join();
    };
    run(warm);
    auto t0 = std::chrono::steady_clock::now();
    run(hops);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

SHIM int ref_hardware_threads() { return int(std::thread::hardware_concurrency()); }

// ---------------------------------------------------------------------------------------------
// PConvRestated<T>: the partitioned-convolution state machine of PartitionedConvolve.cpp:173-426
// written over the reference FFT for any T ("Double (C5)").  Frames are kept as a
// simple 2-hop history rather than the reference's two phase-shifted buffers; the arithmetic fed
// to the FFT, the cMAC order of PartitionedConvolve.cpp:412-413 (separate mul/add, ring order,
// newest frame x partition 0 last) and the 1/(4N) scale of :237 are the reference's.
// ---------------------------------------------------------------------------------------------

template <class T> struct FFTSel;
template <> struct FFTSel<float>  { typedef FFT_SETUP_F Setup; typedef FFT_SPLIT_COMPLEX_F Split; };
template <> struct FFTSel<double> { typedef FFT_SETUP_D Setup; typedef FFT_SPLIT_COMPLEX_D Split; };

template <class T>
struct PConvRestated
{
    typedef typename FFTSel<T>::Setup Setup;
    typedef typename FFTSel<T>::Split Split;

    uintptr_t log2n, N, B, P = 0, valid = 0, head = 0, phase = 0;
    Setup setup;
    std::vector<T> hr, hi, xr, xi, newest, previous, frame, pending, ar, ai, y;

    explicit PConvRestated(uintptr_t fftSize)
    {
        log2n = 0; while ((uintptr_t(1) << log2n) < fftSize) log2n++;
        N = uintptr_t(1) << log2n; B = N >> 1;
        hisstools_create_setup(&setup, log2n);
        newest.assign(B, 0); previous.assign(B, 0); frame.assign(N, 0); pending.assign(B, 0);
        ar.assign(B, 0); ai.assign(B, 0); y.assign(N, 0);
    }
    ~PConvRestated() { hisstools_destroy_setup(setup); }

    void set(const T *ir, uintptr_t len)           // PartitionedConvolve.cpp:203-219
    {
        P = (len + B - 1) / B;
        hr.assign(P * B, 0); hi.assign(P * B, 0); xr.assign(P * B, 0); xi.assign(P * B, 0);
        std::vector<T> tmp(N);
        for (uintptr_t p = 0; p < P; p++)
        {
            uintptr_t count = std::min(B, len - p * B);
            std::copy_n(ir + p * B, count, tmp.begin());
            std::fill(tmp.begin() + count, tmp.end(), T(0));
            Split s(hr.data() + p * B, hi.data() + p * B);
            hisstools_rfft(setup, tmp.data(), &s, N, log2n);
        }
        reset();
    }
    void reset()                                    // PartitionedConvolve.cpp:267-290 with resetOffset 0
    {
        std::fill(newest.begin(), newest.end(), T(0)); std::fill(previous.begin(), previous.end(), T(0));
        std::fill(pending.begin(), pending.end(), T(0));
        valid = 0; head = 0; phase = 0;
    }
    void mac(const T *x_r, const T *x_i, const T *h_r, const T *h_i)   // PartitionedConvolve.cpp:387-426
    {
        ai[0] += x_i[0] * h_i[0];
        ar[0] += x_r[0] * h_r[0];
        for (uintptr_t k = 1; k < B; k++)
        {
            ar[k] += (x_r[k] * h_r[k]) - (x_i[k] * h_i[k]);
            ai[k] += (x_r[k] * h_i[k]) + (x_i[k] * h_r[k]);
        }
    }
    void hop()                                      // PartitionedConvolve.cpp:326-376
    {
        // older frames first, in ring order starting from the most recent old frame
        for (uintptr_t p = 1; p < valid + 1 && p < P; p++)
        {
            uintptr_t slot = (head + p) % P;
            mac(xr.data() + slot * B, xi.data() + slot * B, hr.data() + p * B, hi.data() + p * B);
        }
        std::copy(newest.begin(), newest.end(), frame.begin());
        std::copy(previous.begin(), previous.end(), frame.begin() + B);
        Split xs(xr.data() + head * B, xi.data() + head * B);
        hisstools_rfft(setup, frame.data(), &xs, N, log2n);
        mac(xs.realp, xs.imagp, hr.data(), hi.data());
        Split as(ar.data(), ai.data());
        hisstools_rifft(setup, &as, y.data(), log2n);
        const T scale = T(1) / static_cast<T>(N << 2);
        for (uintptr_t k = 0; k < B; k++) pending[k] = y[k] * scale;
        std::fill(ar.begin(), ar.end(), T(0)); std::fill(ai.begin(), ai.end(), T(0));
        previous = newest;
        head = head ? head - 1 : P - 1;
        valid = std::min(P - 1, valid + 1);
    }
    bool process(const T *in, T *out, uintptr_t n)
    {
        if (!P) return false;
        for (uintptr_t s = 0; s < n; s++)
        {
            newest[phase] = in[s];
            out[s] = pending[phase];
            if (++phase == B) { phase = 0; hop(); }
        }
        return true;
    }
};

SHIM void *ref_restated_create_f32(uintptr_t fftSize) { return new PConvRestated<float>(fftSize); }
SHIM void *ref_restated_create_f64(uintptr_t fftSize) { return new PConvRestated<double>(fftSize); }
SHIM void ref_restated_destroy_f32(void *p) { delete static_cast<PConvRestated<float> *>(p); }
SHIM void ref_restated_destroy_f64(void *p) { delete static_cast<PConvRestated<double> *>(p); }
SHIM void ref_restated_set_f32(void *p, const float *ir, uintptr_t len) { static_cast<PConvRestated<float> *>(p)->set(ir, len); }
SHIM void ref_restated_set_f64(void *p, const double *ir, uintptr_t len) { static_cast<PConvRestated<double> *>(p)->set(ir, len); }
SHIM int ref_restated_process_f32(void *p, const float *in, float *out, uintptr_t n) { return static_cast<PConvRestated<float> *>(p)->process(in, out, n); }
SHIM int ref_restated_process_f64(void *p, const double *in, double *out, uintptr_t n) { return static_cast<PConvRestated<double> *>(p)->process(in, out, n); }

// Time `hops` blocks of `block` samples over `chans` independent double channels with `threads` threads.
SHIM double ref_restated_time_f64(void *const *objs, int chans, const double *const *ins, double *const *outs, size_t block, int warm, int hops, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > chans) threads = chans;
    auto run = [&](int count)
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
        {
            int c0 = int(int64_t(chans) * t / threads), c1 = int(int64_t(chans) * (t + 1) / threads);
            pool.emplace_back([=]()
            {
                for (int h = 0; h < count; h++)
                    for (int c = c0; c < c1; c++)
                        static_cast<PConvRestated<double> *>(objs[c])->process(ins[c], outs[c], block);
            });
        }
        for (auto &th : pool) th.join();
    };
    run(warm);
    auto t0 = std::chrono::steady_clock::now();
    run(hops);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}
