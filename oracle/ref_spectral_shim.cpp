// oracle/ref_spectral_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" window onto the reference's one-shot spectral convolution and correlation
// (SpectralProcessor.hpp:164-184 `spectral_processor<T>::convolve / correlate`, real and complex inputs),
// compiled in place from $REF_ROOT by oracle/Makefile into oracle/_ref/libhisstools_ref_spectral.so.
// Built with -mavx because spectral_processor<float> does not instantiate on an SSE2-only build
// (SpectralFunctions.hpp:52-58 needs SIMDType<float,4> as the half-width type).

#include "SpectralProcessor.hpp"

#include <cstdint>

#define SHIM extern "C" __attribute__((visibility("default")))

template <class T>
static uintptr_t spectral_convolve(T *out, const T *in1, uintptr_t n1, const T *in2, uintptr_t n2, int mode, uintptr_t maxFFT)
{
    typedef spectral_processor<T> Proc;
    Proc proc(maxFFT);
    typename Proc::EdgeMode m = static_cast<typename Proc::EdgeMode>(mode);
    uintptr_t size = proc.convolved_size(n1, n2, m);
    proc.convolve(out, typename Proc::in_ptr(in1, n1), typename Proc::in_ptr(in2, n2), m);
    return size;
}

// mode: 0 Linear, 1 Wrap, 2 WrapCentre, 3 Fold, 4 FoldRepeat (SpectralProcessor.hpp:22).
// Returns convolved_size() (0 means the reference silently did nothing: SpectralProcessor.hpp:651-652).
SHIM uintptr_t ref_spectral_convolve_f32(float *out, const float *in1, uintptr_t n1, const float *in2, uintptr_t n2, int mode, uintptr_t maxFFT)
{ return spectral_convolve<float>(out, in1, n1, in2, n2, mode, maxFFT); }
SHIM uintptr_t ref_spectral_convolve_f64(double *out, const double *in1, uintptr_t n1, const double *in2, uintptr_t n2, int mode, uintptr_t maxFFT)
{ return spectral_convolve<double>(out, in1, n1, in2, n2, mode, maxFFT); }

// op: 0 convolve, 1 correlate; real inputs (SpectralProcessor.hpp:169-172, 181-184)
template <class T>
static uintptr_t spectral_binary(T *out, const T *in1, uintptr_t n1, const T *in2, uintptr_t n2, int mode, int op, uintptr_t maxFFT)
{
    typedef spectral_processor<T> Proc;
    Proc proc(maxFFT);
    typename Proc::EdgeMode m = static_cast<typename Proc::EdgeMode>(mode);
    uintptr_t size = op ? proc.correlated_size(n1, n2, m) : proc.convolved_size(n1, n2, m);
    if (op) proc.correlate(out, typename Proc::in_ptr(in1, n1), typename Proc::in_ptr(in2, n2), m);
    else proc.convolve(out, typename Proc::in_ptr(in1, n1), typename Proc::in_ptr(in2, n2), m);
    return size;
}

// complex inputs (SpectralProcessor.hpp:164-167, 176-179); a plane of length 0 is absent
template <class T>
static uintptr_t spectral_binary_complex(T *r_out, T *i_out, const T *r1, uintptr_t nr1, const T *i1, uintptr_t ni1,
                                         const T *r2, uintptr_t nr2, const T *i2, uintptr_t ni2, int mode, int op, uintptr_t maxFFT)
{
    typedef spectral_processor<T> Proc;
    typedef typename Proc::in_ptr in_ptr;
    Proc proc(maxFFT);
    typename Proc::EdgeMode m = static_cast<typename Proc::EdgeMode>(mode);
    uintptr_t n1 = nr1 > ni1 ? nr1 : ni1, n2 = nr2 > ni2 ? nr2 : ni2;
    uintptr_t size = proc.convolved_size(n1, n2, m);
    if (op) proc.correlate(r_out, i_out, in_ptr(r1, nr1), in_ptr(i1, ni1), in_ptr(r2, nr2), in_ptr(i2, ni2), m);
    else proc.convolve(r_out, i_out, in_ptr(r1, nr1), in_ptr(i1, ni1), in_ptr(r2, nr2), in_ptr(i2, ni2), m);
    return size;
}

SHIM uintptr_t ref_spectral_binary_f32(float *out, const float *in1, uintptr_t n1, const float *in2, uintptr_t n2, int mode, int op, uintptr_t maxFFT)
{ return spectral_binary<float>(out, in1, n1, in2, n2, mode, op, maxFFT); }
SHIM uintptr_t ref_spectral_binary_f64(double *out, const double *in1, uintptr_t n1, const double *in2, uintptr_t n2, int mode, int op, uintptr_t maxFFT)
{ return spectral_binary<double>(out, in1, n1, in2, n2, mode, op, maxFFT); }
SHIM uintptr_t ref_spectral_binary_complex_f32(float *r_out, float *i_out, const float *r1, uintptr_t nr1, const float *i1, uintptr_t ni1,
                                               const float *r2, uintptr_t nr2, const float *i2, uintptr_t ni2, int mode, int op, uintptr_t maxFFT)
{ return spectral_binary_complex<float>(r_out, i_out, r1, nr1, i1, ni1, r2, nr2, i2, ni2, mode, op, maxFFT); }
SHIM uintptr_t ref_spectral_binary_complex_f64(double *r_out, double *i_out, const double *r1, uintptr_t nr1, const double *i1, uintptr_t ni1,
                                               const double *r2, uintptr_t nr2, const double *i2, uintptr_t ni2, int mode, int op, uintptr_t maxFFT)
{ return spectral_binary_complex<double>(r_out, i_out, r1, nr1, i1, ni1, r2, nr2, i2, ni2, mode, op, maxFFT); }

// change_phase (SpectralProcessor.hpp:186-208); out must hold the FFT size, which is returned
template <class T>
static uintptr_t spectral_change_phase(T *out, const T *in, uintptr_t size, double phase, double time_multiplier, uintptr_t maxFFT)
{
    typedef spectral_processor<T> Proc;
    Proc proc(maxFFT);
    proc.change_phase(out, in, size, phase, time_multiplier);
    if (size <= 1) return size;
    return uintptr_t(1) << Proc::calc_fft_size_log2((uintptr_t) std::round(size * time_multiplier));
}
SHIM uintptr_t ref_spectral_change_phase_f32(float *out, const float *in, uintptr_t size, double phase, double tm, uintptr_t maxFFT)
{ return spectral_change_phase<float>(out, in, size, phase, tm, maxFFT); }
SHIM uintptr_t ref_spectral_change_phase_f64(double *out, const double *in, uintptr_t size, double phase, double tm, uintptr_t maxFFT)
{ return spectral_change_phase<double>(out, in, size, phase, tm, maxFFT); }
