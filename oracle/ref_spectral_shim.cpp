// oracle/ref_spectral_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" window onto the reference's one-shot spectral convolution
// (SpectralProcessor.hpp:169-172 `spectral_processor<T>::convolve(T*, in_ptr, in_ptr, EdgeMode)`),
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
