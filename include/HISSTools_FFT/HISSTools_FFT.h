// HISSTools_FFT.h -- B200 drop-in for HISSTools_FFT/HISSTools_FFT.h of HISSTools_Library.
//
// Same overload set, argument order and numerical conventions as the reference header
// (HISSTools_FFT/HISSTools_FFT.h:87-369): split-complex planes, forward kernel exp(-j theta), real
// forward transform = 2*DFT with DC in realp[0] and Nyquist in imagp[0], nothing scaled.  Every
// transform is one call into libhisstools_b200.so (include/hisstools_b200.h), which runs it as a
// shared-memory Stockham FFT kernel on the GPU; zip / unzip are pure re-orderings of the caller's
// memory (reference Core:1185-1287) and stay inline on the host.  Link with -lhisstools_b200.
// A failed CUDA call throws std::runtime_error carrying hb_last_error(): there is no CPU fallback.
#ifndef HISSTOOLS_B200_FFT_HPP
#define HISSTOOLS_B200_FFT_HPP

#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <string>

#include "../hisstools_b200.h"

template <class T> struct Split
{
    Split() : realp(nullptr), imagp(nullptr) {}
    Split(T *real, T *imag) : realp(real), imagp(imag) {}
    T *realp;
    T *imagp;
};

typedef Split<double> DoubleSplit;
typedef Split<float> FloatSplit;
typedef DoubleSplit FFT_SPLIT_COMPLEX_D;
typedef FloatSplit FFT_SPLIT_COMPLEX_F;

// opaque setups (reference: HISSTools_FFT.h:57,63); each wraps one hb_fft_setup
struct DoubleSetup { hb_fft_setup *h; };
struct FloatSetup { hb_fft_setup *h; };
typedef struct DoubleSetup *FFT_SETUP_D;
typedef struct FloatSetup *FFT_SETUP_F;

namespace hisstools_b200_detail
{
    inline void check(int code)
    {
        if (code < 0) throw std::runtime_error(std::string("hisstools_b200: ") + hb_last_error());
    }
    inline int device()
    {
        return 0;
    }
    template <class T, class U> void unzip(const U *input, Split<T> *output, uintptr_t half)
    {
        for (uintptr_t i = 0; i < half; i++)
        {
            output->realp[i] = static_cast<T>(input[2 * i]);
            output->imagp[i] = static_cast<T>(input[2 * i + 1]);
        }
    }
    template <class T, class U> void unzip_zero(const U *input, Split<T> *output, uintptr_t in_length, uintptr_t log2n)
    {
        // reference Core:1258-1287: clamp to the FFT size, an odd last sample lands in realp, zero the rest
        const uintptr_t n = uintptr_t(1) << log2n, half = n >> 1;
        const uintptr_t length = std::min(in_length, n), pairs = length >> 1;
        unzip(input, output, pairs);
        if (length & 1)
        {
            output->realp[pairs] = static_cast<T>(input[length - 1]);
            output->imagp[pairs] = T(0);
        }
        for (uintptr_t i = pairs + (length & 1); i < half; i++) output->realp[i] = output->imagp[i] = T(0);
    }
}

// setup (HISSTools_FFT.h:87,98,108,118)
inline void hisstools_create_setup(FFT_SETUP_D *setup, uintptr_t max_fft_log_2)
{
    *setup = new DoubleSetup{nullptr};
    hisstools_b200_detail::check(hb_fft_setup_create(&(*setup)->h, HB_F64, max_fft_log_2, hisstools_b200_detail::device()));
}
inline void hisstools_create_setup(FFT_SETUP_F *setup, uintptr_t max_fft_log_2)
{
    *setup = new FloatSetup{nullptr};
    hisstools_b200_detail::check(hb_fft_setup_create(&(*setup)->h, HB_F32, max_fft_log_2, hisstools_b200_detail::device()));
}
inline void hisstools_destroy_setup(FFT_SETUP_D setup) { if (setup) { hb_fft_setup_destroy(setup->h); delete setup; } }
inline void hisstools_destroy_setup(FFT_SETUP_F setup) { if (setup) { hb_fft_setup_destroy(setup->h); delete setup; } }

// in-place complex and real transforms (HISSTools_FFT.h:130-166, 220-256)
inline void hisstools_fft(FFT_SETUP_D setup, FFT_SPLIT_COMPLEX_D *input, uintptr_t log2n) { hisstools_b200_detail::check(hb_fft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_fft(FFT_SETUP_F setup, FFT_SPLIT_COMPLEX_F *input, uintptr_t log2n) { hisstools_b200_detail::check(hb_fft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_ifft(FFT_SETUP_D setup, FFT_SPLIT_COMPLEX_D *input, uintptr_t log2n) { hisstools_b200_detail::check(hb_ifft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_ifft(FFT_SETUP_F setup, FFT_SPLIT_COMPLEX_F *input, uintptr_t log2n) { hisstools_b200_detail::check(hb_ifft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_rfft(FFT_SETUP_D setup, FFT_SPLIT_COMPLEX_D *input, uintptr_t log2n) { if (log2n) hisstools_b200_detail::check(hb_rfft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_rfft(FFT_SETUP_F setup, FFT_SPLIT_COMPLEX_F *input, uintptr_t log2n) { if (log2n) hisstools_b200_detail::check(hb_rfft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_rifft(FFT_SETUP_D setup, FFT_SPLIT_COMPLEX_D *input, uintptr_t log2n) { if (log2n) hisstools_b200_detail::check(hb_rifft(setup->h, input->realp, input->imagp, log2n)); }
inline void hisstools_rifft(FFT_SETUP_F setup, FFT_SPLIT_COMPLEX_F *input, uintptr_t log2n) { if (log2n) hisstools_b200_detail::check(hb_rifft(setup->h, input->realp, input->imagp, log2n)); }

// out-of-place real transforms with zero padding / zipping (HISSTools_FFT.h:180-208, 269-282)
inline void hisstools_rfft(FFT_SETUP_D setup, const double *input, FFT_SPLIT_COMPLEX_D *output, uintptr_t in_length, uintptr_t log2n)
{ hisstools_b200_detail::check(hb_rfft_real(setup->h, input, HB_F64, output->realp, output->imagp, in_length, log2n)); }
inline void hisstools_rfft(FFT_SETUP_F setup, const float *input, FFT_SPLIT_COMPLEX_F *output, uintptr_t in_length, uintptr_t log2n)
{ hisstools_b200_detail::check(hb_rfft_real(setup->h, input, HB_F32, output->realp, output->imagp, in_length, log2n)); }
inline void hisstools_rfft(FFT_SETUP_D setup, const float *input, FFT_SPLIT_COMPLEX_D *output, uintptr_t in_length, uintptr_t log2n)
{ hisstools_b200_detail::check(hb_rfft_real(setup->h, input, HB_F32, output->realp, output->imagp, in_length, log2n)); }
inline void hisstools_rifft(FFT_SETUP_D setup, FFT_SPLIT_COMPLEX_D *input, double *output, uintptr_t log2n)
{ hisstools_b200_detail::check(hb_rifft_real(setup->h, input->realp, input->imagp, output, log2n)); }
inline void hisstools_rifft(FFT_SETUP_F setup, FFT_SPLIT_COMPLEX_F *input, float *output, uintptr_t log2n)
{ hisstools_b200_detail::check(hb_rifft_real(setup->h, input->realp, input->imagp, output, log2n)); }

// zip / unzip (HISSTools_FFT.h:295-369)
inline void hisstools_unzip_zero(const double *input, FFT_SPLIT_COMPLEX_D *output, uintptr_t in_length, uintptr_t log2n) { hisstools_b200_detail::unzip_zero(input, output, in_length, log2n); }
inline void hisstools_unzip_zero(const float *input, FFT_SPLIT_COMPLEX_F *output, uintptr_t in_length, uintptr_t log2n) { hisstools_b200_detail::unzip_zero(input, output, in_length, log2n); }
inline void hisstools_unzip_zero(const float *input, FFT_SPLIT_COMPLEX_D *output, uintptr_t in_length, uintptr_t log2n) { hisstools_b200_detail::unzip_zero(input, output, in_length, log2n); }
inline void hisstools_unzip(const double *input, FFT_SPLIT_COMPLEX_D *output, uintptr_t log2n) { hisstools_b200_detail::unzip(input, output, (uintptr_t(1) << log2n) >> 1); }
inline void hisstools_unzip(const float *input, FFT_SPLIT_COMPLEX_F *output, uintptr_t log2n) { hisstools_b200_detail::unzip(input, output, (uintptr_t(1) << log2n) >> 1); }
inline void hisstools_zip(const FFT_SPLIT_COMPLEX_D *input, double *output, uintptr_t log2n)
{
    for (uintptr_t i = 0, half = (uintptr_t(1) << log2n) >> 1; i < half; i++) { output[2 * i] = input->realp[i]; output[2 * i + 1] = input->imagp[i]; }
}
inline void hisstools_zip(const FFT_SPLIT_COMPLEX_F *input, float *output, uintptr_t log2n)
{
    for (uintptr_t i = 0, half = (uintptr_t(1) << log2n) >> 1; i < half; i++) { output[2 * i] = input->realp[i]; output[2 * i + 1] = input->imagp[i]; }
}

#endif
