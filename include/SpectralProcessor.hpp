// SpectralProcessor.hpp -- B200 drop-in for the convolution / correlation part of the reference's
// spectral_processor<T> (SpectralProcessor.hpp:11-683): convolve and correlate for real inputs (T*, in_ptr, in_ptr,
// EdgeMode) and complex inputs (T*, T*, in_ptr x 4, EdgeMode), convolved_size / correlated_size, set_max_fft_size /
// max_fft_size.  The transforms, the per-bin products (SpectralFunctions.hpp:49-84, 265-281) and the edge-mode
// arrangements (:445-538) run on the GPU through hb_spectral_* of hisstools_b200.h, as does change_phase (:186-208).
// The raw fft / rfft members are not provided (use HISSTools_FFT.h).
#ifndef HISSTOOLS_B200_SPECTRALPROCESSOR_HPP
#define HISSTOOLS_B200_SPECTRALPROCESSOR_HPP

#include <cstdint>
#include <type_traits>

#include "HISSTools_FFT/HISSTools_FFT.h"

template <typename T>
class spectral_processor
{
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "float or double");

public:

    enum class EdgeMode { Linear, Wrap, WrapCentre, Fold, FoldRepeat };

    struct in_ptr
    {
        in_ptr(const T* ptr, uintptr_t size) : m_ptr(ptr), m_size(size) {}

        const T* m_ptr;
        const uintptr_t m_size;
    };

    spectral_processor(uintptr_t max_fft_size = 32768, int device = 0) : m_handle(nullptr)
    {
        hisstools_b200_detail::check(hb_spectral_create(&m_handle, std::is_same<T, double>::value ? HB_F64 : HB_F32, max_fft_size, device));
    }
    ~spectral_processor() { hb_spectral_destroy(m_handle); }

    spectral_processor(const spectral_processor&) = delete;
    spectral_processor &operator =(const spectral_processor&) = delete;
    spectral_processor(spectral_processor&& b) : m_handle(b.m_handle) { b.m_handle = nullptr; }
    spectral_processor &operator =(spectral_processor&& b)
    {
        if (this != &b) { hb_spectral_destroy(m_handle); m_handle = b.m_handle; b.m_handle = nullptr; }
        return *this;
    }

    void set_max_fft_size(uintptr_t size) { hisstools_b200_detail::check(hb_spectral_set_max_fft_size(m_handle, size)); }
    uintptr_t max_fft_size() const { return hb_spectral_max_fft_size(m_handle); }

    void convolve(T *output, in_ptr in1, in_ptr in2, EdgeMode mode)
    {
        hisstools_b200_detail::check(hb_spectral_convolve(m_handle, output, in1.m_ptr, in1.m_size, in2.m_ptr, in2.m_size, static_cast<int>(mode), nullptr));
    }

    void convolve(T *r_out, T *i_out, in_ptr r_in1, in_ptr i_in1, in_ptr r_in2, in_ptr i_in2, EdgeMode mode)
    {
        hisstools_b200_detail::check(hb_spectral_convolve_complex(m_handle, r_out, i_out, r_in1.m_ptr, r_in1.m_size, i_in1.m_ptr, i_in1.m_size,
                                                                  r_in2.m_ptr, r_in2.m_size, i_in2.m_ptr, i_in2.m_size, static_cast<int>(mode), nullptr));
    }

    void correlate(T *output, in_ptr in1, in_ptr in2, EdgeMode mode)
    {
        hisstools_b200_detail::check(hb_spectral_correlate(m_handle, output, in1.m_ptr, in1.m_size, in2.m_ptr, in2.m_size, static_cast<int>(mode), nullptr));
    }

    void correlate(T *r_out, T *i_out, in_ptr r_in1, in_ptr i_in1, in_ptr r_in2, in_ptr i_in2, EdgeMode mode)
    {
        hisstools_b200_detail::check(hb_spectral_correlate_complex(m_handle, r_out, i_out, r_in1.m_ptr, r_in1.m_size, i_in1.m_ptr, i_in1.m_size,
                                                                   r_in2.m_ptr, r_in2.m_size, i_in2.m_ptr, i_in2.m_size, static_cast<int>(mode), nullptr));
    }

    uintptr_t convolved_size(uintptr_t size1, uintptr_t size2, EdgeMode mode) const
    {
        return hb_spectral_convolved_size(m_handle, size1, size2, static_cast<int>(mode));
    }

    uintptr_t correlated_size(uintptr_t size1, uintptr_t size2, EdgeMode mode) const { return convolved_size(size1, size2, mode); }

    void change_phase(T *output, const T *input, uintptr_t size, double phase, double time_multiplier = 1.0)
    {
        hisstools_b200_detail::check(hb_spectral_change_phase(m_handle, output, input, size, phase, time_multiplier, nullptr));
    }

private:

    hb_spectral *m_handle;
};

#endif
