// PartitionedConvolve.h -- B200 drop-in for HISSTools::PartitionedConvolve
// (reference: HIRT_Multichannel_Convolution/PartitionedConvolve.h:23-41, .cpp:52-385).
// Same constructor, setters, error codes and process() contract (output = linear convolution delayed
// by FFTSize/2; returns false and leaves `out` untouched when no IR is loaded); the work runs on the
// GPU through hb_conv_* of include/hisstools_b200.h.  Superset: a double-precision engine
// (PartitionedConvolveD) and a device-pointer process.
#pragma once

#include "../HISSTools_FFT/HISSTools_FFT.h"
#include "ConvolveErrors.h"

#include <cstdint>

namespace HISSTools
{
    template <class T, int DTYPE>
    class PartitionedConvolveT
    {
    public:

        PartitionedConvolveT(uintptr_t maxFFTSize, uintptr_t maxLength, uintptr_t offset, uintptr_t length, int device = 0) : mHandle(nullptr)
        {
            hisstools_b200_detail::check(hb_conv_create(&mHandle, DTYPE, 1, 1, 1, maxFFTSize, maxLength, offset, length, device));
        }
        ~PartitionedConvolveT() { hb_conv_destroy(mHandle); }

        // Non-moveable and copyable, as the reference (PartitionedConvolve.h:29-32)
        PartitionedConvolveT(PartitionedConvolveT& obj) = delete;
        PartitionedConvolveT& operator = (PartitionedConvolveT& obj) = delete;
        PartitionedConvolveT(PartitionedConvolveT&& obj) = delete;
        PartitionedConvolveT& operator = (PartitionedConvolveT&& obj) = delete;

        ConvolveError setFFTSize(uintptr_t FFTSize) { return b200::to_error(hb_conv_set_fft_size(mHandle, FFTSize)); }
        ConvolveError setLength(uintptr_t length) { return b200::to_error(hb_conv_set_length(mHandle, length)); }
        void setOffset(uintptr_t offset) { hb_conv_set_offset(mHandle, offset); }
        void setResetOffset(intptr_t offset = -1) { hb_conv_set_reset_offset(mHandle, offset); }

        ConvolveError set(const T *input, uintptr_t length) { return b200::to_error(hb_conv_set_ir(mHandle, 0, 0, 0, input, DTYPE, length)); }
        void reset() { hb_conv_reset(mHandle); }

        bool process(const T *in, T *out, uintptr_t numSamples)
        {
            const void *ins[1] = { in };
            void *outs[1] = { out };
            const int code = hb_conv_process(mHandle, ins, outs, numSamples, 0);
            if (code == HB_ERR_NO_IR) return false;
            hisstools_b200_detail::check(code);
            return true;
        }

        // device-resident variant (superset): enqueues on `stream`, no synchronisation
        bool processDevice(const T *d_in, T *d_out, uintptr_t numSamples, void *stream = nullptr)
        {
            const int code = hb_conv_process_dev(mHandle, d_in, numSamples, d_out, numSamples, numSamples, 0, stream);
            if (code == HB_ERR_NO_IR) return false;
            hisstools_b200_detail::check(code);
            return true;
        }

        // may consecutive processDevice calls of one block run side by side on the GPU?  0 never, 1 (default) on the object's own
        // stream (stream = nullptr), 2 on any stream -- the caller's rows are complete when a call is made (hb_conv_set_hop_overlap)
        void setHopOverlap(int mode) { hisstools_b200_detail::check(hb_conv_set_hop_overlap(mHandle, mode)); }

        hb_conv *handle() { return mHandle; }

    private:

        hb_conv *mHandle;
    };

    typedef PartitionedConvolveT<float, HB_F32> PartitionedConvolve;
    typedef PartitionedConvolveT<double, HB_F64> PartitionedConvolveD;
}
