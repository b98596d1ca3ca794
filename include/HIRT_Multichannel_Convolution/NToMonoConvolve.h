// NToMonoConvolve.h -- B200 drop-in for HISSTools::NToMonoConvolve
// (reference: HIRT_Multichannel_Convolution/NToMonoConvolve.h:18-24, .cpp:4-43): N inputs convolved
// and summed into one output.  The reference sums N MonoConvolve outputs in the time domain; here one
// hb_matrix (N x 1) transforms each input once and sums over inputs in the frequency domain.
// Superset: the custom partition sizes MonoConvolve accepts.
#pragma once

#include "MonoConvolve.h"
#include "ConvolveErrors.h"

#include <algorithm>
#include <cstdint>
#include <vector>

namespace HISSTools
{
    class NToMonoConvolve
    {
    public:

        NToMonoConvolve(uint32_t input_chans, uintptr_t maxLength, LatencyMode latency)
        : mMatrix(1, input_chans, 1, maxLength, latency), mNumInChans(input_chans) {}
        NToMonoConvolve(uint32_t input_chans, uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B = 0, uint32_t C = 0, uint32_t D = 0)
        : mMatrix(1, input_chans, 1, maxLength, zeroLatency, A, B, C, D), mNumInChans(input_chans) {}

        ConvolveError resize(uint32_t inChan, uintptr_t impulse_length)
        {
            if (inChan >= mNumInChans) return CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE;       // NToMonoConvolve.cpp:11-18
            return b200::to_error(hb_matrix_resize(mMatrix.handle(), 0, inChan, 0, impulse_length));
        }
        ConvolveError set(uint32_t inChan, const float *input, uintptr_t impulse_length, bool resize)
        {
            if (inChan >= mNumInChans) return CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE;
            return b200::to_error(hb_matrix_set(mMatrix.handle(), 0, inChan, 0, input, HB_F32, impulse_length, resize ? 1 : 0));
        }
        ConvolveError reset(uint32_t inChan)
        {
            if (inChan >= mNumInChans) return CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE;
            return b200::to_error(hb_matrix_reset_pair(mMatrix.handle(), 0, inChan, 0));     // that input's convolver only (NToMonoConvolve.cpp:28-33)
        }
        void setResetOffset(intptr_t offset = -1) { hb_matrix_set_reset_offset(mMatrix.handle(), offset); }

        // out = sum over the first active_in_chans inputs (NToMonoConvolve.cpp:35-43); `temp` unused
        void process(const float * const* ins, float *out, float *temp, size_t numSamples, size_t active_in_chans)
        {
            (void) temp;
            std::fill_n(out, numSamples, 0.f);
            mRows.assign(mNumInChans, nullptr);
            for (size_t i = 0; i < std::min<size_t>(active_in_chans, mNumInChans); i++) mRows[i] = ins[i];
            void *outs[1] = { out };
            const int code = hb_matrix_process(mMatrix.handle(), mRows.data(), outs, numSamples, 1);
            if (code != HB_ERR_NO_IR && code != HB_ERR_BUSY) hisstools_b200_detail::check(code);
        }

    private:

        b200::Matrix mMatrix;
        std::vector<const void *> mRows;
        uint32_t mNumInChans;
    };
}
