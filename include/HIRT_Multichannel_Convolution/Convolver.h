// Convolver.h -- B200 drop-in for HISSTools::Convolver
// (reference: HIRT_Multichannel_Convolution/Convolver.h:25-50, .cpp:5-195): an N x M matrix convolver or
// numIO parallel channels, float and double I/O.  One hb_matrix handle holds the whole matrix on the
// GPU: N forward FFTs, one frequency-domain multiply-accumulate over (input x partition) for every
// output, M inverse FFTs per hop -- instead of N*M independent MonoConvolve objects.
// Superset: custom partition sizes (as MonoConvolve.h:31) and a pre-sized allocation.
#pragma once

#include "NToMonoConvolve.h"
#include "ConvolveErrors.h"

#include <algorithm>
#include <cstdint>
#include <vector>

namespace HISSTools
{
    class Convolver
    {
    public:

        // the reference gives every pair room for 16384 taps (Convolver.cpp:18,35); longer IRs need set(..., resize = true)
        Convolver(uint32_t numIns, uint32_t numOuts, LatencyMode latency)
        : mNumIns(std::max(numIns, 1u)), mNumOuts(numOuts), mN2M(true), mMatrix(1, std::max(numIns, 1u), numOuts, 16384, latency) {}
        Convolver(uint32_t numIO, LatencyMode latency)
        : mNumIns(std::max(numIO, 1u)), mNumOuts(std::max(numIO, 1u)), mN2M(false), mMatrix(std::max(numIO, 1u), 1, 1, 16384, latency) {}
        Convolver(uint32_t numIns, uint32_t numOuts, uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B = 0, uint32_t C = 0, uint32_t D = 0)
        : mNumIns(std::max(numIns, 1u)), mNumOuts(numOuts), mN2M(true), mMatrix(1, std::max(numIns, 1u), numOuts, maxLength, zeroLatency, A, B, C, D) {}

        // Superset: the same object dealt to several GPUs of this process -- `devices` lists CUDA device ordinals.  The input channels
        // of an N x M matrix (numIns and numOuts multiples of the device count) or the channels of a parallel convolver are shared
        // out; set / resize / process keep their reference signatures and take all channels, as on one device.
        Convolver(uint32_t numIns, uint32_t numOuts, LatencyMode latency, const std::vector<int>& devices)
        : mNumIns(std::max(numIns, 1u)), mNumOuts(numOuts), mN2M(true), mMatrix(1, std::max(numIns, 1u), numOuts, 16384, latency, devices) {}
        Convolver(uint32_t numIO, LatencyMode latency, const std::vector<int>& devices)
        : mNumIns(std::max(numIO, 1u)), mNumOuts(std::max(numIO, 1u)), mN2M(false), mMatrix(std::max(numIO, 1u), 1, 1, 16384, latency, devices) {}
        Convolver(uint32_t numIns, uint32_t numOuts, uintptr_t maxLength, const std::vector<int>& devices, bool zeroLatency, uint32_t A, uint32_t B = 0, uint32_t C = 0, uint32_t D = 0)
        : mNumIns(std::max(numIns, 1u)), mNumOuts(numOuts), mN2M(true), mMatrix(1, std::max(numIns, 1u), numOuts, maxLength, zeroLatency, A, B, C, D, devices) {}

        virtual ~Convolver() throw() {}

        // Clear IRs (Convolver.cpp:51-71)
        void clear(bool resize)
        {
            if (mN2M)
            {
                for (uint32_t i = 0; i < mNumOuts; i++)
                    for (uint32_t j = 0; j < mNumIns; j++)
                        clear(j, i, resize);
            }
            else
                for (uint32_t i = 0; i < mNumOuts; i++) clear(i, i, resize);
        }
        void clear(uint32_t inChan, uint32_t outChan, bool resize) { set(inChan, outChan, static_cast<const float *>(nullptr), 0, resize); }

        // DSP Engine Reset (Convolver.cpp:75-97)
        void reset() { hb_matrix_reset(mMatrix.handle()); }
        ConvolveError reset(uint32_t inChan, uint32_t outChan)
        {
            uint32_t g, i, o;
            const ConvolveError err = pair(inChan, outChan, g, i, o);
            return err ? err : b200::to_error(hb_matrix_reset_pair(mMatrix.handle(), g, i, o));
        }

        // Resize and set IR (Convolver.cpp:101-134)
        ConvolveError resize(uint32_t inChan, uint32_t outChan, uintptr_t impulseLength)
        {
            uint32_t g, i, o;
            // the reference reports any bad channel as IN_CHAN here (Convolver.cpp:109)
            if (pair(inChan, outChan, g, i, o)) return CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE;
            return b200::to_error(hb_matrix_resize(mMatrix.handle(), g, i, o, impulseLength));
        }
        ConvolveError set(uint32_t inChan, uint32_t outChan, const float* input, uintptr_t length, bool resize) { return setT(inChan, outChan, input, HB_F32, length, resize); }
        ConvolveError set(uint32_t inChan, uint32_t outChan, const double* input, uintptr_t length, bool resize) { return setT(inChan, outChan, input, HB_F64, length, resize); }
        void setResetOffset(intptr_t offset = -1) { hb_matrix_set_reset_offset(mMatrix.handle(), offset); }

        // DSP (Convolver.cpp:138-183): outputs below numOuts are always written (silence when nothing is loaded)
        void process(const float * const*  ins, float** outs, size_t numIns, size_t numOuts, size_t numSamples)
        {
            numIns = std::min<size_t>(numIns, mNumIns);
            numOuts = std::min<size_t>(numOuts, mNumOuts);
            mIn.assign(mNumIns, nullptr);
            mOut.assign(mNumOuts, nullptr);
            for (size_t i = 0; i < numIns; i++) mIn[i] = ins[i];
            for (size_t i = 0; i < numOuts; i++) { std::fill_n(outs[i], numSamples, 0.f); mOut[i] = outs[i]; }
            const int code = hb_matrix_process(mMatrix.handle(), mIn.data(), mOut.data(), numSamples, 1);
            if (code != HB_ERR_NO_IR && code != HB_ERR_BUSY) hisstools_b200_detail::check(code);
        }
        // double I/O casts through float, as the reference (Convolver.cpp:156-183)
        void process(const double * const* ins, double** outs, size_t numIns, size_t numOuts, size_t numSamples)
        {
            numIns = std::min<size_t>(numIns, mNumIns);
            numOuts = std::min<size_t>(numOuts, mNumOuts);
            mTemp.resize((numIns + numOuts) * numSamples);
            std::vector<const float *> fin(numIns);
            std::vector<float *> fout(numOuts);
            for (size_t i = 0; i < numIns; i++)
            {
                float *row = mTemp.data() + i * numSamples;
                for (size_t k = 0; k < numSamples; k++) row[k] = static_cast<float>(ins[i][k]);
                fin[i] = row;
            }
            for (size_t i = 0; i < numOuts; i++) fout[i] = mTemp.data() + (numIns + i) * numSamples;
            process(fin.data(), fout.data(), numIns, numOuts, numSamples);
            for (size_t i = 0; i < numOuts; i++)
                for (size_t k = 0; k < numSamples; k++) outs[i][k] = fout[i][k];
        }

        hb_matrix *handle() { return mMatrix.handle(); }

    private:

        // (group, in, out) of the engine for reference channel indices, or an error (Convolver.cpp:86-124)
        ConvolveError pair(uint32_t inChan, uint32_t outChan, uint32_t &g, uint32_t &i, uint32_t &o) const
        {
            if (!mN2M) inChan -= outChan;            // parallel mode: callers pass the same channel twice (Convolver.cpp:92,106,118)
            if (outChan >= mNumOuts) return CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE;
            if (inChan >= (mN2M ? mNumIns : 1u)) return CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE;
            g = mN2M ? 0 : outChan; i = mN2M ? inChan : 0; o = mN2M ? outChan : 0;
            return CONVOLVE_ERR_NONE;
        }
        ConvolveError setT(uint32_t inChan, uint32_t outChan, const void *input, int dtype, uintptr_t length, bool resize)
        {
            uint32_t g, i, o;
            const ConvolveError err = pair(inChan, outChan, g, i, o);
            return err ? err : b200::to_error(hb_matrix_set(mMatrix.handle(), g, i, o, input, dtype, length, resize ? 1 : 0));
        }

        uint32_t mNumIns;
        uint32_t mNumOuts;
        bool mN2M;
        b200::Matrix mMatrix;
        std::vector<const void *> mIn;
        std::vector<void *> mOut;
        std::vector<float> mTemp;
    };
}
