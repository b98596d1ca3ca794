// MonoConvolve.h -- B200 drop-in for HISSTools::MonoConvolve
// (reference: HIRT_Multichannel_Convolution/MonoConvolve.h:14-48, .cpp:18-258): a non-uniform
// partition scheme (zero-latency direct-form head + FFT parts of 256/1024/4096/16384 or custom sizes).
// One hb_matrix handle (include/hisstools_b200.h) with a single pair does the work on the GPU.
#pragma once

#include "PartitionedConvolve.h"
#include "ConvolveErrors.h"

#include <cstdint>
#include <stdexcept>
#include <vector>

enum LatencyMode
{
    kLatencyZero,
    kLatencyShort,
    kLatencyMedium,
} ;

namespace HISSTools
{
    namespace b200
    {
        // owner of one hb_matrix: groups banks of ins x outs pairs sharing a partition scheme
        class Matrix
        {
        public:

            Matrix() : mHandle(nullptr) {}
            Matrix(uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t maxLength, LatencyMode latency, int dtype = HB_F32, int device = 0) : mHandle(nullptr)
            {
                hisstools_b200_detail::check(hb_matrix_create_latency(&mHandle, dtype, groups, ins, outs, maxLength, static_cast<int>(latency), device));
            }
            Matrix(uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B, uint32_t C, uint32_t D,
                   int dtype = HB_F32, int device = 0) : mHandle(nullptr)
            {
                create(groups, ins, outs, maxLength, zeroLatency, A, B, C, D, dtype, device);
            }
            // the same matrix dealt to several GPUs of this process (hb_matrix_create_multi / _latency_multi)
            Matrix(uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t maxLength, LatencyMode latency, const std::vector<int>& devices, int dtype = HB_F32) : mHandle(nullptr)
            {
                hisstools_b200_detail::check(hb_matrix_create_latency_multi(&mHandle, dtype, groups, ins, outs, maxLength, static_cast<int>(latency),
                                                                            devices.data(), static_cast<uint32_t>(devices.size())));
            }
            Matrix(uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B, uint32_t C, uint32_t D,
                   const std::vector<int>& devices, int dtype = HB_F32) : mHandle(nullptr)
            {
                const int code = hb_matrix_create_multi(&mHandle, dtype, groups, ins, outs, maxLength, zeroLatency ? 1 : 0, A, B, C, D,
                                                        devices.data(), static_cast<uint32_t>(devices.size()));
                if (code == HB_ERR_BAD_ARG) throw std::runtime_error(hb_last_error());
                hisstools_b200_detail::check(code);
            }
            ~Matrix() { hb_matrix_destroy(mHandle); }

            Matrix(Matrix& obj) = delete;
            Matrix& operator = (Matrix& obj) = delete;
            Matrix(Matrix&& obj) : mHandle(obj.mHandle) { obj.mHandle = nullptr; }
            Matrix& operator = (Matrix&& obj)
            {
                if (this != &obj) { hb_matrix_destroy(mHandle); mHandle = obj.mHandle; obj.mHandle = nullptr; }
                return *this;
            }

            void create(uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B, uint32_t C, uint32_t D,
                        int dtype = HB_F32, int device = 0)
            {
                hb_matrix *fresh = nullptr;
                const int code = hb_matrix_create(&fresh, dtype, groups, ins, outs, maxLength, zeroLatency ? 1 : 0, A, B, C, D, device);
                // the reference throws std::runtime_error for an invalid size list (MonoConvolve.cpp:212,229)
                if (code == HB_ERR_BAD_ARG) throw std::runtime_error(hb_last_error());
                hisstools_b200_detail::check(code);
                hb_matrix_destroy(mHandle);
                mHandle = fresh;
            }

            hb_matrix *handle() const { return mHandle; }

        private:

            hb_matrix *mHandle;
        };
    }

    class MonoConvolve
    {
    public:

        MonoConvolve(uintptr_t maxLength, LatencyMode latency) : mMatrix(1, 1, 1, maxLength, latency) {}
        MonoConvolve(uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B = 0, uint32_t C = 0, uint32_t D = 0)
        : mMatrix(1, 1, 1, maxLength, zeroLatency, A, B, C, D) {}

        // Moveable but not copyable (MonoConvolve.h:35-38)
        MonoConvolve(MonoConvolve& obj) = delete;
        MonoConvolve& operator = (MonoConvolve& obj) = delete;
        MonoConvolve(MonoConvolve&& obj) = default;
        MonoConvolve& operator = (MonoConvolve&& obj) = default;

        void setResetOffset(intptr_t offset = -1) { hb_matrix_set_reset_offset(mMatrix.handle(), offset); }

        ConvolveError resize(uintptr_t length) { return b200::to_error(hb_matrix_resize(mMatrix.handle(), 0, 0, 0, length)); }
        ConvolveError set(const float *input, uintptr_t length, bool requestResize)
        {
            return b200::to_error(hb_matrix_set(mMatrix.handle(), 0, 0, 0, input, HB_F32, length, requestResize ? 1 : 0));
        }
        ConvolveError reset() { return b200::to_error(hb_matrix_reset(mMatrix.handle())); }

        // `temp` is kept for signature compatibility (MonoConvolve.h:46); the sum over parts happens on the device
        void process(const float *in, float *temp, float *out, uintptr_t numSamples, bool accumulate = false)
        {
            (void) temp;
            const void *ins[1] = { in };
            void *outs[1] = { out };
            const int code = hb_matrix_process(mMatrix.handle(), ins, outs, numSamples, accumulate ? 1 : 0);
            if (code != HB_ERR_NO_IR && code != HB_ERR_BUSY) hisstools_b200_detail::check(code);
        }

        void setPartitions(uintptr_t maxLength, bool zeroLatency, uint32_t A, uint32_t B = 0, uint32_t C = 0, uint32_t D = 0)
        {
            mMatrix.create(1, 1, 1, maxLength, zeroLatency, A, B, C, D);
        }

    private:

        b200::Matrix mMatrix;
    };
}
