// ConvolveErrors.h -- error codes of the convolver classes; values identical to the reference
// (HIRT_Multichannel_Convolution/ConvolveErrors.h:4-19) because they are what the C ABI returns.
#pragma once

enum ConvolveError
{
    CONVOLVE_ERR_NONE = 0,
    CONVOLVE_ERR_IN_CHAN_OUT_OF_RANGE = 1,
    CONVOLVE_ERR_OUT_CHAN_OUT_OF_RANGE = 2,
    CONVOLVE_ERR_MEM_UNAVAILABLE = 3,
    CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL = 4,
    CONVOLVE_ERR_TIME_IMPULSE_TOO_LONG = 5,
    CONVOLVE_ERR_TIME_LENGTH_OUT_OF_RANGE = 6,
    CONVOLVE_ERR_PARTITION_LENGTH_TOO_LARGE = 7,
    CONVOLVE_ERR_FFT_SIZE_MAX_TOO_SMALL = 8,
    CONVOLVE_ERR_FFT_SIZE_MAX_TOO_LARGE = 9,
    CONVOLVE_ERR_FFT_SIZE_MAX_NON_POWER_OF_TWO = 10,
    CONVOLVE_ERR_FFT_SIZE_OUT_OF_RANGE = 11,
    CONVOLVE_ERR_FFT_SIZE_NON_POWER_OF_TWO = 12,
};

namespace HISSTools
{
    namespace b200
    {
        // a negative hb_status (CUDA failure, bad argument) has no reference code: report it as memory unavailable
        inline ConvolveError to_error(int code) { return code < 0 ? CONVOLVE_ERR_MEM_UNAVAILABLE : static_cast<ConvolveError>(code); }
    }
}
