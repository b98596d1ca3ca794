// hb_fft_block.cuh -- device-side CTA-wide FFT in shared memory built from hb_fft_core.cuh.
#pragma once

#include "hb_fft_core.cuh"

namespace hb
{

// In-place forward complex FFT of 2^log2m points held in shared memory `s` (padded indexing
// sidx<PADSH>).  Every thread of the CTA must call it; blockDim.x * EPT >= 2^log2m.
// Ends with a barrier: results are visible to all threads on return.
template <class T, int EPT, int PADSH>
__device__ __forceinline__ void block_fft(Cx<T> *s, int log2m, const Cx<T> *__restrict__ tw, int tw_log2)
{
    const uint32_t M = 1u << log2m;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    Cx<T> v[EPT];
    uint32_t Ns = 1;
    int done = 0;
    while (done < log2m)
    {
        const int R = next_radix(log2m - done);
        if (R == 8)
        {
            pass_load<T, EPT, 8, PADSH>(s, M, tid, nthr, v);
            __syncthreads();
            pass_store<T, EPT, 8, PADSH>(s, M, Ns, done + 3, tid, nthr, v, tw, tw_log2);
            done += 3; Ns <<= 3;
        }
        else if (R == 4)
        {
            pass_load<T, EPT, 4, PADSH>(s, M, tid, nthr, v);
            __syncthreads();
            pass_store<T, EPT, 4, PADSH>(s, M, Ns, done + 2, tid, nthr, v, tw, tw_log2);
            done += 2; Ns <<= 2;
        }
        else
        {
            pass_load<T, EPT, 2, PADSH>(s, M, tid, nthr, v);
            __syncthreads();
            pass_store<T, EPT, 2, PADSH>(s, M, Ns, done + 1, tid, nthr, v, tw, tw_log2);
            done += 1; Ns <<= 1;
        }
        __syncthreads();
    }
}

// default shared-memory padding: one slot per 32 elements
#ifndef HB_PADSH
#define HB_PADSH 5
#endif

// largest complex length the single-CTA shared-memory path handles (227 KB per CTA on sm_100a)
template <class T> struct SmemFftLimit;
template <> struct SmemFftLimit<float>  { static constexpr int max_log2m = 14; };   // 16384 x 8 B = 128 KiB
template <> struct SmemFftLimit<double> { static constexpr int max_log2m = 13; };   //  8192 x 16 B = 128 KiB

inline int fft_threads(int log2m, int ept)
{
    int t = (1 << log2m) / ept;
    if (t < 32) t = 32;
    return (t + 31) & ~31;
}

} // namespace hb
