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

// Real <-> half-complex split pass over the whole array (pairs k and M-k, k = 0 .. M/2), every thread
// taking up to EPT/2 + 1 pairs.  All shared-memory and twiddle loads of a thread are issued before the
// first dependent arithmetic (the pairs are disjoint), so their latencies overlap.  Callers put a
// barrier before and after.
template <class T, int EPT, int PADSH>
__device__ __forceinline__ void block_real_split(Cx<T> *s, uint32_t M, int log2N, bool inverse, const Cx<T> *__restrict__ tw, int tw_log2)
{
    constexpr int NP = EPT / 2 + 1;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    Cx<T> a[NP], b[NP], w[NP];
#pragma unroll
    for (int e = 0; e < NP; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k <= M / 2)
        {
            a[e] = s[sidx<PADSH>(k)];
            b[e] = s[sidx<PADSH>(k ? M - k : 0)];
            w[e] = tw_root(tw, tw_log2, k, log2N);
        }
    }
#pragma unroll
    for (int e = 0; e < NP; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k <= M / 2)
        {
            if (k == 0)
            {
                const T t1 = a[e].x + a[e].y, t2 = a[e].x - a[e].y;
                s[sidx<PADSH>(0)] = inverse ? cx<T>(t1, t2) : cx<T>(t1 + t1, t2 + t2);
            }
            else
            {
                const T sr = a[e].x + b[e].x, si = a[e].y + b[e].y, dr = a[e].x - b[e].x, di = a[e].y - b[e].y;
                const T wr = inverse ? -w[e].x : w[e].x, wi = w[e].y;
                const T u = wr * si + wi * dr;
                const T v = wi * si - wr * dr;
                s[sidx<PADSH>(k)] = cx<T>(sr + u, v + di);
                s[sidx<PADSH>(M - k)] = cx<T>(sr - u, v - di);
            }
        }
    }
}

// Roots of order 2^log2n (half circle, 2^(log2n-1) entries = one per point of the complex array of a real
// transform) copied from the device table into shared memory, EPT per thread.  The loads are returned in
// registers so that the caller can issue them together with its own input loads and store later.
template <class T, int EPT>
__device__ __forceinline__ void twiddle_stage_load(Cx<T> *r, const Cx<T> *__restrict__ tw, int tw_log2, int log2n)
{
    const uint32_t count = 1u << (log2n - 1);
    const int sh = tw_log2 - log2n;
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t q = threadIdx.x + e * blockDim.x;
        if (q < count) r[e] = tw[size_t(q) << sh];
    }
}
template <class T, int EPT>
__device__ __forceinline__ void twiddle_stage_store(Cx<T> *stw, const Cx<T> *r, int log2n)
{
    const uint32_t count = 1u << (log2n - 1);
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t q = threadIdx.x + e * blockDim.x;
        if (q < count) stw[q] = r[e];
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

// Points per thread for a transform of 2^log2m complex points: 8 up to 4096 points, then as many as keep
// the CTA at 512 threads (every FFT kernel is compiled with __launch_bounds__(512): 128 registers per thread).
// Usage: HB_EPT_DISPATCH(log2m, launch<T, EPT>(...));
#define HB_EPT_DISPATCH(log2m, ...)                                            \
    do                                                                         \
    {                                                                          \
        if ((log2m) <= 12) { constexpr int EPT = 8; __VA_ARGS__; }             \
        else if ((log2m) == 13) { constexpr int EPT = 16; __VA_ARGS__; }       \
        else { constexpr int EPT = 32; __VA_ARGS__; }                          \
    } while (0)

} // namespace hb
