// hb_fft_big.cuh -- complex FFTs too long for one CTA's shared memory (the reference transforms any
// size its setup was built for, HISSTools_FFT_Core.h:1293-1374; PartitionedConvolve allows FFT sizes up
// to 2^20, PartitionedConvolve.h:18-19).
//
// Four-step decomposition of M = M1 * M2 points (both powers of two, M1 >= M2), two launches over an
// interleaved complex array in global memory:
//   k_big_cols  M2 column transforms of length M1 (stride M2), times the inter-step twiddle w_M^(n2 k1)
//   k_big_rows  M1 row transforms of length M2, written in natural order X[k1 + M1 k2]
// Every CTA (256 threads x 8 points) holds a tile of 2048 points = several adjacent sub-transforms, so
// that global loads and stores touch runs of adjacent addresses; the sub-transforms are the same
// shared-memory Stockham passes as the single-CTA path (hb_fft_core.cuh).
#pragma once

#include "hb_common.cuh"
#include "hb_fft_block.cuh"

namespace hb
{

constexpr int BIG_THREADS = 256;
constexpr int BIG_EPT = 8;
constexpr int BIG_TILE = BIG_THREADS * BIG_EPT;      // points per CTA
constexpr int BIG_MAX_LOG2 = 22;                     // largest complex transform: sub-transforms stay <= 2048 points

// split of m = log2(M): m1 = ceil(m/2) (columns), m2 = m - m1 (rows)
inline int big_m1(int m) { return (m + 1) / 2; }

// Stockham passes over `count` sub-arrays of 2^log2l points laid out one after the other (stride lp) in
// shared memory; thread `tid` of the CTA works on sub-array tid / nthr_sub.  All threads call it.
template <class T, int EPT, int PADSH>
__device__ __forceinline__ void block_fft_sub(Cx<T> *s, uint32_t lp, int log2l, const Cx<T> *__restrict__ tw, int tw_log2)
{
    const uint32_t L = 1u << log2l;
    const uint32_t nthr_sub = L / EPT ? L / EPT : 1;
    const uint32_t sub = threadIdx.x / nthr_sub, tid = threadIdx.x - sub * nthr_sub;
    Cx<T> *ss = s + size_t(sub) * lp;
    Cx<T> v[EPT];
    uint32_t Ns = 1;
    int done = 0;
    while (done < log2l)
    {
        const int R = next_radix(log2l - done);
        if (R == 8)
        {
            pass_load<T, EPT, 8, PADSH>(ss, L, tid, nthr_sub, v);
            __syncthreads();
            pass_store<T, EPT, 8, PADSH>(ss, L, Ns, done + 3, tid, nthr_sub, v, tw, tw_log2);
            done += 3; Ns <<= 3;
        }
        else if (R == 4)
        {
            pass_load<T, EPT, 4, PADSH>(ss, L, tid, nthr_sub, v);
            __syncthreads();
            pass_store<T, EPT, 4, PADSH>(ss, L, Ns, done + 2, tid, nthr_sub, v, tw, tw_log2);
            done += 2; Ns <<= 2;
        }
        else
        {
            pass_load<T, EPT, 2, PADSH>(ss, L, tid, nthr_sub, v);
            __syncthreads();
            pass_store<T, EPT, 2, PADSH>(ss, L, Ns, done + 1, tid, nthr_sub, v, tw, tw_log2);
            done += 1; Ns <<= 1;
        }
        __syncthreads();
    }
}

// step A.  grid = (M2 / CW, batch), CW = BIG_TILE / M1 adjacent columns per CTA.
template <class T>
__global__ void __launch_bounds__(BIG_THREADS) k_big_cols(const Cx<T> *__restrict__ in, Cx<T> *__restrict__ out, int m, int m1,
                                                          const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t M1 = 1u << m1, M2 = 1u << (m - m1);
    const uint32_t CW = BIG_TILE / M1, lp = padded_elems<HB_PADSH>(M1);
    const uint32_t c0 = blockIdx.x * CW;
    const size_t base = size_t(blockIdx.y) << m;
    Cx<T> v[BIG_EPT];
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        const uint32_t sub = idx % CW, n1 = idx / CW;
        v[e] = in[base + size_t(n1) * M2 + c0 + sub];
    }
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        const uint32_t sub = idx % CW, n1 = idx / CW;
        s[sub * lp + sidx<HB_PADSH>(n1)] = v[e];
    }
    __syncthreads();
    block_fft_sub<T, BIG_EPT, HB_PADSH>(s, lp, m1, tw, tw_log2);
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        const uint32_t sub = idx % CW, k1 = idx / CW;
        const uint32_t n2 = c0 + sub;
        const Cx<T> w = tw_root(tw, tw_log2, n2 * k1, m);            // n2 * k1 < M
        out[base + size_t(k1) * M2 + n2] = cmul(s[sub * lp + sidx<HB_PADSH>(k1)], w);
    }
}

// step B.  grid = (M1 / RW, batch), RW = BIG_TILE / M2 adjacent rows per CTA.
template <class T>
__global__ void __launch_bounds__(BIG_THREADS) k_big_rows(const Cx<T> *__restrict__ in, Cx<T> *__restrict__ out, int m, int m1,
                                                          const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const int m2 = m - m1;
    const uint32_t M1 = 1u << m1, M2 = 1u << m2;
    const uint32_t RW = BIG_TILE / M2, lp = padded_elems<HB_PADSH>(M2);
    const uint32_t r0 = blockIdx.x * RW;
    const size_t base = size_t(blockIdx.y) << m;
    Cx<T> v[BIG_EPT];
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        v[e] = in[base + size_t(r0) * M2 + idx];                      // RW whole rows are one contiguous run
    }
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        const uint32_t row = idx >> m2, n2 = idx & (M2 - 1);
        s[row * lp + sidx<HB_PADSH>(n2)] = v[e];
    }
    __syncthreads();
    block_fft_sub<T, BIG_EPT, HB_PADSH>(s, lp, m2, tw, tw_log2);
#pragma unroll
    for (int e = 0; e < BIG_EPT; e++)
    {
        const uint32_t idx = threadIdx.x + e * BIG_THREADS;
        const uint32_t row = idx % RW, k2 = idx / RW;
        out[base + size_t(k2) * M1 + r0 + row] = s[row * lp + sidx<HB_PADSH>(k2)];
    }
}

// split / merge pass of a real transform of N = 2M points on the interleaved array (pairs k, M - k):
// the global-memory form of block_real_split.  grid = (ceil((M/2 + 1) / 256), batch).
template <class T>
__global__ void k_big_split(Cx<T> *__restrict__ z, int m, int inverse, const Cx<T> *__restrict__ tw, int tw_log2)
{
    const uint32_t M = 1u << m;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= M / 2) real_split_pair<T, 31>(z + (size_t(blockIdx.y) << m), M, m + 1, k, inverse != 0, tw, tw_log2);
}

template <class T> inline size_t big_smem_bytes(int m)
{
    const int m1 = big_m1(m), m2 = m - m1;
    const size_t a = size_t(BIG_TILE >> m1) * padded_elems<HB_PADSH>(1u << m1);
    const size_t b = size_t(BIG_TILE >> m2) * padded_elems<HB_PADSH>(1u << m2);
    return (a > b ? a : b) * sizeof(Cx<T>);
}

// forward complex FFT of `batch` arrays of 2^m points: in -> out via scratch (in may equal out; scratch
// must be distinct from both).  12 <= m <= BIG_MAX_LOG2, tw_log2 >= m.
template <class T>
int big_cfft(const Cx<T> *in, Cx<T> *scratch, Cx<T> *out, int m, size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st)
{
    if (m < 12 || m > BIG_MAX_LOG2 || tw_log2 < m) { set_error("internal: four-step FFT of 2^%d points (table order %d)", m, tw_log2); return HB_ERR_UNSUPPORTED; }
    const int m1 = big_m1(m), m2 = m - m1;
    const size_t smem = big_smem_bytes<T>(m);
    if (smem > 48 * 1024)
    {
        HB_CUDA(cudaFuncSetAttribute(k_big_cols<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        HB_CUDA(cudaFuncSetAttribute(k_big_rows<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    }
    const unsigned cw = BIG_TILE >> m1, rw = BIG_TILE >> m2;
    k_big_cols<T><<<dim3((1u << m2) / cw, (unsigned) batch), BIG_THREADS, smem, st>>>(in, scratch, m, m1, tw, tw_log2);
    HB_LAUNCH_CHECK();
    k_big_rows<T><<<dim3((1u << m1) / rw, (unsigned) batch), BIG_THREADS, smem, st>>>(scratch, out, m, m1, tw, tw_log2);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int big_split(Cx<T> *z, int m, int inverse, size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st)
{
    const unsigned pairs = (1u << m) / 2 + 1;
    k_big_split<T><<<dim3((pairs + 255) / 256, (unsigned) batch), 256, 0, st>>>(z, m, inverse, tw, tw_log2);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// exchange real and imaginary parts in place (the unscaled inverse is a forward transform of the exchanged planes)
template <class T>
__global__ void k_big_exchange(Cx<T> *__restrict__ z, size_t count)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
    {
        const Cx<T> v = z[i];
        z[i] = cx<T>(v.y, v.x);
    }
}

// scratch of the four-step path: two interleaved arrays of batch * M points
struct BigScratch
{
    DevBuf z1, z2;
    template <class T> int ensure(int m, size_t batch)
    {
        const size_t bytes = (batch << m) * sizeof(Cx<T>);
        int rc;
        if ((rc = z1.ensure(bytes)) || (rc = z2.ensure(bytes))) return rc;
        return HB_OK;
    }
    void release() { z1.release(); z2.release(); }
};

inline dim3 big_grid(int m, size_t batch)
{
    const size_t blocks = ((size_t(1) << m) + 255) / 256;
    return dim3((unsigned) (blocks < 1024 ? blocks : 1024), (unsigned) batch);
}

} // namespace hb
