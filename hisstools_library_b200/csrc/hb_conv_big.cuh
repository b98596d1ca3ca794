// hb_conv_big.cuh -- the convolution engine at FFT sizes whose half-length complex transform does not fit one
// CTA's shared memory (real FFT above 2^15 float / 2^14 double, up to the reference's 2^20,
// PartitionedConvolve.h:18-19).  k_fwd / k_inv / k_ir become short chains of launches around the four-step
// transform of hb_fft_big.cuh over an interleaved scratch array z[channel][B]:
//   forward  pack frame -> cols, rows -> split -> scatter into the FDL tiles (+ Nyquist side array)
//   inverse  Nyquist dot products -> sum of the stream-K partials into z (+ hand-over of the previous block)
//            -> inverse split -> plane exchange -> cols, rows -> scale and store the first B samples
//   IR       pack chunks (one per partition) -> cols, rows -> split -> scatter into the unit layout
// The multiply-accumulate kernels are size-agnostic and shared with the small path.
#pragma once

#include "hb_conv_kernels.cuh"
#include "hb_fft_big.cuh"

namespace hb
{

// rotated frame [newest B | previous B] of every input channel, de-interleaved (k_fwd's load stage)
template <class T>
__global__ void k_bigc_pack(const Geom g, const T *__restrict__ prev, size_t prev_ld, const T *__restrict__ newest, size_t new_ld,
                            T *__restrict__ save, size_t save_ld, Cx<T> *__restrict__ z)
{
    const uint32_t ch = blockIdx.y, B = g.B;
    const T *pn = newest + size_t(ch) * new_ld, *pp = prev + size_t(ch) * prev_ld;
    T *ps = save ? save + size_t(ch) * save_ld : nullptr;
    Cx<T> *zc = z + size_t(ch) * B;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x)
    {
        const uint32_t j = 2 * k;
        T a, b;
        if (j < B)
        {
            a = pn[j]; b = pn[j + 1];
            if (ps) { ps[j] = a; ps[j + 1] = b; }
        }
        else { a = pp[j - B]; b = pp[j - B + 1]; }
        zc[k] = cx<T>(a, b);
    }
}

// packed spectrum of every input channel -> newest FDL slot (k_fwd's store stage)
template <class T>
__global__ void k_bigc_to_fdl(const Geom g, const Cx<T> *__restrict__ z, Cx<T> *__restrict__ X, T *__restrict__ Xnyq)
{
    const uint32_t ch = blockIdx.y, B = g.B;
    const uint32_t TB = tile_bins<T>(g);
    const Cx<T> *zc = z + size_t(ch) * B;
    Cx<T> *xrow = X + size_t(ch) * g.n_bt * g.R * TB;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x)
    {
        Cx<T> v = zc[k];
        if (k == 0)
        {
            Xnyq[size_t(ch) * g.R + g.slot] = v.y;
            v.y = T(0);
        }
        const uint32_t bt = k / TB, j = k - bt * TB;
        xrow[(size_t(bt) * g.R + g.slot) * TB + j] = v;
    }
}

// chunks of one impulse response, one per partition (blockIdx.y), zero-padded and de-interleaved
template <class T>
__global__ void k_bigc_ir_pack(const Geom g, const T *__restrict__ ir, size_t taps, Cx<T> *__restrict__ z)
{
    const uint32_t p = blockIdx.y, B = g.B;
    const size_t begin = size_t(p) * B;
    const size_t have = taps > begin ? (taps - begin < B ? taps - begin : B) : 0;
    const T *src = ir + begin;
    Cx<T> *zc = z + size_t(p) * B;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x)
    {
        const size_t j = 2 * size_t(k);
        zc[k] = cx<T>(j < have ? src[j] : T(0), j + 1 < have ? src[j + 1] : T(0));
    }
}

// partition spectra of pair (grp, in, o) -> unit layout (k_ir's store stage)
template <class T>
__global__ void k_bigc_to_h(const Geom g, const Cx<T> *__restrict__ z, uint32_t grp, uint32_t in, uint32_t o,
                            Cx<T> *__restrict__ H, T *__restrict__ Hnyq)
{
    const uint32_t p = blockIdx.y, B = g.B;
    const uint32_t TB = tile_bins<T>(g);
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    const Cx<T> *zc = z + size_t(p) * B;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x)
    {
        Cx<T> v = zc[k];
        if (k == 0)
        {
            Hnyq[((size_t(grp) * g.outs + o) * g.ins + in) * g.Pcap + p] = v.y;
            v.y = T(0);
        }
        const uint32_t bt = k / TB, j = k - bt * TB;
        const uint64_t tile = (uint64_t(grp) * g.n_ot + ot) * g.n_bt + bt;
        const uint64_t unit = (tile * g.ins + in) * g.Pcap + p;
        H[unit * (uint64_t(g.Q) * VecOf<T>::CPV) + size_t(row) * TB + j] = v;
    }
}

// Nyquist bin of every output: the real dot product over (in, partition) that k_inv does in its prologue
template <class T>
__global__ void __launch_bounds__(256) k_bigc_nyq(const Geom g, const T *__restrict__ Xnyq, const T *__restrict__ Hnyq, T *__restrict__ nyq)
{
    __shared__ T red[40];
    const uint32_t ch = blockIdx.x;
    const uint32_t grp = ch / g.outs, o = ch - grp * g.outs;
    const T *hn = Hnyq + (size_t(grp) * g.outs + o) * g.ins * g.Pcap;
    const T *xn = Xnyq + size_t(grp) * g.ins * g.R;
    T part = T(0);
    for (uint32_t idx = threadIdx.x; idx < g.upt; idx += blockDim.x)
    {
        const uint32_t in = idx / g.P, p = idx - in * g.P;
        uint32_t sl = g.slot + p;
        if (sl >= g.R) sl -= g.R;
        part += xn[size_t(in) * g.R + sl] * hn[size_t(in) * g.Pcap + p];
    }
    const T total = block_sum<T>(part, red);
    if (threadIdx.x == 0) nyq[ch] = total;
}

// sum of the stream-K partial segments of one spectrum vector (segment order = CTA order, set after set)
template <class T>
__device__ __forceinline__ typename VecOf<T>::type sum_segments(const Geom &g, const SegSets &sets, uint32_t grp, uint32_t ot, uint32_t row, uint32_t v)
{
    typedef typename VecOf<T>::type V;
    V sum;
    vzero(sum);
    const uint32_t bt = v / g.TBV, xa = v - bt * g.TBV;
    const uint32_t tile = (grp * g.n_ot + ot) * g.n_bt + bt;
    const uint64_t base = uint64_t(tile) * g.Q + row * g.TBV + xa;
    for (int q = 0; q < sets.n; q++)
    {
        const V *__restrict__ S = reinterpret_cast<const V *>(sets.s[q].S);
        const uint64_t ulo = uint64_t(tile) * sets.s[q].upt, uhi = ulo + sets.s[q].upt - 1;
        const uint32_t lo = (uint32_t) unit_owner(ulo, sets.s[q].U, sets.s[q].G), hi = (uint32_t) unit_owner(uhi, sets.s[q].U, sets.s[q].G);
        for (uint32_t c = lo; c <= hi; c++) vadd(sum, S[uint64_t(c) * g.Q + base]);
    }
    return sum;
}

// k_inv's first stage: hand the previous block to the caller, then the packed output spectrum into z
template <class T>
__global__ void k_bigc_gather(const Geom g, const SegSets sets, const T *__restrict__ nyq, Cx<T> *__restrict__ z,
                              const T *__restrict__ carry_src, size_t carry_src_ld, T *__restrict__ carry_dst, size_t carry_dst_ld, int add_carry)
{
    typedef typename VecOf<T>::type V;
    constexpr int CPV = VecOf<T>::CPV;
    const uint32_t ch = blockIdx.y, B = g.B;
    const uint32_t grp = ch / g.outs, o = ch - grp * g.outs;
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    if (carry_dst)
    {
        const T *cs = carry_src + size_t(ch) * carry_src_ld;
        T *cd = carry_dst + size_t(ch) * carry_dst_ld;
        for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B; k += gridDim.x * blockDim.x) cd[k] = add_carry ? cd[k] + cs[k] : cs[k];
    }
    Cx<T> *zc = z + size_t(ch) * B;
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < B / CPV; v += gridDim.x * blockDim.x)
    {
        const V sum = sum_segments<T>(g, sets, grp, ot, row, v);
        if constexpr (CPV == 2)
        {
            zc[2 * v] = cx<T>(sum.x, v == 0 ? nyq[ch] : sum.y);
            zc[2 * v + 1] = cx<T>(sum.z, sum.w);
        }
        else
            zc[v] = cx<T>(sum.x, v == 0 ? nyq[ch] : sum.y);
    }
}

// k_inv's last stage: scale 1/(4N) and the first B samples of the planes-exchanged transform to the output row
template <class T>
__global__ void k_bigc_store(const Geom g, const Cx<T> *__restrict__ z, T *__restrict__ yout, size_t ld, size_t off, int add_result)
{
    const uint32_t ch = blockIdx.y, B = g.B;
    const T scale = T(1) / T(size_t(4) << g.log2n);
    const Cx<T> *zc = z + size_t(ch) * B;
    T *dst = yout + size_t(ch) * ld + off;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < B / 2; k += gridDim.x * blockDim.x)
    {
        const Cx<T> v = zc[k];
        if (add_result) { dst[2 * k] += v.y * scale; dst[2 * k + 1] += v.x * scale; }
        else { dst[2 * k] = v.y * scale; dst[2 * k + 1] = v.x * scale; }
    }
}

} // namespace hb
