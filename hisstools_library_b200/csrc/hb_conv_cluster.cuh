// hb_conv_cluster.cuh -- forward / inverse transforms of the convolution engine with every transform spread over a
// thread-block CLUSTER of 8 CTAs (distributed shared memory), for engines whose transforms are few and expensive.
//
// Why: one CTA per channel (k_fwd / k_inv) leaves a 16-channel double-precision engine (BASELINE config 5: 16 x 8192
// complex points) on 16 SMs, where each transform is a chain of dependent shared-memory passes and global round trips on
// 16 warps (ncu: 25 % issue slots busy, 16 cycles per issued instruction; 38 us forward, 60-86 us inverse beside the
// streaming multiply-accumulate -- profiles/r1_c5_cluster_fft.txt); the FFT kernels, not HBM, then set the hop period.
// Eight CTAs per transform put the same work on 128 SMs with an eighth of the passes' length each.
//
// Decomposition of the M-point complex transform (M = 8 L), decimation in time over the cluster rank r:
//   rank r      local L-point Stockham transform (hb_fft_block.cuh) of the decimated sequence z[8 n + r]  ->  Y_r[k2]
//   cluster barrier
//   every rank  for its share of the columns k2: the eight Y_r[k2] straight out of the peers' shared memory, times
//               W_M^(r k2), one radix-8 butterfly  ->  X[k2 + L k1], k1 = 0..7
// forward (k_fwd_cl): the columns of a rank are a mirror-symmetric set {k2} u {L - k2}, so that the real split pass
//   (pairs k, M - k; behaviour of HISSTools_FFT_Core.h:934-988) stays inside the CTA; spectra go to the newest FDL slot.
// inverse (k_inv_cl): every rank sums the stream-K partial segments of its decimated bins, fetches their mirrors (held by
//   rank 8 - r) through distributed shared memory, does the inverse split pass in registers, exchanges the planes
//   (Core:1341-1346), and after the cross pass keeps k1 = 0..3 only:
//   the first B samples are all the overlap-save scheme uses (PartitionedConvolve.cpp:235-240,357-360).
// Same layouts, same summation order of the partial segments as k_fwd / k_inv; single hops only (no multi-hop batches,
// no fused multi-GPU exchange -- the callers fall back to the one-CTA kernels for those).
#pragma once

#include <cooperative_groups.h>

#include "hb_conv_big.cuh"

namespace hb
{
namespace cgc = cooperative_groups;

constexpr int CL_CS = 8;           // CTAs per transform
constexpr int CL_EPT = 8;          // points per thread of the local transform
constexpr int CL_MIN_LOG2M = 11;   // local transforms of at least 256 points (one warp)
constexpr int CL_MAX_SEG_TILES = 1024;

template <class T> inline size_t cl_fwd_smem(int log2m)
{
    const uint32_t L = 1u << (log2m - 3);
    return (size_t(padded_elems<HB_PADSH>(L)) + L / 2 + L) * sizeof(Cx<T>);
}
template <class T> inline size_t cl_inv_smem(int log2m, uint32_t n_bt)
{
    const uint32_t L = 1u << (log2m - 3);
    return (2 * size_t(padded_elems<HB_PADSH>(L)) + L / 2) * sizeof(Cx<T>) + size_t(n_bt) * 4 * sizeof(uint32_t);
}

template <class T>
__device__ __forceinline__ void fdl_store(const Geom &g, Cx<T> *__restrict__ xrow, T *__restrict__ Xnyq, uint32_t ch, uint32_t slot, uint32_t TB, uint32_t k, Cx<T> z)
{
    if (k == 0)
    {
        Xnyq[size_t(ch) * g.R + slot] = z.y;
        z.y = T(0);
    }
    const uint32_t bt = k / TB, j = k - bt * TB;
    xrow[(size_t(bt) * g.R + slot) * TB + j] = z;
}

// ---------------------------------------------------------------------------------------------
// k_fwd_cl: grid = 8 x channels, cluster (8,1,1), L/8 threads.  Arguments as k_fwd.
// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void __cluster_dims__(CL_CS, 1, 1) __launch_bounds__(256)
k_fwd_cl(const Geom g, const T *__restrict__ prev, size_t prev_ld, const T *__restrict__ newest, size_t new_ld,
         T *__restrict__ save, size_t save_ld, Cx<T> *__restrict__ X, T *__restrict__ Xnyq, const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cgc::cluster_group cluster = cgc::this_cluster();
    const uint32_t r = cluster.block_rank();
    const uint32_t ch = blockIdx.x / CL_CS;
    const uint32_t B = g.B, L = B / CL_CS;
    const int m = (int) g.log2n - 1, l = m - 3;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;            // nthr == L / 8
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);                 // local transform (padded)
    Cx<T> *stw = s + padded_elems<HB_PADSH>(L);                     // roots of order L, half circle
    Cx<T> *xs = stw + L / 2;                                        // [8][nthr]: this rank's columns after the cross pass
    trace_mark(g, 0, 0);

    Cx<T> twr[CL_EPT];
    twiddle_stage_load<T, CL_EPT>(twr, tw, tw_log2, l);
    // decimated, de-interleaved rotated frame [newest B | previous B]: local point n holds z[8 n + r]
    const T *pn = newest + size_t(ch) * new_ld, *pp = prev + size_t(ch) * prev_ld;
    T *ps = save ? save + size_t(ch) * save_ld : nullptr;
    const bool vn = pair_aligned(newest, new_ld), vp = pair_aligned(prev, prev_ld), vs = save && pair_aligned(save, save_ld);
    {
        Pair<T> v[CL_EPT];
#pragma unroll
        for (int e = 0; e < CL_EPT; e++)
        {
            const uint32_t j = 2 * (CL_CS * (tid + e * nthr) + r);
            v[e] = j < B ? ld_pair(pn + j, vn) : ld_pair(pp + (j - B), vp);
        }
#pragma unroll
        for (int e = 0; e < CL_EPT; e++)
        {
            const uint32_t n = tid + e * nthr, j = 2 * (CL_CS * n + r);
            s[sidx<HB_PADSH>(n)] = cx<T>(v[e].a, v[e].b);
            if (ps && j < B) st_pair(ps + j, v[e].a, v[e].b, vs);   // becomes the previous hop of the next call
        }
    }
    twiddle_stage_store<T, CL_EPT>(stw, twr, l);
    __syncthreads();
    block_fft<T, CL_EPT, HB_PADSH>(s, l, stw, l);
    cluster.sync();                                                 // every Y_r is complete

    // cross pass for one column per thread.  Columns of rank r: k2 = r h + t (t < h) and their mirrors L - k2; the
    // mirror of column 0 is column 0 itself, its place (rank 0, thread h) takes the self-mirrored column L/2.
    const uint32_t h = nthr / 2;
    const uint32_t a0 = r * h + (tid < h ? tid : tid - h);
    const uint32_t k2 = tid < h ? a0 : (a0 ? L - a0 : L / 2);
    Cx<T> v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = cluster.map_shared_rank(s, q)[sidx<HB_PADSH>(k2)];
    {
        Cx<T> w[8];
#pragma unroll
        for (int q = 1; q < 8; q++) w[q] = tw_root(tw, tw_log2, q * k2, m);
#pragma unroll
        for (int q = 1; q < 8; q++) v[q] = cmul(v[q], w[q]);
    }
    dft8(v);                                                        // v[k1] = Z[k2 + L k1]
#pragma unroll
    for (int q = 0; q < 8; q++) xs[q * nthr + tid] = v[q];
    cluster.sync();                                                 // xs published; nobody reads a peer's s any more

    // split pass: pairs (k, M - k) = (column k2, k1) with (column L - k2, 7 - k1); each thread takes k1 = 0..3 of its own
    // column.  Column 0 pairs k1 with 8 - k1 (k1 = 0: the packed DC / Nyquist bin, k1 = 4: self-paired).
    const bool col0 = r == 0 && tid == 0, colh = r == 0 && tid == h;
    const uint32_t pslot = (col0 || colh) ? tid : (tid < h ? tid + h : tid - h);
    const uint32_t TB = tile_bins<T>(g);
    Cx<T> *xrow = X + size_t(ch) * g.n_bt * g.R * TB;
    const uint32_t slot = g.slot;
    const int npair = col0 ? 5 : 4;
    Cx<T> b[5], w[5];
#pragma unroll
    for (int k1 = 0; k1 < 5; k1++)
    {
        if (k1 < npair)
        {
            const uint32_t pk = col0 ? ((8 - k1) & 7) : 7 - k1;
            b[k1] = xs[pk * nthr + pslot];
            w[k1] = tw_root(tw, tw_log2, k2 + L * k1, m + 1);
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < 5; k1++)
    {
        if (k1 < npair)
        {
            const uint32_t k = k2 + L * k1, q = (B - k) & (B - 1);
            // v[] is indexed with compile-time k1 (registers); k1 == 4 happens for column 0 only
            const Cx<T> a = v[k1];
            if (k == 0)
            {
                const T t1 = a.x + a.y, t2 = a.x - a.y;
                fdl_store<T>(g, xrow, Xnyq, ch, slot, TB, 0, cx<T>(t1 + t1, t2 + t2));
            }
            else
            {
                const T sr = a.x + b[k1].x, si = a.y + b[k1].y, dr = a.x - b[k1].x, di = a.y - b[k1].y;
                const T wr = w[k1].x, wi = w[k1].y;
                const T u = wr * si + wi * dr;
                const T vv = wi * si - wr * dr;
                fdl_store<T>(g, xrow, Xnyq, ch, slot, TB, k, cx<T>(sr + u, vv + di));
                if (q != k) fdl_store<T>(g, xrow, Xnyq, ch, slot, TB, q, cx<T>(sr - u, vv - di));
            }
        }
    }
    trace_mark(g, 0, 1);
}

// ---------------------------------------------------------------------------------------------
// k_inv_cl: grid = 8 x output channels, cluster (8,1,1), L/8 threads.  Arguments as k_inv (single hop, not sharded).
// ---------------------------------------------------------------------------------------------
// Capped at 152 registers (144 used, no spills): two of these CTAs (128 threads at config 5) then fit on an SM beside a
// resident multiply-accumulate CTA (256 x 96 registers) and all 16 clusters of config 5 are placed at once.  A 192-register
// build (one CTA per SM) left one cluster in 16 waiting for the multiply-accumulate launch to drain, and a build that met
// the cap by spilling ran 75 us beside the tail instead of 35-45 (profiles/r1_c5_cluster_fft.txt).
template <class T>
__global__ void __cluster_dims__(CL_CS, 1, 1) __maxnreg__(152)
k_inv_cl(const Geom g, const SegSets sets, const T *__restrict__ Xnyq, const T *__restrict__ Hnyq,
         T *__restrict__ yout, size_t ld, size_t off, int add_result,
         const T *__restrict__ carry_src, size_t carry_src_ld, T *__restrict__ carry_dst, size_t carry_dst_ld, int add_carry,
         const Cx<T> *__restrict__ tw, int tw_log2)
{
    typedef typename VecOf<T>::type V;
    constexpr int CPV = VecOf<T>::CPV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ T red[40];
    cgc::cluster_group cluster = cgc::this_cluster();
    const uint32_t r = cluster.block_rank();
    const uint32_t ch = blockIdx.x / CL_CS;
    const uint32_t grp = ch / g.outs, o = ch - grp * g.outs;
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    const uint32_t B = g.B, L = B / CL_CS;
    const int m = (int) g.log2n - 1, l = m - 3;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    Cx<T> *raw = reinterpret_cast<Cx<T> *>(smem_raw);               // summed spectrum bins 8 n + r of this rank (read by the mirror rank)
    Cx<T> *s = raw + padded_elems<HB_PADSH>(L);                     // local transform (padded; read by every rank in the cross pass)
    Cx<T> *stw = s + padded_elems<HB_PADSH>(L);
    uint32_t *seg = reinterpret_cast<uint32_t *>(stw + L / 2);      // [set][bin tile][lo, hi]: CTAs whose partial segments cover the tile
    trace_mark(g, 3, 0);

    Cx<T> twr[CL_EPT];
    twiddle_stage_load<T, CL_EPT>(twr, tw, tw_log2, l);
    for (uint32_t i = tid; i < uint32_t(sets.n) * g.n_bt; i += nthr)
    {
        const uint32_t q = i / g.n_bt, bt = i - q * g.n_bt;
        const uint32_t tile = (grp * g.n_ot + ot) * g.n_bt + bt;
        const uint64_t ulo = uint64_t(tile) * sets.s[q].upt, uhi = ulo + sets.s[q].upt - 1;
        seg[2 * i] = (uint32_t) unit_owner(ulo, sets.s[q].U, sets.s[q].G);
        seg[2 * i + 1] = (uint32_t) unit_owner(uhi, sets.s[q].U, sets.s[q].G);
    }
    if (carry_dst)
    {
        // hand the block computed by the previous hop to the caller before this hop's result replaces it: this rank's eighth
        const T *cs = carry_src + size_t(ch) * carry_src_ld;
        T *cd = carry_dst + size_t(ch) * carry_dst_ld;
        const bool vs = pair_aligned(carry_src, carry_src_ld), vd = pair_aligned(carry_dst, carry_dst_ld);
        const uint32_t per = B / (2 * CL_CS);                       // sample pairs per rank = 4 per thread
        Pair<T> cv[CL_EPT / 2], dv[CL_EPT / 2];
#pragma unroll
        for (int e = 0; e < CL_EPT / 2; e++)
        {
            const uint32_t k = r * per + tid + e * nthr;
            cv[e] = ld_pair(cs + 2 * k, vs);
            if (add_carry) dv[e] = ld_pair(cd + 2 * k, vd);
        }
#pragma unroll
        for (int e = 0; e < CL_EPT / 2; e++)
        {
            const uint32_t k = r * per + tid + e * nthr;
            if (add_carry) st_pair(cd + 2 * k, dv[e].a + cv[e].a, dv[e].b + cv[e].b, vd);
            else st_pair(cd + 2 * k, cv[e].a, cv[e].b, vd);
        }
    }
    // Nyquist bin: a real dot product over (in, partition); only the rank that holds bin 0 uses it
    T part = T(0);
    if (r == 0)
    {
        const T *hn = Hnyq + (size_t(grp) * g.outs + o) * g.ins * g.Pcap;
        const T *xn = Xnyq + size_t(grp) * g.ins * g.R;
        for (uint32_t idx = tid; idx < g.upt; idx += nthr)
        {
            const uint32_t in = idx / g.P, p = idx - in * g.P;
            uint32_t sl = g.slot + p;
            if (sl >= g.R) sl -= g.R;
            part += xn[size_t(in) * g.R + sl] * hn[size_t(in) * g.Pcap + p];
        }
    }
    const T nyq = block_sum<T>(part, red);                          // its barriers also publish seg[]
    twiddle_stage_store<T, CL_EPT>(stw, twr, l);

    // bins 8 n + r of this rank: fixed-order sums of the partial segments (set after set, CTA order -- the order of k_inv).
    // Beside a multiply-accumulate launch that saturates HBM every dependent round trip costs microseconds, so the
    // segments of both sets form one list per bin and four of them are in flight per bin and round (16 loads a thread).
    constexpr int SR = 4, HB = CL_EPT / 2;                          // two halves of four bins, four segments per bin in flight
#pragma unroll 1
    for (int half = 0; half < 2; half++)
    {
        V sum[HB];
        uint32_t cnt0[HB], cntt[HB], most = 0;
        uint64_t b0[HB];
        int32_t off1[HB];                                           // segment i >= cnt0 of the list is segment i + off1 of set 1, counted from b0
        const V *__restrict__ S0 = reinterpret_cast<const V *>(sets.s[0].S);
        const V *__restrict__ S1 = reinterpret_cast<const V *>(sets.s[sets.n > 1 ? 1 : 0].S);
#pragma unroll
        for (int e = 0; e < HB; e++)
        {
            const uint32_t vi = (CL_CS * (tid + (half * HB + e) * nthr) + r) / CPV;
            const uint32_t bt = vi / g.TBV, xa = vi - bt * g.TBV;
            const uint64_t tile = (uint64_t(grp) * g.n_ot + ot) * g.n_bt + bt;
            const uint32_t lo0 = seg[2 * bt];
            cnt0[e] = seg[2 * bt + 1] - lo0 + 1;
            b0[e] = (tile + lo0) * g.Q + row * g.TBV + xa;
            cntt[e] = cnt0[e];
            off1[e] = 0;
            if (sets.n > 1)
            {
                const uint32_t lo1 = seg[2 * (g.n_bt + bt)];
                cntt[e] += seg[2 * (g.n_bt + bt) + 1] - lo1 + 1;
                off1[e] = int32_t(lo1) - int32_t(lo0) - int32_t(cnt0[e]);
            }
            most = cntt[e] > most ? cntt[e] : most;
            vzero(sum[e]);
        }
        for (uint32_t i0 = 0; i0 < most; i0 += SR)
        {
            V p[HB][SR];
#pragma unroll
            for (int e = 0; e < HB; e++)
#pragma unroll
                for (int j = 0; j < SR; j++)
                {
                    const uint32_t i = i0 + j;
                    if (i < cnt0[e]) p[e][j] = S0[b0[e] + uint64_t(i) * g.Q];
                    else if (i < cntt[e]) p[e][j] = S1[int64_t(b0[e]) + int64_t(int32_t(i) + off1[e]) * int64_t(g.Q)];
                    else vzero(p[e][j]);
                }
#pragma unroll
            for (int e = 0; e < HB; e++)
#pragma unroll
                for (int j = 0; j < SR; j++) vadd(sum[e], p[e][j]);
        }
#pragma unroll
        for (int e = 0; e < HB; e++)
        {
            const uint32_t n = tid + (half * HB + e) * nthr;
            if constexpr (CPV == 2) raw[sidx<HB_PADSH>(n)] = ((CL_CS * n + r) & 1) ? cx<T>(sum[e].z, sum[e].w) : cx<T>(sum[e].x, sum[e].y);
            else raw[sidx<HB_PADSH>(n)] = cx<T>(sum[e].x, sum[e].y);
        }
    }
    cluster.sync();                                                 // every rank's raw bins are complete
    // inverse split pass: the partner of bin k = 8 n + r is M - k = 8 (L - n - 1) + (8 - r), held by rank 8 - r (r = 0: bin
    // 8 (L - n) of rank 0 itself); only this rank's member of each pair is kept, planes exchanged on the way into the
    // local transform
    {
        const Cx<T> *praw = cluster.map_shared_rank(raw, (CL_CS - r) & (CL_CS - 1));
        Cx<T> a[CL_EPT], b[CL_EPT], w[CL_EPT];
#pragma unroll
        for (int e = 0; e < CL_EPT; e++)
        {
            const uint32_t n = tid + e * nthr;
            a[e] = raw[sidx<HB_PADSH>(n)];                          // written by this thread
            b[e] = praw[sidx<HB_PADSH>(r ? L - 1 - n : ((L - n) & (L - 1)))];
            w[e] = tw_root(tw, tw_log2, CL_CS * n + r, m + 1);
        }
#pragma unroll
        for (int e = 0; e < CL_EPT; e++)
        {
            const uint32_t n = tid + e * nthr, k = CL_CS * n + r;
            Cx<T> z;
            if (k == 0) z = cx<T>(a[e].x + nyq, a[e].x - nyq);
            else
            {
                const T sr = a[e].x + b[e].x, si = a[e].y + b[e].y, dr = a[e].x - b[e].x, di = a[e].y - b[e].y;
                const T wr = -w[e].x, wi = w[e].y;
                const T u = wr * si + wi * dr;
                const T vv = wi * si - wr * dr;
                z = cx<T>(sr + u, vv + di);
            }
            s[sidx<HB_PADSH>(n)] = cx<T>(z.y, z.x);
        }
    }
    __syncthreads();
    block_fft<T, CL_EPT, HB_PADSH>(s, l, stw, l);
    cluster.sync();

    // cross pass: column k2 = r nthr + tid, outputs k1 = 0..3 = the first B samples
    const uint32_t k2 = r * nthr + tid;
    Cx<T> v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = cluster.map_shared_rank(s, q)[sidx<HB_PADSH>(k2)];
    {
        Cx<T> w[8];
#pragma unroll
        for (int q = 1; q < 8; q++) w[q] = tw_root(tw, tw_log2, q * k2, m);
#pragma unroll
        for (int q = 1; q < 8; q++) v[q] = cmul(v[q], w[q]);
    }
    dft8(v);
    const T scale = T(1) / T(size_t(4) << g.log2n);
    T *dst = yout + size_t(ch) * ld + off;
    const bool vd = pair_aligned(yout + off, ld);
    Pair<T> old[4];
    if (add_result)
    {
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) old[k1] = ld_pair(dst + 2 * (k2 + L * k1), vd);
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++)
    {
        T *d = dst + 2 * (k2 + L * k1);
        if (add_result) st_pair(d, old[k1].a + v[k1].y * scale, old[k1].b + v[k1].x * scale, vd);
        else st_pair(d, v[k1].y * scale, v[k1].x * scale, vd);
    }
    cluster.sync();                                                 // no rank leaves while a peer still reads its s
    trace_mark(g, 3, 1);
}

} // namespace hb
