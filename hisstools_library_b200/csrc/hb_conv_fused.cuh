// hb_conv_fused.cuh -- one launch per hop for small engines (PartitionedConvolve, MonoConvolve parts, NToMonoConvolve,
// small Convolver matrices: groups x ins x up to 8 outputs, one bin tile per spectrum).
//
// Such hops are bound by launch latency, not by HBM: the three-kernel hop (k_fwd -> k_cmac -> k_inv) spends ~7 us per
// launch on a few hundred KiB of L2-resident data.  Here one thread-block CLUSTER per (group, output) does the whole
// hop (PartitionedConvolve.cpp:352-377):
//   (with several outputs every output's cluster transforms the inputs for itself -- a few microseconds of redundant
//   arithmetic instead of a grid-wide dependency -- and only output 0's cluster stores the spectra into the delay line)
//   every rank  forward FFT of the inputs it owns (rank = input mod cluster size) into the newest FDL slot, partition 0
//               against that fresh spectrum, then its share of the (input, partition >= 1) products against spectra that
//               are already in the delay line -- all accumulated in registers, one complex bin set per thread;
//   cluster barrier; rank 0 adds the other ranks' partial spectra and Nyquist sums straight out of their shared memory
//               (distributed shared memory), then inverse split, inverse FFT, scale 1/(4N), first B samples out.
// Layouts are the engine's own (hb_conv_kernels.cuh) with OT = 1 and one bin tile, so IR loading and the other
// schedules are unchanged.
//
// Programmatic dependent launch (round 2).  Back-to-back hops of a stream are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and the kernel is ordered in two phases around griddepcontrol.wait:
//   before the wait  what does not depend on the previous hop: twiddles into shared memory, the newest input block into
//                    registers, and the products of partitions p >= 2 -- they meet spectra that are at least two hops old,
//                    written by kernels that had completed before this one could start (a kernel releases its successor only
//                    after its own wait has returned);
//   after the wait   the previous block (saved by the previous hop), the forward FFT, partitions 0 and 1 (p = 1 meets the
//                    spectrum the previous hop wrote), the reduction, the inverse FFT and the store.
// So the launch latency of hop t+1 and most of its L2 round trips run beside the critical part of hop t.  Without the launch
// attribute (or with anything else between two hops in the stream) the wait returns at once and the kernel is what it was.
#pragma once

#include <cooperative_groups.h>

#include "hb_conv_kernels.cuh"

namespace hb
{
namespace cg = cooperative_groups;

struct FusedArgs
{
    uint32_t cs;               // cluster size
    uint32_t tail_items;       // ins * (P - 2): the products of partitions p >= 2 (p = 1 waits for the previous hop)
};

template <class T, int EPT>
__global__ void __launch_bounds__(512) k_hop_fused(const Geom g, const FusedArgs fa,
                                                   const T *__restrict__ prev, size_t prev_ld, const T *__restrict__ newest, size_t new_ld,
                                                   T *__restrict__ save, size_t save_ld,
                                                   const Cx<T> *__restrict__ H, Cx<T> *__restrict__ X, const T *__restrict__ Hnyq, T *__restrict__ Xnyq,
                                                   T *__restrict__ yout, size_t ld, size_t off, int add_result,
                                                   const T *__restrict__ carry_src, size_t carry_src_ld, T *__restrict__ carry_dst, size_t carry_dst_ld, int add_carry,
                                                   const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank(), cs = fa.cs;
    const uint32_t cl = blockIdx.x / cs;                        // one cluster per (group, output)
    const uint32_t grp = cl / g.outs, o = cl - grp * g.outs;
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    const uint32_t tile = grp * g.n_ot + ot;                    // one bin tile: tile = (group, output tile)
    const bool writer = o == 0;                                 // this cluster keeps the delay line
    const uint32_t B = g.B, P = g.P, R = g.R;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);             // FFT work array (padded)
    Cx<T> *stw = s + padded_elems<HB_PADSH>(B);                 // twiddles of this size
    Cx<T> *xch = stw + B;                                       // this rank's partial spectrum, read by rank 0
    __shared__ T red[40];
    __shared__ T nyq_part;
    trace_mark(g, 0, 0);

    // twiddles of this transform size into shared memory (published by the first barrier below)
    {
        Cx<T> twr[EPT];
        twiddle_stage_load<T, EPT>(twr, tw, tw_log2, (int) g.log2n);
        twiddle_stage_store<T, EPT>(stw, twr, (int) g.log2n);
    }
    const Cx<T> *twl = stw;
    const int twl_log2 = (int) g.log2n;

    Cx<T> acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) acc[e] = cx<T>(T(0), T(0));
    T nyq = T(0);

    // ================= before the wait: nothing here depends on the previous hop =================
    // the newest block of the first input this rank transforms (the caller's rows are complete before the launch)
    // (float only: the double instance has no registers to carry values across the transform -- 128 at 512 threads)
    constexpr bool EARLY = sizeof(T) == 4;
    T na[EARLY ? EPT : 1], nbv[EARLY ? EPT : 1];
    if (EARLY && rank < g.ins)
    {
        const T *pn = newest + size_t(grp * g.ins + rank) * new_ld;
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr, j = 2 * k;
            if (k < B && j < B) { na[e] = pn[j]; nbv[e] = pn[j + 1]; }
        }
    }
    trace_mark(g, 1, 0);
    // this rank's share of the (input, partition >= 2) products: spectra that are at least two hops old
    if (P > 2)
    {
        const uint32_t q0 = (uint32_t) ((uint64_t(rank) * fa.tail_items) / cs), q1 = (uint32_t) ((uint64_t(rank + 1) * fa.tail_items) / cs);
        const uint32_t pm2 = P - 2;
        for (uint32_t q = q0; q < q1; q++)
        {
            const uint32_t in = q / pm2, p = 2 + (q - in * pm2);
            const uint32_t ch = grp * g.ins + in;
            uint32_t sl = g.slot + p;
            if (sl >= R) sl -= R;
            const Cx<T> *hp = H + (((size_t(tile) * g.ins + in) * g.Pcap + p) * g.OT + row) * B;
            const Cx<T> *xp = X + (size_t(ch) * R + sl) * B;
            Cx<T> hv[EPT], xv[EPT];
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B) { hv[e] = hp[k]; xv[e] = xp[k]; }
            }
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B)
                {
                    acc[e].x = fma(xv[e].x, hv[e].x, acc[e].x); acc[e].x = fma(-xv[e].y, hv[e].y, acc[e].x);
                    acc[e].y = fma(xv[e].x, hv[e].y, acc[e].y); acc[e].y = fma(xv[e].y, hv[e].x, acc[e].y);
                }
            }
            if (tid == 0) nyq += Xnyq[size_t(ch) * R + sl] * Hnyq[(size_t(cl) * g.ins + in) * g.Pcap + p];
        }
    }

    // ================= the previous hop must be complete from here on; release the next one =================
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (rank == 0 && carry_dst)
    {
        // hand the block computed by the previous hop to the caller (the output-ring read of PartitionedConvolve.cpp:307)
        const T *cs_ = carry_src + size_t(cl) * carry_src_ld;
        T *cd = carry_dst + size_t(cl) * carry_dst_ld;
        for (uint32_t k = tid; k < B; k += nthr) cd[k] = add_carry ? cd[k] + cs_[k] : cs_[k];
    }

    // ---- inputs this rank transforms: forward FFT, newest FDL slot, partitions 0 and 1 ----
    for (uint32_t in = rank; in < g.ins; in += cs)
    {
        const uint32_t ch = grp * g.ins + in;
        const T *pn = newest + size_t(ch) * new_ld, *pp = prev + size_t(ch) * prev_ld;
        T *ps = (save && writer) ? save + size_t(ch) * save_ld : nullptr;
        // partition 1 meets the spectrum the previous hop wrote: fetched now, used after the transform
        Cx<T> x1[EPT], h1[EPT];
        T xn1 = T(0), hn1 = T(0);
        uint32_t sl1 = g.slot + 1;
        if (sl1 >= R) sl1 -= R;
        const Cx<T> *hp1 = H + (((size_t(tile) * g.ins + in) * g.Pcap + 1) * g.OT + row) * B;
        const Cx<T> *xp1 = X + (size_t(ch) * R + sl1) * B;
        if (EARLY && P > 1)
        {
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B) { h1[e] = hp1[k]; x1[e] = xp1[k]; }
            }
        }
        if (P > 1 && tid == 0) { xn1 = Xnyq[size_t(ch) * R + sl1]; hn1 = Hnyq[(size_t(cl) * g.ins + in) * g.Pcap + 1]; }
        __syncthreads();                                        // s is free (previous input's spectrum consumed)
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr, j = 2 * k;
            if (k < B)
            {
                T a, b;
                if (j < B)
                {
                    if (EARLY && in == rank) { a = na[EARLY ? e : 0]; b = nbv[EARLY ? e : 0]; }      // loaded before the wait
                    else { a = pn[j]; b = pn[j + 1]; }
                    if (ps) { ps[j] = a; ps[j + 1] = b; }
                }
                else { a = pp[j - B]; b = pp[j - B + 1]; }
                s[sidx<HB_PADSH>(k)] = cx<T>(a, b);
            }
        }
        __syncthreads();
        block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, twl, twl_log2);
        block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, false, twl, twl_log2);
        __syncthreads();
        Cx<T> *xrow = X + (size_t(ch) * R + g.slot) * B;
        // unit (tile, in, p) holds OT rows of B bins; this output is row `row` of it
        const Cx<T> *h0 = H + ((size_t(tile) * g.ins + in) * g.Pcap * g.OT + row) * B;
        const T *hnq = Hnyq + (size_t(cl) * g.ins + in) * g.Pcap;
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B)
            {
                Cx<T> z = s[sidx<HB_PADSH>(k)];
                if (k == 0)
                {
                    const T xn = z.y;
                    if (writer) Xnyq[size_t(ch) * R + g.slot] = xn;
                    nyq += xn * hnq[0];
                    z.y = T(0);
                }
                if (writer) xrow[k] = z;
                const Cx<T> h = h0[k];
                acc[e].x = fma(z.x, h.x, acc[e].x); acc[e].x = fma(-z.y, h.y, acc[e].x);
                acc[e].y = fma(z.x, h.y, acc[e].y); acc[e].y = fma(z.y, h.x, acc[e].y);
            }
        }
        if (P > 1)
        {
            if (!EARLY)
            {
#pragma unroll
                for (int e = 0; e < EPT; e++)
                {
                    const uint32_t k = tid + e * nthr;
                    if (k < B) { h1[e] = hp1[k]; x1[e] = xp1[k]; }
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B)
                {
                    acc[e].x = fma(x1[e].x, h1[e].x, acc[e].x); acc[e].x = fma(-x1[e].y, h1[e].y, acc[e].x);
                    acc[e].y = fma(x1[e].x, h1[e].y, acc[e].y); acc[e].y = fma(x1[e].y, h1[e].x, acc[e].y);
                }
            }
            if (tid == 0) nyq += xn1 * hn1;
        }
    }

    trace_mark(g, 1, 1);                                        // products done
    // ---- publish the partial spectrum and Nyquist sum; rank 0 reduces over the cluster ----
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B) xch[k] = acc[e];
    }
    const T nyq_sum = block_sum<T>(nyq, red);
    if (tid == 0) nyq_part = nyq_sum;
    cluster.sync();
    trace_mark(g, 2, 0);                                        // first cluster barrier passed
    if (rank == 0)
    {
        T nyq_total = nyq_sum;
        for (uint32_t r = 1; r < cs; r++)
        {
            const Cx<T> *rx = cluster.map_shared_rank(xch, r);
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B) { const Cx<T> v = rx[k]; acc[e].x += v.x; acc[e].y += v.y; }
            }
            nyq_total += *cluster.map_shared_rank(&nyq_part, r);
        }
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B) s[sidx<HB_PADSH>(k)] = k ? acc[e] : cx<T>(acc[e].x, nyq_total);
        }
    }
    cluster.sync();                                             // remote shared memory may go away from here on
    trace_mark(g, 2, 1);                                        // reduction done
    if (rank != 0) { trace_mark(g, 0, 1); return; }

    // ---- inverse real FFT, scale, first B samples (k_inv's tail) ----
    block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, true, twl, twl_log2);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B)
        {
            const Cx<T> z = s[sidx<HB_PADSH>(k)];
            s[sidx<HB_PADSH>(k)] = cx<T>(z.y, z.x);
        }
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, twl, twl_log2);
    trace_mark(g, 3, 0);                                        // inverse transform done
    const T scale = T(1) / T(size_t(4) << g.log2n);
    T *dst = yout + size_t(cl) * ld + off;
#pragma unroll
    for (int e = 0; e < EPT / 2; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B / 2)
        {
            const Cx<T> z = s[sidx<HB_PADSH>(k)];
            if (add_result) { dst[2 * k] += z.y * scale; dst[2 * k + 1] += z.x * scale; }
            else { dst[2 * k] = z.y * scale; dst[2 * k + 1] = z.x * scale; }
        }
    }
    trace_mark(g, 0, 1);
}

} // namespace hb
