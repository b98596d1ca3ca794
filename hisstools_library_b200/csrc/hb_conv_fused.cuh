// hb_conv_fused.cuh -- one launch per hop for small engines (PartitionedConvolve, MonoConvolve parts, NToMonoConvolve,
// small Convolver matrices: groups x ins x up to 8 outputs, one bin tile per spectrum).
//
// Such hops are bound by launch latency, not by HBM: the three-kernel hop (k_fwd -> k_cmac -> k_inv) spends ~7 us per
// launch on a few hundred KiB of L2-resident data.  Here one thread-block CLUSTER per (group, output) does the whole
// hop (PartitionedConvolve.cpp:352-377):
//   (with several outputs every output's cluster transforms the inputs for itself -- a few microseconds of redundant
//   arithmetic instead of a grid-wide dependency -- and only output 0's cluster stores the spectra into the delay line)
//   every rank  forward FFT of the inputs it owns (rank = input mod cluster size) into the newest FDL slot, partition 0
//               against that fresh spectrum, partition 1 of those inputs, and its share of the (input, partition >= 2)
//               products against spectra that are already in the delay line (the rows of several items fetched together:
//               these loads come from L2 and their latency is what a rank spends) -- accumulated in registers;
//   cluster barrier; every rank sums its slice of the bins over all ranks' partial spectra (distributed shared memory, in rank
//               order) and stores it into rank 0's work array; rank 0: inverse split, inverse FFT, scale 1/(4N), first B samples out.
// Layouts are the engine's own (hb_conv_kernels.cuh) with OT = 1 and one bin tile, so IR loading and the other
// schedules are unchanged.
//
// Programmatic dependent launch (round 2).  Back-to-back hops of a stream are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization.  The kernel runs in one of two orders (FusedArgs::chained):
//
//  strict   (any stream; what a caller gets who only promises stream order for its rows)
//           before griddepcontrol.wait: twiddles into shared memory and the products of partitions p >= 2 -- they meet spectra
//           that are at least two hops old, written by kernels that had completed before this one could start (a strict kernel
//           releases its successor only after its own wait has returned; behind a chained hop, which does not wait, these
//           products run after the wait as well: FusedArgs::tail_early);  after it: the caller's rows, the forward FFT,
//           partitions 0 and 1, the reduction, the inverse FFT and the store.
//  chained  (the engine's own stream, or a caller that declares its rows complete when the call is made:
//           hb_conv_set_hop_overlap)  no wait up front.  The hop depends on its predecessors through three counters in
//           device memory (FusedArgs::sync), bumped by every hop that a chained one may follow (FusedArgs::bump), strict or chained:
//             [0] input blocks saved      -> the previous block of this hop's frame
//             [1] spectra stored          -> partition 1 (previous hop's spectrum), partitions >= 2 (two hops back and older)
//             [2] hops finished           -> the delay-line slot this hop overwrites is no longer read (at most `depth` hops
//                                            are in flight), and -- through griddepcontrol.wait just before the store -- hops
//                                            leave their blocks in stream order.
//           It builds its frame and releases its successor at once, transforms, and only then waits for the spectra it
//           multiplies with: the forward FFT of hop t+1 does not wait for anything hop t computes, so consecutive hops run
//           side by side on different SMs and a stream of single-block calls is bound by the launch rate, not by the
//           latency of one hop.  Spectra and saved blocks written by a hop that may still be running are read with
//           ld.global.cg (L2) after an acquire load of the counter.
#pragma once

#include <cooperative_groups.h>

#include "hb_conv_kernels.cuh"

namespace hb
{
namespace cg = cooperative_groups;

struct FusedArgs
{
    uint32_t cs;               // cluster size
    uint32_t tail_items;       // ins * (P - 2): the products of partitions p >= 2
    uint32_t chained;          // 1: consecutive hops overlap (counters); 0: griddepcontrol.wait first
    uint32_t bump;             // 1: this hop counts (a chained hop may follow it); 0: nothing on this stream will look at the counters
    uint32_t tail_early;       // strict order: 1 = the products of partitions >= 2 run before the wait (the hop before this one was strict
                               // too, so the spectra two hops back are complete when this kernel starts); 0 = after it
    uint32_t writers;          // CTAs per hop that bump sync[0] and sync[1]: groups * min(cs, ins)
    uint32_t clusters;         // CTAs per hop that bump sync[2]: groups * outs
    uint32_t depth;            // hops that may be in flight behind the one whose delay-line slot is reused
    unsigned long long n;      // number of this hop since the counters were zeroed (1, 2, ...)
    unsigned long long *sync;  // [0] blocks saved, [1] spectra stored, [2] hops finished
};

// one thread's wait for a counter of the hop chain (the hop that bumps it was launched earlier on this stream and is resident
// or finished, so the wait is short; a wait of half a minute means a broken chain and stops the context instead of hanging the device)
__device__ __forceinline__ void chain_wait(const unsigned long long *p, unsigned long long target)
{
    if (!target) return;
    unsigned long long v, t0 = 0;
    for (uint32_t spins = 0;; spins++)
    {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= target) return;
        if ((spins & 255u) == 255u)
        {
            const unsigned long long t = global_ns();
            if (!t0) t0 = t;
            else if (t - t0 > 30000000000ull) __trap();            // 30 s
        }
    }
}

__device__ __forceinline__ void chain_bump(unsigned long long *p)
{
    __threadfence();
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

__device__ __forceinline__ Cx<float> ld_l2(const Cx<float> *p) { const float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); return cx<float>(v.x, v.y); }
__device__ __forceinline__ Cx<double> ld_l2(const Cx<double> *p) { const double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); return cx<double>(v.x, v.y); }

// MAXT: 256 for spectra of up to 2048 bins (the whole register file of a thread is available), 512 above
template <class T, int EPT, int MAXT>
__global__ void __launch_bounds__(MAXT) k_hop_fused(const Geom g, const FusedArgs fa,
                                                   const T *__restrict__ prev, size_t prev_ld, const T *__restrict__ newest, size_t new_ld,
                                                   T *__restrict__ save, size_t save_ld,
                                                   const Cx<T> *__restrict__ H, Cx<T> *X, const T *__restrict__ Hnyq, T *Xnyq,
                                                   T *yout, size_t ld, size_t off, int add_result,
                                                   const T *carry_src, size_t carry_src_ld, T *carry_dst, size_t carry_dst_ld, int add_carry,
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
    Cx<T> *xch = stw + B;                                       // this rank's partial spectrum, read by every rank of the cluster
    __shared__ T red[40];
    __shared__ T nyq_part;
    const bool chained = fa.chained != 0;
    const unsigned long long n1 = fa.n >= 1 ? (fa.n - 1) * fa.writers : 0;      // what the previous hop leaves in sync[0] and sync[1]
    const unsigned long long n2 = fa.n >= 2 ? (fa.n - 2) * fa.writers : 0;
    trace_mark(g, 0, 0);

    // Two running sums, so that the samples do not depend on the order the kernel runs in: `acc` / `nyq` take partitions 0 and 1
    // (in input order), the products of partitions >= 2 are summed on their own and parked in this rank's exchange row.
    Cx<T> acc[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) acc[e] = cx<T>(T(0), T(0));
    T nyq = T(0), nyq_tail = T(0);

    // twiddles of this transform size: fetched now, put into shared memory further down (published by the first barrier after that)
    Cx<T> twr[EPT];
    twiddle_stage_load<T, EPT>(twr, tw, tw_log2, (int) g.log2n);
    const Cx<T> *twl = stw;
    const int twl_log2 = (int) g.log2n;

    // this rank's share of the (input, partition >= 2) products: spectra that are at least two hops old
    // (these loads come from L2 and a product needs a few instructions: what a rank spends here is the latency of its loads, so
    // the rows of U items are fetched together -- as many as the registers of this instance hold)
    constexpr int U = MAXT == 256 ? (sizeof(T) == 4 ? 4 : 2) : 1;
    auto tail_products = [&]()
    {
        Cx<T> t[EPT];
#pragma unroll
        for (int e = 0; e < EPT; e++) t[e] = cx<T>(T(0), T(0));
        const uint32_t q0 = (uint32_t) ((uint64_t(rank) * fa.tail_items) / cs), q1 = (uint32_t) ((uint64_t(rank + 1) * fa.tail_items) / cs);
        const uint32_t pm2 = P > 2 ? P - 2 : 1;
        for (uint32_t qb = q0; qb < q1; qb += U)
        {
            Cx<T> hv[U][EPT], xv[U][EPT];
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                const uint32_t q = qb + u;
                if (q < q1)
                {
                    const uint32_t in = q / pm2, p = 2 + (q - in * pm2);
                    const uint32_t ch = grp * g.ins + in;
                    uint32_t sl = g.slot + p;
                    if (sl >= R) sl -= R;
                    const Cx<T> *hp = H + (((size_t(tile) * g.ins + in) * g.Pcap + p) * g.OT + row) * B;
                    const Cx<T> *xp = X + (size_t(ch) * R + sl) * B;
#pragma unroll
                    for (int e = 0; e < EPT; e++)
                    {
                        const uint32_t k = tid + e * nthr;
                        if (k < B) { hv[u][e] = hp[k]; xv[u][e] = ld_l2(xp + k); }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++)
            {
                if (qb + u < q1)
                {
#pragma unroll
                    for (int e = 0; e < EPT; e++)
                    {
                        const uint32_t k = tid + e * nthr;
                        if (k < B)
                        {
                            t[e].x = fma(xv[u][e].x, hv[u][e].x, t[e].x); t[e].x = fma(-xv[u][e].y, hv[u][e].y, t[e].x);
                            t[e].y = fma(xv[u][e].x, hv[u][e].y, t[e].y); t[e].y = fma(xv[u][e].y, hv[u][e].x, t[e].y);
                        }
                    }
                }
            }
        }
        // the Nyquist products of these items, one item per thread (summed over the block further down)
        for (uint32_t q = q0 + tid; q < q1; q += nthr)
        {
            const uint32_t in = q / pm2, p = 2 + (q - in * pm2);
            const uint32_t ch = grp * g.ins + in;
            uint32_t sl = g.slot + p;
            if (sl >= R) sl -= R;
            nyq_tail += __ldcg(Xnyq + size_t(ch) * R + sl) * Hnyq[(size_t(cl) * g.ins + in) * g.Pcap + p];
        }
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B) xch[k] = t[e];                       // each thread parks and later fetches its own bins
        }
    };

    // hand the block computed by the previous hop to the caller (the output-ring read of PartitionedConvolve.cpp:307): fetched as
    // early as the order allows, stored once the loads of the frame are under way
    Cx<T> cv[EPT / 2];
    auto carry_load = [&]()
    {
        if (!carry_dst) return;
        const T *cs_ = carry_src + size_t(cl) * carry_src_ld;
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B / 2) cv[e] = cx<T>(__ldcg(cs_ + 2 * k), __ldcg(cs_ + 2 * k + 1));
        }
    };
    auto carry_store = [&]()
    {
        if (!carry_dst) return;
        T *cd = carry_dst + size_t(cl) * carry_dst_ld;
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B / 2)
            {
                if (add_carry) { cd[2 * k] += cv[e].x; cd[2 * k + 1] += cv[e].y; }
                else { cd[2 * k] = cv[e].x; cd[2 * k + 1] = cv[e].y; }
            }
        }
    };

    if (!chained)
    {
        twiddle_stage_store<T, EPT>(stw, twr, (int) g.log2n);
        if (fa.tail_early) tail_products();
        trace_mark(g, 1, 0);
        // ================= the previous hop must be complete from here on; release the next one =================
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // (a hop that a chained one may follow releases it only once the saved block is in shared memory, as a chained hop does:
        // the successor saves its own block over it)
        if (!fa.bump || rank >= g.ins) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (!fa.tail_early) tail_products();
        if (rank == 0) carry_load();
    }
    else
    {
        if (tid == 0) chain_wait(fa.sync + 0, n1);                                                 // the previous block is saved
        if (tid == 32 && fa.n > fa.depth + 1) chain_wait(fa.sync + 2, (fa.n - 1 - fa.depth) * fa.clusters);   // the slot written below is free
        if (rank >= g.ins) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");         // no rows of the caller's to read
        twiddle_stage_store<T, EPT>(stw, twr, (int) g.log2n);
    }
    trace_mark(g, 4, 0);

    // ---- inputs this rank transforms: forward FFT, newest FDL slot, partition 0 ----
    constexpr bool KEEP = sizeof(T) == 4;                       // float: the rows of partitions 0 and 1 are fetched ahead of the transform
    Cx<T> h0v[KEEP ? EPT : 1], h1v[KEEP ? EPT : 1], x1v[KEEP ? EPT : 1];   // (x1v: strict order only -- the previous hop's spectrum is complete)
    uint32_t sl1 = g.slot + 1;
    if (sl1 >= R) sl1 -= R;
    for (uint32_t in = rank; in < g.ins; in += cs)
    {
        const uint32_t ch = grp * g.ins + in;
        const T *pn = newest + size_t(ch) * new_ld, *pp = prev + size_t(ch) * prev_ld;
        T *ps = (save && writer) ? save + size_t(ch) * save_ld : nullptr;
        const bool last_in = in + cs >= g.ins;
        // unit (tile, in, p) holds OT rows of B bins; this output is row `row` of it
        const Cx<T> *h0 = H + ((size_t(tile) * g.ins + in) * g.Pcap * g.OT + row) * B;
        const T *hnq = Hnyq + (size_t(cl) * g.ins + in) * g.Pcap;
        __syncthreads();                                        // s is free (previous input's spectrum consumed); the chain waits are over
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr, j = 2 * k;
            if (k < B)
            {
                T a, b;
                if (j < B)
                {
                    a = pn[j]; b = pn[j + 1];
                    if (ps) { ps[j] = a; ps[j + 1] = b; }
                }
                else { a = __ldcg(pp + (j - B)); b = __ldcg(pp + (j - B + 1)); }
                s[sidx<HB_PADSH>(k)] = cx<T>(a, b);
            }
        }
        __syncthreads();
        if (last_in)
        {
            // every row this CTA reads from the previous hop's save area is in shared memory and its own blocks are saved
            if (tid == 0 && writer && fa.bump) chain_bump(fa.sync + 0);
            if (fa.bump) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            if (!chained && rank == 0) carry_store();
        }
        if (KEEP)
        {
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B)
                {
                    h0v[KEEP ? e : 0] = h0[k];
                    if (P > 1 && last_in)
                    {
                        h1v[KEEP ? e : 0] = h0[size_t(g.OT) * B + k];
                        if (!chained) x1v[KEEP ? e : 0] = ld_l2(X + (size_t(ch) * R + sl1) * B + k);
                    }
                }
            }
        }
        block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, twl, twl_log2);
        block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, false, twl, twl_log2);
        __syncthreads();
        if (last_in) trace_mark(g, 4, 1);
        Cx<T> *xrow = X + (size_t(ch) * R + g.slot) * B;
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
                const Cx<T> h = KEEP ? h0v[KEEP ? e : 0] : h0[k];
                acc[e].x = fma(z.x, h.x, acc[e].x); acc[e].x = fma(-z.y, h.y, acc[e].x);
                acc[e].y = fma(z.x, h.y, acc[e].y); acc[e].y = fma(z.y, h.x, acc[e].y);
            }
        }
    }
    if (writer && rank < g.ins && fa.bump)
    {
        __syncthreads();                                        // every thread's spectrum stores precede the bump
        if (tid == 0) chain_bump(fa.sync + 1);
    }

    if (chained)
    {
        trace_mark(g, 1, 0);
        if (tid == 0) chain_wait(fa.sync + 1, n2);              // spectra up to two hops back are stored
        __syncthreads();
        tail_products();
        if (P > 1)
        {
            if (tid == 0) chain_wait(fa.sync + 1, n1);          // the previous hop's spectrum is stored
            __syncthreads();
        }
    }

    // ---- partition 1 meets the spectrum the previous hop wrote ----
    if (P > 1)
    {
        for (uint32_t in = rank; in < g.ins; in += cs)
        {
            const uint32_t ch = grp * g.ins + in;
            const bool kept = KEEP && in + cs >= g.ins;          // the last input's row was fetched ahead
            const Cx<T> *hp1 = H + (((size_t(tile) * g.ins + in) * g.Pcap + 1) * g.OT + row) * B;
            const Cx<T> *xp1 = X + (size_t(ch) * R + sl1) * B;
#pragma unroll
            for (int e = 0; e < EPT; e++)
            {
                const uint32_t k = tid + e * nthr;
                if (k < B)
                {
                    const Cx<T> h = kept ? h1v[KEEP ? e : 0] : hp1[k];
                    const Cx<T> x = (kept && !chained) ? x1v[KEEP ? e : 0] : ld_l2(xp1 + k);
                    acc[e].x = fma(x.x, h.x, acc[e].x); acc[e].x = fma(-x.y, h.y, acc[e].x);
                    acc[e].y = fma(x.x, h.y, acc[e].y); acc[e].y = fma(x.y, h.x, acc[e].y);
                }
            }
            if (tid == 0) nyq += __ldcg(Xnyq + size_t(ch) * R + sl1) * Hnyq[(size_t(cl) * g.ins + in) * g.Pcap + 1];
        }
    }

    trace_mark(g, 1, 1);                                        // products done
    // ---- publish the partial spectrum and Nyquist sum; rank 0 reduces over the cluster ----
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B) { const Cx<T> t = xch[k]; acc[e] = cx<T>(t.x + acc[e].x, t.y + acc[e].y); xch[k] = acc[e]; }
    }
    const T nyq_sum = block_sum<T>(nyq_tail + nyq, red);
    if (cs == 1)
    {
        // a single CTA: the sums are complete in registers (block_sum's barriers separate the last reads of s from these writes)
#pragma unroll
        for (int e = 0; e < EPT; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B) s[sidx<HB_PADSH>(k)] = k ? acc[e] : cx<T>(acc[e].x, nyq_sum);
        }
        __syncthreads();
        trace_mark(g, 2, 0);
    }
    else
    {
    if (tid == 0) nyq_part = nyq_sum;
    cluster.sync();
    trace_mark(g, 2, 0);                                        // first cluster barrier passed
    // Every rank sums its slice of the bins over all ranks (in rank order, whoever does it) and stores it into rank 0's work array:
    // the reads through distributed shared memory are spread over the SMs of the cluster instead of queueing on one.
    {
        __shared__ T nq[16];
        T nyq_total = T(0);
        if (rank == 0)
        {
            if (tid < cs) nq[tid] = *cluster.map_shared_rank(&nyq_part, tid);
            __syncthreads();
            for (uint32_t r = 0; r < cs; r++) nyq_total += nq[r];
        }
        const uint32_t lo = (uint32_t) (uint64_t(rank) * B / cs), hi = (uint32_t) (uint64_t(rank + 1) * B / cs);
        Cx<T> *s0 = cluster.map_shared_rank(s, 0);
        constexpr int RB = sizeof(T) == 4 ? 8 : 4;                  // reads in flight per bin
        for (uint32_t k = lo + tid; k < hi; k += nthr)
        {
            Cx<T> a = cx<T>(T(0), T(0));
            for (uint32_t r0 = 0; r0 < cs; r0 += RB)
            {
                Cx<T> v[RB];
#pragma unroll
                for (int j = 0; j < RB; j++)
                    if (r0 + j < cs) v[j] = cluster.map_shared_rank(xch, r0 + j)[k];
#pragma unroll
                for (int j = 0; j < RB; j++)
                    if (r0 + j < cs) { a.x += v[j].x; a.y += v[j].y; }
            }
            if (k == 0) a.y = nyq_total;
            s0[sidx<HB_PADSH>(k)] = a;
        }
    }
    cluster.sync();                                             // remote shared memory may go away from here on
    }
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

    // chained: the previous hop (and with it every hop before) has finished -- its block may be handed on, the staging row it
    // read may be overwritten, and blocks reach the caller's rows in stream order
    if (chained)
    {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        carry_load();
        carry_store();
    }
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
    if (fa.bump)
    {
        __syncthreads();
        if (tid == 0) chain_bump(fa.sync + 2);
    }
    trace_mark(g, 0, 1);
}

} // namespace hb
