// hb_conv_kernels.cuh -- device side of the uniform partitioned convolution engine.
//
// What the reference does per hop (PartitionedConvolve.cpp:352-377; NToMonoConvolve.cpp:39-42 for the
// sum over inputs) becomes three kernels here:
//
//   k_fwd   one CTA per input channel : frame [newest B | previous B] -> real FFT -> newest slot of
//                                       the frequency-domain delay line (FDL)
//   k_cmac  persistent, one CTA per SM: Y[o] = sum_i sum_p H[o][i][p] (.) X[i][t-p]; every CTA streams
//                                       one contiguous range of IR "units" from HBM (stream-K)
//   k_inv   one CTA per output channel: reduce the stream-K partials in a fixed order, real inverse
//                                       FFT, scale by 1/(4N), keep the first B samples
//
// Data layout in HBM (T = float/double, V = one 16-byte vector = 2 float bins or 1 double bin,
// complex values interleaved re,im):
//
//   IR spectra  H[tile][in][partition < Pcap][row < OT][TBV vectors]       tile = (group, out-tile, bin-tile)
//               one "unit" = the OT x TBV block of one (tile, in, partition) = Q vectors (<= 32 KiB),
//               i.e. exactly what one CTA consumes per step, and consecutive steps are consecutive
//               in memory -> each CTA reads one long contiguous stream (1-D bulk TMA copies).
//   FDL         X[group][in][bin-tile][slot < P][TBV vectors]               ring over slots, newest at `slot`
//   bin 0 holds (DC, 0); the Nyquist values the reference packs into imagp[0]
//   (PartitionedConvolve.cpp:398-406) live in two small real side arrays Hnyq[group][out][in][Pcap]
//   and Xnyq[group][in][P], reduced by k_inv, so the hot loop is a pure complex multiply-accumulate.
//   partials    S[cta + tile][row][TBV] : a CTA writes one segment per tile its unit range touches.
#pragma once

#include "hb_fft_block.cuh"

namespace hb
{

template <class T> struct VecOf;
template <> struct VecOf<float>  { typedef float4 type;  static constexpr int CPV = 2; };
template <> struct VecOf<double> { typedef double2 type; static constexpr int CPV = 1; };

struct Geom
{
    uint32_t groups, ins, outs;
    uint32_t log2n;        // log2 of the FFT size N
    uint32_t B;            // hop = N/2 = complex bins per spectrum
    uint32_t P;            // partitions in use
    uint32_t R;            // FDL ring length (slots) >= P: P + the hops a multi-hop launch looks ahead
    uint32_t Pcap;         // partition stride of the IR layout (capacity at this FFT size)
    uint32_t OT, n_ot;     // output rows per tile, output tiles
    uint32_t TBV, n_bt;    // vectors along bins per tile, bin tiles
    uint32_t TX, TY;       // thread grid inside a CTA (bins x rows); TX*XA = TBV, TY*OB = OT
    uint32_t XA, OB;
    uint32_t Q;            // vectors per unit = OT*TBV
    uint32_t upt;          // units per tile = ins*P
    uint32_t tiles;        // groups*n_ot*n_bt
    uint32_t G;            // CTAs of the multiply-accumulate kernel
    uint32_t slot;         // FDL slot holding the newest spectrum
    uint32_t hop;          // hop counter (trace only)
    uint64_t U;            // units in total = tiles*upt
    unsigned long long *trace;   // optional timeline buffer (hb_conv_set_trace), nullptr = off
};

// Timeline tracing (debug): thread 0 of every CTA stamps %globaltimer at entry and exit into
// trace[(((hop % TRACE_HOPS) * TRACE_KINDS + kind) * 2 + exit) * TRACE_CTAS + cta]; kind 0 forward FFT, 1 head,
// 2 tail / whole multiply-accumulate, 3 inverse FFT, 4 owner-side sum of the multi-GPU exchange.  Shows which kernels really share the machine in the overlapped schedule.
constexpr uint32_t TRACE_HOPS = 16, TRACE_CTAS = 256, TRACE_KINDS = 5;
__device__ __forceinline__ void trace_mark(unsigned long long *trace, uint32_t hop, uint32_t kind, uint32_t is_exit)
{
    if (trace && threadIdx.x == 0 && blockIdx.x < TRACE_CTAS)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace[(((hop % TRACE_HOPS) * TRACE_KINDS + kind) * 2 + is_exit) * TRACE_CTAS + blockIdx.x] = t;
    }
}
__device__ __forceinline__ void trace_mark(const Geom &g, uint32_t kind, uint32_t is_exit) { trace_mark(g.trace, g.hop, kind, is_exit); }

// One launch of the multiply-accumulate kernel covers the partitions [p0, p0 + pc) of every (tile, input):
// the whole IR (p0 = 0, pc = P) in the serial schedule; in the overlapped schedule the newest spectrum
// against partition 0 ("head", on the critical path of a hop) and partitions 1..P-1 against the spectra that
// are already in the delay line ("tail", computed one hop ahead, beside the FFT kernels).
struct Range
{
    uint32_t p0, pc;       // first partition, partition count
    uint32_t upt;          // units per tile = ins*pc
    uint32_t G;            // CTAs of this launch
    uint32_t slot;         // FDL slot of the spectrum that meets partition 0
    uint32_t kind;         // trace only: 1 head, 2 tail / whole
    uint32_t pin_p;        // L2 residency (k_cmac_tma): IR units of partitions p < pin_p are loaded evict_last, the others evict_first
    uint32_t pin_s;        // FDL tiles of ring slots s < pin_s are loaded evict_last, the others evict_first
    uint64_t U;            // units of this launch = tiles*upt
};

// ---------------------------------------------------------------------------------------------
// vector complex multiply-accumulate
// ---------------------------------------------------------------------------------------------
// float: two packed fma.rn.f32x2 (SASS FFMA2) per complex bin -- acc(re, im) += xr * (hr, hi); acc(re, im) += (-hi, hr) * xi.  The
// broadcast of xr / xi, the swap of (hr, hi) and the negated half are operand modifiers of FFMA2 (R.F32, .LO_HI, .NP), so a bin costs
// two instructions of the FMA pipe instead of four (a three-register FFMA issues at half rate on this part; hb_conv_mh.cu).
__device__ __forceinline__ unsigned long long pk_f32x2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void cmac(float4 &a, const float4 x, const float4 h)
{
    unsigned long long a0 = pk_f32x2(a.x, a.y), a1 = pk_f32x2(a.z, a.w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(pk_f32x2(x.x, x.x)), "l"(pk_f32x2(h.x, h.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(pk_f32x2(x.z, x.z)), "l"(pk_f32x2(h.z, h.w)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(pk_f32x2(-h.y, h.x)), "l"(pk_f32x2(x.y, x.y)));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(pk_f32x2(-h.w, h.z)), "l"(pk_f32x2(x.w, x.w)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(a0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a.z), "=f"(a.w) : "l"(a1));
}
__device__ __forceinline__ void cmac(double2 &a, const double2 x, const double2 h)
{
    a.x = fma(x.x, h.x, a.x); a.x = fma(-x.y, h.y, a.x);
    a.y = fma(x.x, h.y, a.y); a.y = fma(x.y, h.x, a.y);
}
__device__ __forceinline__ void vzero(float4 &a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(double2 &a) { a = make_double2(0.0, 0.0); }
__device__ __forceinline__ void vadd(float4 &a, const float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vadd(double2 &a, const double2 b) { a.x += b.x; a.y += b.y; }

// streaming (read-once) 16-byte load that does not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a lost transaction traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---------------------------------------------------------------------------------------------
// unit cursor: walks units u = ((tile*ins + in)*P + p) in order and yields the stream addresses
// ---------------------------------------------------------------------------------------------
struct Cursor
{
    uint32_t tile, in, p;
    __device__ __forceinline__ void seek(const Range &r, uint64_t u)
    {
        tile = (uint32_t) (u / r.upt);
        uint32_t rem = (uint32_t) (u - uint64_t(tile) * r.upt);
        in = rem / r.pc;
        p = r.p0 + (rem - in * r.pc);
    }
    // returns true when the step that was just left was the last unit of its tile
    __device__ __forceinline__ bool advance(const Geom &g, const Range &r)
    {
        if (++p < r.p0 + r.pc) return false;
        p = r.p0;
        if (++in < g.ins) return false;
        in = 0;
        tile++;
        return true;
    }
    // vector offset of this unit in the IR array
    __device__ __forceinline__ uint64_t h_off(const Geom &g) const
    {
        return ((uint64_t(tile) * g.ins + in) * g.Pcap + p) * g.Q;
    }
    // ring slot of the matching FDL tile
    __device__ __forceinline__ uint32_t x_slot(const Geom &g, const Range &r) const
    {
        uint32_t s = r.slot + p;
        if (s >= g.R) s -= g.R;
        return s;
    }
    // vector offset of the matching FDL tile
    __device__ __forceinline__ uint64_t x_off(const Geom &g, const Range &r) const
    {
        uint32_t bt = tile % g.n_bt;
        uint32_t grp = tile / (g.n_bt * g.n_ot);
        uint32_t s = r.slot + p;
        if (s >= g.R) s -= g.R;
        return (((uint64_t(grp) * g.ins + in) * g.n_bt + bt) * g.R + s) * g.TBV;
    }
};

// ---------------------------------------------------------------------------------------------
// k_cmac, variant "tma": IR units and FDL tiles are staged in a shared-memory ring by bulk TMA copies
// issued by one thread; all threads consume from shared memory.  One barrier per step.
// dynamic shared memory: nstages * (Q + TBV) vectors, then nstages mbarriers.
// ---------------------------------------------------------------------------------------------
// Register cap: in the overlapped schedule an FFT CTA (512 threads x 64 registers = 32 K) has to be placed on an SM
// beside a resident multiply-accumulate CTA.  At the 111 registers ptxas picks on its own (256 x 112 = 28 K) the
// block scheduler does not co-schedule the two although 60 K < 64 K; at <= 96 (24 K) it does (measured with
// hb_conv_set_trace, profiles/r1_overlap_trace.txt).  No instance spills at 96 and the streaming rate is unchanged.
#ifndef HB_CMAC_MAXREG
#define HB_CMAC_MAXREG 96
#endif
#define HB_CMAC_BOUNDS __maxnreg__(HB_CMAC_MAXREG)
template <class T, int XA, int OB>
__global__ void HB_CMAC_BOUNDS k_cmac_tma(const Geom g, const Range rg, const typename VecOf<T>::type *__restrict__ H,
                                                     const typename VecOf<T>::type *__restrict__ X,
                                                     typename VecOf<T>::type *__restrict__ S, const int nstages)
{
    typedef typename VecOf<T>::type V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_vecs = g.Q + g.TBV;
    V *ring = reinterpret_cast<V *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(nstages) * stage_vecs * sizeof(V));

    const uint32_t tid = threadIdx.x;
    const uint32_t tx = tid % g.TX, ty = tid / g.TX;
    const bool active = ty < g.TY;
    trace_mark(g, rg.kind, 0);

    const uint64_t u0 = unit_begin(blockIdx.x, rg.U, rg.G), u1 = unit_begin(blockIdx.x + 1, rg.U, rg.G);
    const uint32_t n = (uint32_t) (u1 - u0);

    if (tid == 0)
    {
        for (int s = 0; s < nstages; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    Cursor prod;
    uint64_t pol_first = 0, pol_last = 0;
    uint32_t issued = 0;
    const uint32_t h_bytes = g.Q * (uint32_t) sizeof(V), x_bytes = g.TBV * (uint32_t) sizeof(V);
    if (tid == 0)
    {
        pol_first = l2_policy_evict_first();
        pol_last = l2_policy_evict_last();
        prod.seek(rg, u0);
        for (; issued < n && issued + 1 < (uint32_t) nstages; issued++)
        {
            V *dst = ring + size_t(issued) * stage_vecs;
            mbar_expect_tx(&full[issued], h_bytes + x_bytes);
            bulk_g2s(dst, H + prod.h_off(g), h_bytes, &full[issued], prod.p < rg.pin_p ? pol_last : pol_first);
            bulk_g2s(dst + g.Q, X + prod.x_off(g, rg), x_bytes, &full[issued], prod.x_slot(g, rg) < rg.pin_s ? pol_last : pol_first);
            prod.advance(g, rg);
        }
    }

    Cursor cons;
    cons.seek(rg, u0);
    V acc[XA * OB];
#pragma unroll
    for (int r = 0; r < XA * OB; r++) vzero(acc[r]);

    uint32_t stage = 0, parity = 0;
    for (uint32_t k = 0; k < n; k++)
    {
        if (tid == 0 && issued < n)
        {
            // stage (k + nstages - 1) % nstages was drained in step k-1 (barrier at the end of that step)
            uint32_t ps = stage ? stage - 1 : nstages - 1;
            V *dst = ring + size_t(ps) * stage_vecs;
            mbar_expect_tx(&full[ps], h_bytes + x_bytes);
            bulk_g2s(dst, H + prod.h_off(g), h_bytes, &full[ps], prod.p < rg.pin_p ? pol_last : pol_first);
            bulk_g2s(dst + g.Q, X + prod.x_off(g, rg), x_bytes, &full[ps], prod.x_slot(g, rg) < rg.pin_s ? pol_last : pol_first);
            prod.advance(g, rg);
            issued++;
        }
        mbar_wait(&full[stage], parity);
        if (active)
        {
            const V *hs = ring + size_t(stage) * stage_vecs;
            const V *xs = hs + g.Q;
            V xv[XA];
#pragma unroll
            for (int a = 0; a < XA; a++) xv[a] = xs[tx + g.TX * a];
#pragma unroll
            for (int b = 0; b < OB; b++)
#pragma unroll
                for (int a = 0; a < XA; a++) cmac(acc[b * XA + a], xv[a], hs[(ty + g.TY * b) * g.TBV + tx + g.TX * a]);
        }
        const uint32_t tile_done = cons.tile;
        const bool last = cons.advance(g, rg) || (k + 1 == n);
        if (last && active)
        {
            V *seg = S + (uint64_t(blockIdx.x) + tile_done) * g.Q;
#pragma unroll
            for (int b = 0; b < OB; b++)
#pragma unroll
                for (int a = 0; a < XA; a++)
                {
                    seg[(ty + g.TY * b) * g.TBV + tx + g.TX * a] = acc[b * XA + a];
                    vzero(acc[b * XA + a]);
                }
        }
        if (++stage == (uint32_t) nstages) { stage = 0; parity ^= 1; }
        __syncthreads();
    }
    trace_mark(g, rg.kind, 1);
}

// ---------------------------------------------------------------------------------------------
// k_cmac_tma_mh: the TMA variant for NH consecutive hops in ONE pass over the impulse-response spectra ("multi-hop
// reuse").  A call that brings several hops at once (numSamples >= 2B) has all their input spectra in the delay line
// before any output is needed, so every IR unit is streamed from HBM once and multiplied into NH accumulator sets:
// hop j (0 = oldest) of partition p meets slot rg.slot - j + p.  The kernel stays HBM-bound up to NH = 4 (8 x NH x 8
// FMAs per thread and unit against 1290 cycles of unit transfer time), i.e. NH hops cost about one.
// shared memory: nstages * (Q + NH * TBV) vectors, then nstages mbarriers.  Partials: S[hop][cta + tile][row][TBV].
// ---------------------------------------------------------------------------------------------
template <class T, int XA, int OB, int NH>
__global__ void __launch_bounds__(288, 1) k_cmac_tma_mh(const Geom g, const Range rg, const typename VecOf<T>::type *__restrict__ H,
                                                        const typename VecOf<T>::type *__restrict__ X,
                                                        typename VecOf<T>::type *__restrict__ S, const int nstages, const uint64_t set_stride)
{
    typedef typename VecOf<T>::type V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t stage_vecs = g.Q + NH * g.TBV;
    V *ring = reinterpret_cast<V *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(nstages) * stage_vecs * sizeof(V));

    // 8 consumer warps (the thread grid of the single-hop kernel) + 1 producer warp whose lane 0 issues the copies, so
    // that the address arithmetic of the NH + 1 copies of a stage runs beside the arithmetic instead of in front of it
    const uint32_t tid = threadIdx.x;
    const bool producer = tid >= 256;
    const uint32_t tx = tid % g.TX, ty = tid / g.TX;
    const bool active = !producer && ty < g.TY;
    trace_mark(g, rg.kind, 0);

    const uint64_t u0 = unit_begin(blockIdx.x, rg.U, rg.G), u1 = unit_begin(blockIdx.x + 1, rg.U, rg.G);
    const uint32_t n = (uint32_t) (u1 - u0);

    if (tid == 0)
    {
        for (int s = 0; s < nstages; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    Cursor prod;
    uint64_t pol_h = 0, pol_x = 0;
    uint32_t issued = 0;
    const uint32_t h_bytes = g.Q * (uint32_t) sizeof(V), x_bytes = g.TBV * (uint32_t) sizeof(V);
    // one stage = the IR unit and the FDL tiles of the NH hops that meet it.  Hop j reads slot s - j (mod R): the NH tiles
    // are adjacent in memory (ascending from hop NH - 1 to hop 0) unless the ring wraps inside them, so they normally
    // arrive as one copy; shared-memory order is therefore hop NH - 1 first.
    auto issue = [&](uint32_t st)
    {
        V *dst = ring + size_t(st) * stage_vecs;
        mbar_expect_tx(&full[st], h_bytes + NH * x_bytes);
        bulk_g2s(dst, H + prod.h_off(g), h_bytes, &full[st], pol_h);
        const uint64_t x0 = prod.x_off(g, rg);                       // hop 0: slot rg.slot + p
        uint32_t s = rg.slot + prod.p;
        if (s >= g.R) s -= g.R;
        if (s >= uint32_t(NH - 1))
            bulk_g2s(dst + g.Q, X + x0 - uint64_t(NH - 1) * g.TBV, NH * x_bytes, &full[st], pol_x);
        else
        {
#pragma unroll
            for (int j = 0; j < NH; j++)
            {
                const uint32_t sj = s >= (uint32_t) j ? s - j : s + g.R - j;
                bulk_g2s(dst + g.Q + (NH - 1 - j) * g.TBV, X + x0 + (int64_t(sj) - int64_t(s)) * g.TBV, x_bytes, &full[st], pol_x);
            }
        }
        prod.advance(g, rg);
        issued++;
    };
    if (tid == 256)
    {
        pol_h = l2_policy_evict_first();
        pol_x = rg.pin_s ? l2_policy_evict_last() : l2_policy_evict_first();     // delay line kept in L2 only while it fits (plan_geometry)
        prod.seek(rg, u0);
        while (issued < n && issued + 1 < (uint32_t) nstages) issue(issued);
    }

    Cursor cons;
    cons.seek(rg, u0);
    V acc[NH][XA * OB];
#pragma unroll
    for (int j = 0; j < NH; j++)
#pragma unroll
        for (int r = 0; r < XA * OB; r++) vzero(acc[j][r]);

    uint32_t stage = 0, parity = 0;
    for (uint32_t k = 0; k < n; k++)
    {
        if (producer)
        {
            // stage (k + nstages - 1) % nstages was drained in step k-1 (barrier at the end of that step)
            if (tid == 256 && issued < n) issue(stage ? stage - 1 : nstages - 1);
        }
        else
        {
            mbar_wait(&full[stage], parity);
            if (active)
            {
                const V *hs = ring + size_t(stage) * stage_vecs;
                const V *xs = hs + g.Q;
                V xv[NH][XA];
#pragma unroll
                for (int j = 0; j < NH; j++)
#pragma unroll
                    for (int a = 0; a < XA; a++) xv[j][a] = xs[(NH - 1 - j) * g.TBV + tx + g.TX * a];
#pragma unroll
                for (int b = 0; b < OB; b++)
#pragma unroll
                    for (int a = 0; a < XA; a++)
                    {
                        const V h = hs[(ty + g.TY * b) * g.TBV + tx + g.TX * a];
#pragma unroll
                        for (int j = 0; j < NH; j++) cmac(acc[j][b * XA + a], xv[j][a], h);
                    }
            }
            const uint32_t tile_done = cons.tile;
            const bool last = cons.advance(g, rg) || (k + 1 == n);
            if (last && active)
            {
#pragma unroll
                for (int j = 0; j < NH; j++)
                {
                    V *seg = S + uint64_t(j) * set_stride + (uint64_t(blockIdx.x) + tile_done) * g.Q;
#pragma unroll
                    for (int b = 0; b < OB; b++)
#pragma unroll
                        for (int a = 0; a < XA; a++)
                        {
                            seg[(ty + g.TY * b) * g.TBV + tx + g.TX * a] = acc[j][b * XA + a];
                            vzero(acc[j][b * XA + a]);
                        }
                }
            }
        }
        if (++stage == (uint32_t) nstages) { stage = 0; parity ^= 1; }
        __syncthreads();
    }
    trace_mark(g, rg.kind, 1);
}

// ---------------------------------------------------------------------------------------------
// k_cmac, variant "ldg": same decomposition, every thread loads its own vectors straight from global
// memory (read-once, L1 bypass).  Kept as the comparison point for the TMA ring (profiles/).
// ---------------------------------------------------------------------------------------------
// blockIdx.y = hop j of a batch of hops (small engines, hb_conv.cu process_core): the frame of hop j sits j slots below rg.slot and
// its partial segments go set_stride vectors further on -- nb hops of a launch-latency-bound engine in one launch.
template <class T, int XA, int OB>
__global__ void __launch_bounds__(256) k_cmac_ldg(const Geom g, const Range rg_in, const typename VecOf<T>::type *__restrict__ H,
                                                  const typename VecOf<T>::type *__restrict__ X,
                                                  typename VecOf<T>::type *__restrict__ S, const uint64_t set_stride)
{
    typedef typename VecOf<T>::type V;
    Range rg = rg_in;
    if (blockIdx.y)
    {
        const uint32_t j = blockIdx.y;
        rg.slot = rg.slot >= j ? rg.slot - j : rg.slot + g.R - j;
        S += uint64_t(j) * set_stride;
    }
    const uint32_t tid = threadIdx.x;
    const uint32_t tx = tid % g.TX, ty = tid / g.TX;
    trace_mark(g, rg.kind, 0);
    if (ty >= g.TY) return;

    const uint64_t u0 = unit_begin(blockIdx.x, rg.U, rg.G), u1 = unit_begin(blockIdx.x + 1, rg.U, rg.G);
    Cursor cur;
    cur.seek(rg, u0);
    V acc[XA * OB];
#pragma unroll
    for (int r = 0; r < XA * OB; r++) vzero(acc[r]);
    const uint32_t lane_off = ty * g.TBV + tx;

    uint64_t u = u0;
    while (u < u1)
    {
        // a run = consecutive partitions of one (tile, in): the IR pointer just advances by Q per step
        uint32_t run = rg.p0 + rg.pc - cur.p;
        if (uint64_t(run) > u1 - u) run = (uint32_t) (u1 - u);
        const V *hp = H + cur.h_off(g) + lane_off;
        const V *xbase = X + cur.x_off(g, rg) + tx;      // slot (rg.slot + p) of this (group, in, bin-tile)
        uint32_t s = rg.slot + cur.p;
        if (s >= g.R) s -= g.R;
        const V *xp = xbase;
        const V *xwrap = xbase - uint64_t(s) * g.TBV;    // slot 0
#pragma unroll 2
        for (uint32_t r = 0; r < run; r++)
        {
            V xv[XA], hv[XA * OB];
#pragma unroll
            for (int a = 0; a < XA; a++) xv[a] = xp[g.TX * a];
#pragma unroll
            for (int b = 0; b < OB; b++)
#pragma unroll
                for (int a = 0; a < XA; a++) hv[b * XA + a] = ld_stream(hp + (g.TY * b) * g.TBV + g.TX * a);
#pragma unroll
            for (int b = 0; b < OB; b++)
#pragma unroll
                for (int a = 0; a < XA; a++) cmac(acc[b * XA + a], xv[a], hv[b * XA + a]);
            hp += g.Q;
            xp += g.TBV;
            if (++s == g.R) { s = 0; xp = xwrap; }
        }
        u += run;
        const uint32_t tile_done = cur.tile;
        bool last = false;
        cur.p += run - 1;
        last = cur.advance(g, rg) || (u == u1);
        if (last)
        {
            V *seg = S + (uint64_t(blockIdx.x) + tile_done) * g.Q + lane_off;
#pragma unroll
            for (int b = 0; b < OB; b++)
#pragma unroll
                for (int a = 0; a < XA; a++)
                {
                    seg[(g.TY * b) * g.TBV + g.TX * a] = acc[b * XA + a];
                    vzero(acc[b * XA + a]);
                }
        }
    }
    trace_mark(g, rg.kind, 1);
}

// ---------------------------------------------------------------------------------------------
// block reduction of one value (all threads call; result valid in every thread)
// ---------------------------------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T block_sum(T v, T *red /* >= 33 elements of shared memory */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const uint32_t w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        T t = threadIdx.x < nw ? red[threadIdx.x] : T(0);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// complex-element address of bin k inside the tiled layouts (TB = TBV*CPV complex bins per tile)
template <class T> __device__ __forceinline__ uint32_t tile_bins(const Geom &g) { return g.TBV * VecOf<T>::CPV; }

// ---------------------------------------------------------------------------------------------
// k_fwd: real FFT of the newest frame of every input channel into FDL slot g.slot.
// Row `ch` of the previous hop's B samples starts at prev + ch*prev_ld, of the newest B samples at
// newest + ch*new_ld (staging rows or the caller's own buffer).  The reference transforms the rotated
// frame [newest B | previous B] (PartitionedConvolve.cpp:304-305,357), so that the valid half of the
// inverse is the FIRST B samples.  `save` (optional) receives a copy of the newest samples.
// ---------------------------------------------------------------------------------------------
// two consecutive samples at p (vector load when the row is aligned for it)
template <class T> struct Pair { T a, b; };
__device__ __forceinline__ Pair<float> ld_pair(const float *p, bool vec)
{
    Pair<float> r;
    if (vec) { const float2 v = *reinterpret_cast<const float2 *>(p); r.a = v.x; r.b = v.y; }
    else { r.a = p[0]; r.b = p[1]; }
    return r;
}
__device__ __forceinline__ Pair<double> ld_pair(const double *p, bool vec)
{
    Pair<double> r;
    if (vec) { const double2 v = *reinterpret_cast<const double2 *>(p); r.a = v.x; r.b = v.y; }
    else { r.a = p[0]; r.b = p[1]; }
    return r;
}
__device__ __forceinline__ void st_pair(float *p, float a, float b, bool vec)
{
    if (vec) *reinterpret_cast<float2 *>(p) = make_float2(a, b);
    else { p[0] = a; p[1] = b; }
}
__device__ __forceinline__ void st_pair(double *p, double a, double b, bool vec)
{
    if (vec) *reinterpret_cast<double2 *>(p) = make_double2(a, b);
    else { p[0] = a; p[1] = b; }
}
template <class T> __device__ __forceinline__ bool pair_aligned(const T *p, size_t ld)
{
    return ((reinterpret_cast<uintptr_t>(p) | (ld * sizeof(T))) & (2 * sizeof(T) - 1)) == 0;
}

template <class T, int EPT>
__global__ void __launch_bounds__(512) k_fwd(const Geom g, const T *__restrict__ prev, size_t prev_ld,
                                              const T *__restrict__ newest, size_t new_ld, T *__restrict__ save, size_t save_ld,
                                              Cx<T> *__restrict__ X, T *__restrict__ Xnyq,
                                              const Cx<T> *__restrict__ tw, int tw_log2, int stage_tw)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t ch = blockIdx.x;
    const uint32_t B = g.B;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    trace_mark(g, 0, 0);
    // blockIdx.y = hop j of a multi-hop batch (gridDim.y hops, consecutive blocks of the caller's rows): hop j's frame is
    // [block j | block j - 1] and goes to slot g.slot - j; only the last hop leaves its block behind in `save`
    uint32_t slot = g.slot;
    if (blockIdx.y)
    {
        const uint32_t j = blockIdx.y;
        prev = newest + size_t(j - 1) * B; prev_ld = new_ld;
        newest += size_t(j) * B;
        slot = slot >= j ? slot - j : slot + g.R - j;
    }
    if (blockIdx.y + 1 != gridDim.y) save = nullptr;
    // twiddles of this transform size staged in shared memory behind the data (one global round trip
    // instead of one per pass); the loads are issued together with the input loads below
    Cx<T> *stw = s + padded_elems<HB_PADSH>(B);
    Cx<T> twr[EPT];
    if (stage_tw) twiddle_stage_load<T, EPT>(twr, tw, tw_log2, (int) g.log2n);
    // rotated frame [newest B | previous B], de-interleaved on the way in: z[k] = frame[2k] + i frame[2k+1].
    // Every thread issues all of its loads before the first store, so their latencies overlap.
    const T *pn = newest + size_t(ch) * new_ld, *pp = prev + size_t(ch) * prev_ld;
    T *ps = save ? save + size_t(ch) * save_ld : nullptr;
    const bool vn = pair_aligned(newest, new_ld), vp = pair_aligned(prev, prev_ld), vs = save && pair_aligned(save, save_ld);
    Pair<T> v[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr, j = 2 * k;
        if (k < B) v[e] = j < B ? ld_pair(pn + j, vn) : ld_pair(pp + (j - B), vp);
    }
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr, j = 2 * k;
        if (k < B)
        {
            s[sidx<HB_PADSH>(k)] = cx<T>(v[e].a, v[e].b);
            if (ps && j < B) st_pair(ps + j, v[e].a, v[e].b, vs);      // becomes the previous hop of the next call
        }
    }
    if (stage_tw)
    {
        twiddle_stage_store<T, EPT>(stw, twr, (int) g.log2n);
        tw = stw;
        tw_log2 = (int) g.log2n;
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, false, tw, tw_log2);
    __syncthreads();
    const uint32_t TB = tile_bins<T>(g);
    Cx<T> *xrow = X + size_t(ch) * g.n_bt * g.R * TB;
#pragma unroll
    for (int e = 0; e < EPT; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B)
        {
            Cx<T> z = s[sidx<HB_PADSH>(k)];
            if (k == 0)
            {
                Xnyq[size_t(ch) * g.R + slot] = z.y;
                z.y = T(0);
            }
            const uint32_t bt = k / TB, j = k - bt * TB;
            xrow[(size_t(bt) * g.R + slot) * TB + j] = z;
        }
    }
    trace_mark(g, 0, 1);
}

// ---------------------------------------------------------------------------------------------
// k_inv: one CTA per output channel.  Sums the stream-K partial segments of every bin tile in CTA
// order (deterministic), adds the Nyquist dot product, inverse real FFT, scale 1/(4N)
// (PartitionedConvolve.cpp:232-241,359-360) and store of the first B samples at yout row ch, offset off
// (added to what is there when add_result).  carry_dst (optional): first copy (or add) the B samples at
// carry_src row ch to carry_dst row ch -- the output ring read of PartitionedConvolve.cpp:307.
// ---------------------------------------------------------------------------------------------
// Multi-GPU exchange fused into the inverse-FFT epilogue (input channels sharded over `world` ranks, one
// process per GPU).  Every rank holds an "inbox" [2 parities][world sources][outs/world][slot] plus arrival
// counters [2][world]; the CTA of output o stores its partial block straight into the inbox of the rank that
// owns o -- a plain st.global on a peer-mapped pointer, NVLink carries it -- and then bumps that rank's
// counter for (parity, source).  k_gather on the owner waits for the counters and sums the sources in rank
// order (deterministic).  world == 0: not sharded.
constexpr int HB_MAX_WORLD = 16;
// The inbox keeps HB_INBOX_DEPTH hops apart ("parity" = hop sequence number mod depth).  A rank may be ahead of an owner by
// the hops of its own launch plus those of the owner's launch in flight; with at most 8 hops per multi-hop launch, 16 slots
// are never overwritten before the owner's sum has read them (the stream order of gather -> inverse on every rank does the
// flow control: DESIGN.md 6).
constexpr uint32_t HB_INBOX_DEPTH = 16;
struct PeerOut
{
    void *data[HB_MAX_WORLD];          // inbox data of every rank (own entry = local memory)
    uint32_t *count[HB_MAX_WORLD];     // arrival counters of every rank
    uint32_t world, rank, outs_local, parity;      // parity: inbox slot of hop 0 of the launch (hop j of a batch: parity + j mod depth)
    uint64_t slot;                     // elements per inbox block (hop capacity)
};

// multi-hop batch description for k_inv (all zero / last_j = -1 for a single hop)
struct InvBatch
{
    uint64_t set_stride;   // vectors between the partial-segment sets of consecutive hops
    void *last_yout;       // where the hop flagged last_j leaves its block (staging row), leading dimension last_ld
    uint64_t last_ld;
    int32_t last_j;
    int32_t pad;
};

// partial segments k_inv sums: one set per multiply-accumulate launch that contributed to this hop
struct SegSet
{
    const void *S;         // [cta + tile][row][TBV] of that launch
    uint64_t U;            // its work-item count, items per (virtual) tile and grid (Range)
    uint32_t upt, G;
    uint32_t split;        // 0 / 1: one segment holds all OT rows of a tile; 2: a launch that worked on half units (hb_conv_mh.cuh):
    uint32_t pad;          //        virtual tile = tile * 2 + (row >= OT / 2), segments of Q / 2 vectors
};
struct SegSets
{
    SegSet s[2];
    int n;
};

template <class T, int EPT>
__global__ void __launch_bounds__(512) k_inv(const Geom g, const SegSets sets,
                                              const T *__restrict__ Xnyq, const T *__restrict__ Hnyq,
                                              T *__restrict__ yout, size_t ld, size_t off, int add_result,
                                              const T *__restrict__ carry_src, size_t carry_src_ld,
                                              T *__restrict__ carry_dst, size_t carry_dst_ld, int add_carry,
                                              const Cx<T> *__restrict__ tw, int tw_log2, int stage_tw, const PeerOut peer, const InvBatch ib)
{
    typedef typename VecOf<T>::type V;
    constexpr int CPV = VecOf<T>::CPV;
    constexpr int VPT = EPT / CPV;                  // 16-byte vectors of the spectrum per thread
    // blockIdx.y = hop j of a multi-hop batch: its partial segments are set_stride vectors further on, its Nyquist products
    // run from slot g.slot - j, its block goes B samples further in the caller's row -- except the hop flagged as the last of
    // the call, which stays behind in the staging row (ib.last_*); only hop 0 hands over the previous block
    uint32_t nyq_slot = g.slot;
    uint64_t seg_shift = 0;
    if (blockIdx.y)
    {
        const uint32_t j = blockIdx.y;
        nyq_slot = nyq_slot >= j ? nyq_slot - j : nyq_slot + g.R - j;
        seg_shift = uint64_t(j) * ib.set_stride;
        off += size_t(j) * g.B;
        carry_dst = nullptr;
    }
    if (ib.last_j >= 0 && blockIdx.y == (uint32_t) ib.last_j)
    {
        yout = reinterpret_cast<T *>(ib.last_yout); ld = ib.last_ld; off = 0; add_result = 0;
    }
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    __shared__ T red[40];
    const uint32_t ch = blockIdx.x;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t grp = ch / g.outs, o = ch - grp * g.outs;
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    const uint32_t B = g.B;
    trace_mark(g, 3, 0);
    if (stage_tw)
    {
        // twiddles of this transform size into shared memory (published by the barriers of block_sum below)
        Cx<T> *stw = s + padded_elems<HB_PADSH>(B);
        Cx<T> twr[EPT];
        twiddle_stage_load<T, EPT>(twr, tw, tw_log2, (int) g.log2n);
        twiddle_stage_store<T, EPT>(stw, twr, (int) g.log2n);
        tw = stw;
        tw_log2 = (int) g.log2n;
    }

    if (carry_dst)
    {
        // hand the block computed by the previous hop to the caller before this hop's result replaces it
        const T *cs = carry_src + size_t(ch) * carry_src_ld;
        T *cd = carry_dst + size_t(ch) * carry_dst_ld;
        const bool vs = pair_aligned(carry_src, carry_src_ld), vd = pair_aligned(carry_dst, carry_dst_ld);
        Pair<T> cv[EPT / 2], dv[EPT / 2];
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (2 * k < B)
            {
                cv[e] = ld_pair(cs + 2 * k, vs);
                if (add_carry) dv[e] = ld_pair(cd + 2 * k, vd);
            }
        }
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (2 * k < B)
            {
                if (add_carry) st_pair(cd + 2 * k, dv[e].a + cv[e].a, dv[e].b + cv[e].b, vd);
                else st_pair(cd + 2 * k, cv[e].a, cv[e].b, vd);
            }
        }
    }

    // stream-K partials of this output row, four spectrum vectors of the thread at a time: up to four
    // segments per vector are fetched per round so that the loads overlap; the order of the additions is
    // fixed (CTA order), so results are reproducible
    constexpr int GV = VPT < 4 ? VPT : 4;
#pragma unroll 1
    for (int e0 = 0; e0 < VPT; e0 += GV)
    {
        V sum[GV];
#pragma unroll
        for (int e = 0; e < GV; e++) vzero(sum[e]);
#pragma unroll 1
        for (int q = 0; q < sets.n; q++)
        {
            const V *__restrict__ S = reinterpret_cast<const V *>(sets.s[q].S) + seg_shift;
            const uint64_t sU = sets.s[q].U;
            const uint32_t sG = sets.s[q].G, supt = sets.s[q].upt;
            const uint32_t split = sets.s[q].split > 1 ? sets.s[q].split : 1u;
            const uint32_t rh = g.OT / split, Qs = g.Q / split;      // rows and vectors per segment
            const uint32_t vhalf = row / rh, vrow = row - vhalf * rh;
            uint32_t seg_lo[GV], seg_hi[GV];
            uint64_t base[GV];
            uint32_t rounds = 0;
#pragma unroll
            for (int e = 0; e < GV; e++)
            {
                const uint32_t v = tid + (e0 + e) * nthr;
                seg_lo[e] = 1; seg_hi[e] = 0; base[e] = 0;
                if (v < B / CPV)
                {
                    const uint32_t bt = v / g.TBV, xa = v - bt * g.TBV;
                    const uint32_t tile = ((grp * g.n_ot + ot) * g.n_bt + bt) * split + vhalf;
                    const uint64_t ulo = uint64_t(tile) * supt, uhi = ulo + supt - 1;
                    seg_lo[e] = (uint32_t) unit_owner(ulo, sU, sG);
                    seg_hi[e] = (uint32_t) unit_owner(uhi, sU, sG);
                    base[e] = uint64_t(tile) * Qs + vrow * g.TBV + xa;
                    const uint32_t need = seg_hi[e] - seg_lo[e] + 1;
                    rounds = need > rounds ? need : rounds;
                }
            }
            for (uint32_t r = 0; r < rounds; r += 4)
            {
                V part[GV][4];
#pragma unroll
                for (int e = 0; e < GV; e++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const uint32_t c = seg_lo[e] + r + k;
                        if (c <= seg_hi[e]) part[e][k] = S[uint64_t(c) * Qs + base[e]];
                        else vzero(part[e][k]);
                    }
#pragma unroll
                for (int e = 0; e < GV; e++)
#pragma unroll
                    for (int k = 0; k < 4; k++) vadd(sum[e], part[e][k]);
            }
        }
#pragma unroll
        for (int e = 0; e < GV; e++)
        {
            const uint32_t v = tid + (e0 + e) * nthr;
            if (v < B / CPV)
            {
                if constexpr (CPV == 2)
                {
                    s[sidx<HB_PADSH>(2 * v)] = cx<T>(sum[e].x, sum[e].y);
                    s[sidx<HB_PADSH>(2 * v + 1)] = cx<T>(sum[e].z, sum[e].w);
                }
                else
                    s[sidx<HB_PADSH>(v)] = cx<T>(sum[e].x, sum[e].y);
            }
        }
    }
    // Nyquist bin: a real dot product over (in, partition)
    T part = T(0);
    const T *hn = Hnyq + (size_t(grp) * g.outs + o) * g.ins * g.Pcap;
    const T *xn = Xnyq + size_t(grp) * g.ins * g.R;
    for (uint32_t idx0 = tid; idx0 < g.upt; idx0 += 4 * nthr)
    {
        T xv[4], hv[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const uint32_t idx = idx0 + q * nthr;
            xv[q] = hv[q] = T(0);
            if (idx < g.upt)
            {
                const uint32_t in = idx / g.P, p = idx - in * g.P;
                uint32_t sl = nyq_slot + p;
                if (sl >= g.R) sl -= g.R;
                xv[q] = xn[size_t(in) * g.R + sl];
                hv[q] = hn[size_t(in) * g.Pcap + p];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) part += xv[q] * hv[q];
    }
    const T nyq = block_sum<T>(part, red);          // contains the barriers that publish s[]
    if (tid == 0) s[sidx<HB_PADSH>(0)].y = nyq;
    __syncthreads();
    block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, true, tw, tw_log2);
    __syncthreads();
    // hisstools_ifft = forward transform on exchanged planes (Core:1341-1346); each thread swaps its own elements
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
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    const T scale = T(1) / T(size_t(4) << g.log2n);
    if (peer.world)
    {
        // partial block of output `ch` -> inbox of its owner, slot (parity, this rank, local output index)
        const uint32_t owner = ch / peer.outs_local, o_loc = ch - owner * peer.outs_local;
        const uint32_t par = (peer.parity + blockIdx.y) % HB_INBOX_DEPTH;
        T *pd = reinterpret_cast<T *>(peer.data[owner]) + ((size_t(par) * peer.world + peer.rank) * peer.outs_local + o_loc) * peer.slot;
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B / 2)
            {
                const Cx<T> z = s[sidx<HB_PADSH>(k)];
                st_pair(pd + 2 * k, z.y * scale, z.x * scale, true);
            }
        }
        // publish: every thread's stores are ordered before the arrival count at system scope
        __threadfence_system();
        __syncthreads();
        if (tid == 0)
        {
            __threadfence_system();
            atomicAdd_system(peer.count[owner] + par * peer.world + peer.rank, 1u);
        }
        trace_mark(g, 3, 1);
        return;
    }
    T *dst = yout + size_t(ch) * ld + off;
    const bool vd = pair_aligned(yout + off, ld);
    Pair<T> old[EPT / 2];
    if (add_result)
    {
#pragma unroll
        for (int e = 0; e < EPT / 2; e++)
        {
            const uint32_t k = tid + e * nthr;
            if (k < B / 2) old[e] = ld_pair(dst + 2 * k, vd);
        }
    }
#pragma unroll
    for (int e = 0; e < EPT / 2; e++)
    {
        const uint32_t k = tid + e * nthr;
        if (k < B / 2)
        {
            const Cx<T> z = s[sidx<HB_PADSH>(k)];
            if (add_result) st_pair(dst + 2 * k, old[e].a + z.y * scale, old[e].b + z.x * scale, vd);
            else st_pair(dst + 2 * k, z.y * scale, z.x * scale, vd);
        }
    }
    trace_mark(g, 3, 1);
}

// ---------------------------------------------------------------------------------------------
// k_gather: the owner's half of the fused exchange.  One CTA per owned output: wait until every source
// rank has delivered this hop's blocks (arrival counter >= expected; bounded wait, a lost peer raises a flag
// instead of hanging or trapping the device), hand the previously finished block to the caller (carry), then sum the
// sources in rank order into `yout` (the caller's rows or the staging rows kept for the next call).
// ---------------------------------------------------------------------------------------------
// A rank with no impulse response loaded still takes part in every hop of the fused exchange: it delivers silence to every
// owner and bumps the arrival counters, so that the peers' k_gather never waits for it and the hop sequence stays in step
// (the NCCL exchange contributes silence in the same state).  One CTA per output channel.
template <class T>
__global__ void __launch_bounds__(256) k_shard_silence(const PeerOut peer, uint32_t B)
{
    const uint32_t ch = blockIdx.x;
    const uint32_t owner = ch / peer.outs_local, o_loc = ch - owner * peer.outs_local;
    const uint32_t par = peer.parity % HB_INBOX_DEPTH;
    T *pd = reinterpret_cast<T *>(peer.data[owner]) + ((size_t(par) * peer.world + peer.rank) * peer.outs_local + o_loc) * peer.slot;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) pd[k] = T(0);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        atomicAdd_system(peer.count[owner] + par * peer.world + peer.rank, 1u);
    }
}

// The fused exchange at FFT sizes above the single-CTA limit: the inverse transform is a chain of launches (hb_conv_big.cuh) whose
// last kernel has many CTAs per row, so it leaves this rank's partial blocks in a local buffer and this kernel -- one CTA per output
// channel, as the epilogue of k_inv -- delivers each block to its owner's inbox and counts the arrival.
template <class T>
__global__ void __launch_bounds__(256) k_shard_deliver(const T *__restrict__ part, size_t ld, const PeerOut peer, uint32_t B)
{
    const uint32_t ch = blockIdx.x;
    const uint32_t owner = ch / peer.outs_local, o_loc = ch - owner * peer.outs_local;
    const uint32_t par = peer.parity % HB_INBOX_DEPTH;
    T *pd = reinterpret_cast<T *>(peer.data[owner]) + ((size_t(par) * peer.world + peer.rank) * peer.outs_local + o_loc) * peer.slot;
    const T *src = part + size_t(ch) * ld;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) pd[k] = src[k];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        atomicAdd_system(peer.count[owner] + par * peer.world + peer.rank, 1u);
    }
}

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// `timeout_ns`: how long the owner waits for a peer's blocks; a peer that is later than that (stopped, crashed) raises
// *late (read by hb_conv_shard_status) and the hop is summed from what has arrived -- the device is never trapped.
// blockIdx.y = hop j of a multi-hop batch: inbox slot parity + j, arrival count expected.e[j], block j of the caller's rows --
// except the hop flagged as the last of the call, which stays behind in the staging row (as InvBatch for k_inv)
struct GatherBatch
{
    uint32_t e[8];         // expected arrival count per hop
    void *last_yout;
    uint64_t last_ld;
    int32_t last_j;
    int32_t pad;
};

template <class T>
__global__ void __launch_bounds__(256) k_gather(const T *__restrict__ inbox, const uint32_t *count, uint32_t world, uint32_t outs_local,
                                                uint32_t parity, uint64_t slot, const GatherBatch gb, uint32_t B,
                                                T *__restrict__ yout, size_t ld, size_t off, int add_result,
                                                const T *__restrict__ carry_src, size_t carry_src_ld,
                                                T *__restrict__ carry_dst, size_t carry_dst_ld, int add_carry,
                                                unsigned long long *trace, uint32_t hop, unsigned long long timeout_ns, uint32_t *late)
{
    const uint32_t o = blockIdx.x;
    const uint32_t expected = gb.e[blockIdx.y];
    parity = (parity + blockIdx.y) % HB_INBOX_DEPTH;
    if (blockIdx.y)
    {
        off += size_t(blockIdx.y) * B;
        carry_dst = nullptr;
    }
    if (gb.last_j >= 0 && blockIdx.y == (uint32_t) gb.last_j)
    {
        yout = reinterpret_cast<T *>(gb.last_yout); ld = gb.last_ld; off = 0; add_result = 0;
    }
    trace_mark(trace, hop, 4, 0);
    if (carry_dst)
    {
        const T *cs = carry_src + size_t(o) * carry_src_ld;
        T *cd = carry_dst + size_t(o) * carry_dst_ld;
        for (uint32_t k = threadIdx.x; k < B; k += blockDim.x) cd[k] = add_carry ? cd[k] + cs[k] : cs[k];
    }
    if (threadIdx.x < world)
    {
        const volatile uint32_t *cnt = count + parity * world + threadIdx.x;
        const unsigned long long t0 = global_ns();
        for (uint32_t spins = 0; int32_t(*cnt - expected) < 0; spins++)
            if ((spins & 1023u) == 1023u && global_ns() - t0 > timeout_ns) { atomicOr(late, 1u << threadIdx.x); break; }
        __threadfence_system();
    }
    __syncthreads();
    T *dst = yout + size_t(o) * ld + off;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        T sum = T(0);
        for (uint32_t r = 0; r < world; r++) sum += __ldcg(inbox + ((size_t(parity) * world + r) * outs_local + o) * slot + k);
        dst[k] = add_result ? dst[k] + sum : sum;
    }
    trace_mark(trace, hop, 4, 1);
}

// ---------------------------------------------------------------------------------------------
// k_ir: impulse response -> partition spectra of one (group, in, out) pair
// (PartitionedConvolve.cpp:203-219).  blockIdx.x = partition; `taps` effective taps at `ir`
// (offset / length clipping already applied by the host).  Partitions past the end are written as zeros.
// ---------------------------------------------------------------------------------------------
// side != nullptr: the spectra go to a private buffer instead ([partition][B] bins in natural order, then one Nyquist value per
// partition) -- a pair that is replaced while the stream runs is revealed partition by partition (hb_conv.cu, k_pair_copy).
template <class T, int EPT>
__global__ void __launch_bounds__(512) k_ir(const Geom g, const T *__restrict__ ir, size_t taps,
                                             uint32_t grp, uint32_t in, uint32_t o,
                                             Cx<T> *__restrict__ H, T *__restrict__ Hnyq,
                                             const Cx<T> *__restrict__ tw, int tw_log2, Cx<T> *__restrict__ side, uint32_t side_parts)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t p = blockIdx.x;
    const uint32_t B = g.B;
    const size_t begin = size_t(p) * B;
    const size_t have = taps > begin ? (taps - begin < B ? taps - begin : B) : 0;   // block-uniform
    if (have)
    {
        const T *src = ir + begin;
        for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
        {
            size_t j = 2 * size_t(k);
            T a = j < have ? src[j] : T(0);
            T b = j + 1 < have ? src[j + 1] : T(0);
            s[sidx<HB_PADSH>(k)] = cx<T>(a, b);
        }
        __syncthreads();
        block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
        block_real_split<T, EPT, HB_PADSH>(s, B, (int) g.log2n, false, tw, tw_log2);
        __syncthreads();
    }
    const uint32_t TB = tile_bins<T>(g);
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        Cx<T> v = have ? s[sidx<HB_PADSH>(k)] : cx<T>(T(0), T(0));
        if (side)
        {
            if (k == 0) { reinterpret_cast<T *>(side + size_t(side_parts) * B)[p] = v.y; v.y = T(0); }
            side[size_t(p) * B + k] = v;
            continue;
        }
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

// Spectra of ONE pair between the engine's unit layout and a private buffer (layout as k_ir's `side`), partitions p0 + blockIdx.x:
// mode 0 reveal (H <- side), 1 gather (side <- H), 2 hide (H <- 0).  A pair whose impulse response is replaced (or that is reset)
// while the stream runs starts again from silence without disturbing the other pairs: its partitions are hidden and come back one
// per hop, partition p at the hop where the frame it meets is the first one recorded after the restart (hb_conv.cu advance_reveals).
template <class T>
__global__ void __launch_bounds__(256) k_pair_copy(const Geom g, Cx<T> *__restrict__ side, uint32_t side_parts, uint32_t grp, uint32_t in, uint32_t o,
                                                   uint32_t p0, Cx<T> *__restrict__ H, T *__restrict__ Hnyq, int mode)
{
    const uint32_t p = p0 + blockIdx.x;
    const uint32_t B = g.B, TB = tile_bins<T>(g);
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    T *side_nyq = reinterpret_cast<T *>(side + size_t(side_parts) * B);
    T *hn = Hnyq + ((size_t(grp) * g.outs + o) * g.ins + in) * g.Pcap + p;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        const uint32_t bt = k / TB, j = k - bt * TB;
        const uint64_t tile = (uint64_t(grp) * g.n_ot + ot) * g.n_bt + bt;
        const uint64_t unit = (tile * g.ins + in) * g.Pcap + p;
        Cx<T> *h = H + unit * (uint64_t(g.Q) * VecOf<T>::CPV) + size_t(row) * TB + j;
        if (mode == 0) *h = side[size_t(p) * B + k];
        else if (mode == 1) side[size_t(p) * B + k] = *h;
        else *h = cx<T>(T(0), T(0));
    }
    if (threadIdx.x == 0)
    {
        if (mode == 0) *hn = side_nyq[p];
        else if (mode == 1) side_nyq[p] = *hn;
        else *hn = T(0);
    }
}

// rows copy / accumulate: dst[r][0..n) (+)= src[r][0..n)
template <class T>
__global__ void k_rows(T *__restrict__ dst, size_t dld, const T *__restrict__ src, size_t sld, size_t n, int add)
{
    const size_t r = blockIdx.y;
    T *d = dst + r * dld;
    const T *s = src + r * sld;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        d[i] = add ? d[i] + s[i] : s[i];
}

} // namespace hb
