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
    uint32_t P;            // partitions in use = FDL ring length
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
    uint64_t U;            // units in total = tiles*upt
};

// ---------------------------------------------------------------------------------------------
// vector complex multiply-accumulate
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cmac(float4 &a, const float4 x, const float4 h)
{
    a.x = fmaf(x.x, h.x, a.x); a.x = fmaf(-x.y, h.y, a.x);
    a.y = fmaf(x.x, h.y, a.y); a.y = fmaf(x.y, h.x, a.y);
    a.z = fmaf(x.z, h.z, a.z); a.z = fmaf(-x.w, h.w, a.z);
    a.w = fmaf(x.z, h.w, a.w); a.w = fmaf(x.w, h.z, a.w);
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
    __device__ __forceinline__ void seek(const Geom &g, uint64_t u)
    {
        tile = (uint32_t) (u / g.upt);
        uint32_t rem = (uint32_t) (u - uint64_t(tile) * g.upt);
        in = rem / g.P;
        p = rem - in * g.P;
    }
    // returns true when the step that was just left was the last unit of its tile
    __device__ __forceinline__ bool advance(const Geom &g)
    {
        if (++p < g.P) return false;
        p = 0;
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
    // vector offset of the matching FDL tile
    __device__ __forceinline__ uint64_t x_off(const Geom &g) const
    {
        uint32_t bt = tile % g.n_bt;
        uint32_t grp = tile / (g.n_bt * g.n_ot);
        uint32_t s = g.slot + p;
        if (s >= g.P) s -= g.P;
        return (((uint64_t(grp) * g.ins + in) * g.n_bt + bt) * g.P + s) * g.TBV;
    }
};

// ---------------------------------------------------------------------------------------------
// k_cmac, variant "tma": IR units and FDL tiles are staged in a shared-memory ring by bulk TMA copies
// issued by one thread; all threads consume from shared memory.  One barrier per step.
// dynamic shared memory: nstages * (Q + TBV) vectors, then nstages mbarriers.
// ---------------------------------------------------------------------------------------------
template <class T, int XA, int OB>
__global__ void __launch_bounds__(256, 1) k_cmac_tma(const Geom g, const typename VecOf<T>::type *__restrict__ H,
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

    const uint64_t u0 = unit_begin(blockIdx.x, g.U, g.G), u1 = unit_begin(blockIdx.x + 1, g.U, g.G);
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
    if (tid == 0)
    {
        pol_h = l2_policy_evict_first();
        pol_x = l2_policy_evict_last();
        prod.seek(g, u0);
        for (; issued < n && issued + 1 < (uint32_t) nstages; issued++)
        {
            V *dst = ring + size_t(issued) * stage_vecs;
            mbar_expect_tx(&full[issued], h_bytes + x_bytes);
            bulk_g2s(dst, H + prod.h_off(g), h_bytes, &full[issued], pol_h);
            bulk_g2s(dst + g.Q, X + prod.x_off(g), x_bytes, &full[issued], pol_x);
            prod.advance(g);
        }
    }

    Cursor cons;
    cons.seek(g, u0);
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
            bulk_g2s(dst, H + prod.h_off(g), h_bytes, &full[ps], pol_h);
            bulk_g2s(dst + g.Q, X + prod.x_off(g), x_bytes, &full[ps], pol_x);
            prod.advance(g);
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
        const bool last = cons.advance(g) || (k + 1 == n);
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
}

// ---------------------------------------------------------------------------------------------
// k_cmac, variant "ldg": same decomposition, every thread loads its own vectors straight from global
// memory (read-once, L1 bypass).  Kept as the comparison point for the TMA ring (profiles/).
// ---------------------------------------------------------------------------------------------
template <class T, int XA, int OB>
__global__ void __launch_bounds__(256) k_cmac_ldg(const Geom g, const typename VecOf<T>::type *__restrict__ H,
                                                  const typename VecOf<T>::type *__restrict__ X,
                                                  typename VecOf<T>::type *__restrict__ S)
{
    typedef typename VecOf<T>::type V;
    const uint32_t tid = threadIdx.x;
    const uint32_t tx = tid % g.TX, ty = tid / g.TX;
    if (ty >= g.TY) return;

    const uint64_t u0 = unit_begin(blockIdx.x, g.U, g.G), u1 = unit_begin(blockIdx.x + 1, g.U, g.G);
    Cursor cur;
    cur.seek(g, u0);
    V acc[XA * OB];
#pragma unroll
    for (int r = 0; r < XA * OB; r++) vzero(acc[r]);
    const uint32_t lane_off = ty * g.TBV + tx;

    uint64_t u = u0;
    while (u < u1)
    {
        // a run = consecutive partitions of one (tile, in): the IR pointer just advances by Q per step
        uint32_t run = g.P - cur.p;
        if (uint64_t(run) > u1 - u) run = (uint32_t) (u1 - u);
        const V *hp = H + cur.h_off(g) + lane_off;
        const V *xbase = X + cur.x_off(g) + tx;          // slot (g.slot + p) of this (group, in, bin-tile)
        uint32_t s = g.slot + cur.p;
        if (s >= g.P) s -= g.P;
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
            if (++s == g.P) { s = 0; xp = xwrap; }
        }
        u += run;
        const uint32_t tile_done = cur.tile;
        bool last = false;
        cur.p += run - 1;
        last = cur.advance(g) || (u == u1);
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
// xin: time-ordered staging rows, row `ch` starts at xin + ch*ld; the 2B samples of this hop start at
// `off`.  The reference transforms the rotated frame [newest B | previous B]
// (PartitionedConvolve.cpp:304-305,357), so that the valid half of the inverse is the FIRST B samples.
// ---------------------------------------------------------------------------------------------
template <class T, int EPT>
__global__ void __launch_bounds__(1024) k_fwd(const Geom g, const T *__restrict__ xin, size_t ld, size_t off,
                                              Cx<T> *__restrict__ X, T *__restrict__ Xnyq,
                                              const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t ch = blockIdx.x;
    const uint32_t B = g.B;
    const T *src = xin + size_t(ch) * ld + off;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        uint32_t j = 2 * k;
        uint32_t q = j < B ? j + B : j - B;
        s[sidx<HB_PADSH>(k)] = cx<T>(src[q], src[q + 1]);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    for (uint32_t k = threadIdx.x; k <= B / 2; k += blockDim.x) real_split_pair<T, HB_PADSH>(s, B, (int) g.log2n, k, false, tw, tw_log2);
    __syncthreads();
    const uint32_t TB = tile_bins<T>(g);
    Cx<T> *xrow = X + size_t(ch) * g.n_bt * g.P * TB;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(k)];
        if (k == 0)
        {
            Xnyq[size_t(ch) * g.P + g.slot] = v.y;
            v.y = T(0);
        }
        uint32_t bt = k / TB, j = k - bt * TB;
        xrow[(size_t(bt) * g.P + g.slot) * TB + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// k_inv: one CTA per output channel.  Sums the stream-K partial segments of every bin tile in CTA
// order (deterministic), adds the Nyquist dot product, inverse real FFT, scale 1/(4N)
// (PartitionedConvolve.cpp:232-241,359-360) and store of the first B samples at yout row ch, offset off.
// ---------------------------------------------------------------------------------------------
template <class T, int EPT>
__global__ void __launch_bounds__(1024) k_inv(const Geom g, const typename VecOf<T>::type *__restrict__ S,
                                              const T *__restrict__ Xnyq, const T *__restrict__ Hnyq,
                                              T *__restrict__ yout, size_t ld, size_t off,
                                              const Cx<T> *__restrict__ tw, int tw_log2)
{
    typedef typename VecOf<T>::type V;
    constexpr int CPV = VecOf<T>::CPV;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    __shared__ T red[40];
    const uint32_t ch = blockIdx.x;
    const uint32_t grp = ch / g.outs, o = ch - grp * g.outs;
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    const uint32_t B = g.B;

    for (uint32_t v = threadIdx.x; v < B / CPV; v += blockDim.x)
    {
        const uint32_t bt = v / g.TBV, xa = v - bt * g.TBV;
        const uint32_t tile = (grp * g.n_ot + ot) * g.n_bt + bt;
        const uint64_t ulo = uint64_t(tile) * g.upt, uhi = ulo + g.upt - 1;
        const uint32_t clo = (uint32_t) unit_owner(ulo, g.U, g.G), chi = (uint32_t) unit_owner(uhi, g.U, g.G);
        V sum;
        vzero(sum);
        for (uint32_t c = clo; c <= chi; c++) vadd(sum, S[(uint64_t(c) + tile) * g.Q + row * g.TBV + xa]);
        if constexpr (CPV == 2)
        {
            s[sidx<HB_PADSH>(2 * v)] = cx<T>(sum.x, sum.y);
            s[sidx<HB_PADSH>(2 * v + 1)] = cx<T>(sum.z, sum.w);
        }
        else
            s[sidx<HB_PADSH>(v)] = cx<T>(sum.x, sum.y);
    }
    // Nyquist bin: a real dot product over (in, partition)
    T part = T(0);
    const T *hn = Hnyq + (size_t(grp) * g.outs + o) * g.ins * g.Pcap;
    const T *xn = Xnyq + size_t(grp) * g.ins * g.P;
    for (uint32_t idx = threadIdx.x; idx < g.upt; idx += blockDim.x)
    {
        uint32_t in = idx / g.P, p = idx - in * g.P;
        uint32_t sl = g.slot + p;
        if (sl >= g.P) sl -= g.P;
        part += xn[size_t(in) * g.P + sl] * hn[size_t(in) * g.Pcap + p];
    }
    const T nyq = block_sum<T>(part, red);          // contains the barriers that publish s[]
    if (threadIdx.x == 0) s[sidx<HB_PADSH>(0)].y = nyq;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= B / 2; k += blockDim.x) real_split_pair<T, HB_PADSH>(s, B, (int) g.log2n, k, true, tw, tw_log2);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(k)];
        s[sidx<HB_PADSH>(k)] = cx<T>(v.y, v.x);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    const T scale = T(1) / T(size_t(4) << g.log2n);
    T *dst = yout + size_t(ch) * ld + off;
    for (uint32_t k = threadIdx.x; k < B / 2; k += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(k)];
        dst[2 * k] = v.y * scale;
        dst[2 * k + 1] = v.x * scale;
    }
}

// ---------------------------------------------------------------------------------------------
// k_ir: impulse response -> partition spectra of one (group, in, out) pair
// (PartitionedConvolve.cpp:203-219).  blockIdx.x = partition; `taps` effective taps at `ir`
// (offset / length clipping already applied by the host).  Partitions past the end are written as zeros.
// ---------------------------------------------------------------------------------------------
template <class T, int EPT>
__global__ void __launch_bounds__(1024) k_ir(const Geom g, const T *__restrict__ ir, size_t taps,
                                             uint32_t grp, uint32_t in, uint32_t o,
                                             Cx<T> *__restrict__ H, T *__restrict__ Hnyq,
                                             const Cx<T> *__restrict__ tw, int tw_log2)
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
        for (uint32_t k = threadIdx.x; k <= B / 2; k += blockDim.x) real_split_pair<T, HB_PADSH>(s, B, (int) g.log2n, k, false, tw, tw_log2);
        __syncthreads();
    }
    const uint32_t TB = tile_bins<T>(g);
    const uint32_t ot = o / g.OT, row = o - ot * g.OT;
    for (uint32_t k = threadIdx.x; k < B; k += blockDim.x)
    {
        Cx<T> v = have ? s[sidx<HB_PADSH>(k)] : cx<T>(T(0), T(0));
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
