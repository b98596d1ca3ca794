// hb_conv_mh.cu -- multi-hop reuse for float engines on the packed FP32 pipe (FFMA2): kernel k_cmac_mh2 and its launcher.
//
// A call that brings NH hops at once streams every impulse-response spectrum from HBM ONCE and multiplies it into NH
// accumulator sets (hop j of partition p meets FDL slot rg.slot - j + p).  The bytes per call are those of one hop, so the
// multiply-accumulate arithmetic, not HBM, decides how many hops one pass can carry: one complex MAC is 4 FMAs, and a
// three-register FFMA issues at half rate on this part (one warp instruction per 2 cycles and SM sub-partition), so the
// scalar kernel (k_cmac_tma_mh, kept for double) is pipe-bound at NH = 4: ncu showed 44 % issue slots = 88 % of the FMA pipe
// (profiles/r1_cmac_mh4_c4_ncu_details.txt).  Here the complex MAC is two packed fma.rn.f32x2 (SASS FFMA2) on (re, im) pairs:
//
//     acc(re, im) += (xr, xr) * (hr, hi)          acc(re, im) += (xi, xi) * (-hi, hr)
//
// -- half the pipe time per MAC.  Both operand shapes are free in SASS: FFMA2 takes a scalar register broadcast to both halves
// (R.F32) for (xr, xr) / (xi, xi) and selects the halves of a register pair (R.F32x2.LO_HI) for the swap, so the delay-line
// tiles are used exactly as the bulk copies deliver them and the h side costs one negation per bin and thread ((hr, -hi), kept
// for the NH hops).  Per work item a thread issues 128 FFMA2, 12 LDS.128 and ~20 other instructions; every shared-memory address
// is a compile-time offset from one base register (the tile width TBV is a template parameter).
// (A first version expanded the tiles into (xr, xr, -xi, xi) in shared memory: twice the LDS traffic for x -- at NH = 8 the
// 128 B/clk of shared-memory bandwidth, not the FMA pipe, then set the pace: profiles/r2_mh_kernel.txt.)
//
// NH = 8 needs NH x 32 KiB = 256 KiB of accumulators per IR unit -- the whole register file -- so a CTA then works on HALF
// units (the first or the last OT / 2 rows of a unit: 16 KiB, contiguous in the unit layout) with 4 rows per thread instead
// of 8; the work items of the stream-K decomposition become (tile, half) pairs ("virtual tiles", SPLIT = 2) and k_inv finds
// the partial segments of a row through SegSet::split.
//
// Thread grid as the single-hop kernel for XA = 1: tx = tid % TBV over the 16-byte vectors (2 bins) of a tile row,
// ty = tid / TBV, RB rows per thread (rows ty + TY * b, TY * TBV = 256): 8 consumer warps + 1 producer warp that does nothing
// but issue the bulk copies (address arithmetic on warp-uniform values, one elected lane per copy).  With the producer's
// ~90 instructions per item on a consumer warp, that warp spent half its time issuing and the other seven waited for it at
// the barrier of every step (profiles/r2_mh_kernel.txt).  Nine warps cap a thread at 168 registers: 128 accumulators, the
// resident operand side (16) and the streamed one (4) fit.
// shared memory per item: [H: 256 * RB vectors][FDL tiles: NH * TBV]; a stage holds IPS items; then nstages mbarriers.
#include "hb_common.cuh"
#include "hb_conv_kernels.cuh"
#include "hb_conv_mh.h"

#include <algorithm>
#include <map>
#include <type_traits>
#include <mutex>

namespace hb
{

typedef unsigned long long pk2;        // two floats in a 64-bit register pair: the operand type of fma.rn.f32x2
// (inline PTX throughout: with float2 / __ffma2_rn the compiler builds every non-adjacent pair with MOVs, up to 125 per item)

__device__ __forceinline__ pk2 pk(float lo, float hi)
{
    pk2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(pk2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// (lo, hi) -> (hi, lo): becomes the operand selector .F32x2.LO_HI of the consuming FFMA2
__device__ __forceinline__ pk2 swp(pk2 v)
{
    pk2 r;
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tmov.b64 %0, {hi, lo};\n\t}" : "=l"(r) : "l"(v));
    return r;
}
// (-v, v)
__device__ __forceinline__ pk2 neg_pos(float v)
{
    pk2 r;
    asm("{\n\t.reg .f32 n;\n\tneg.f32 n, %1;\n\tmov.b64 %0, {n, %1};\n\t}" : "=l"(r) : "f"(v));
    return r;
}
// acc += a * b on both lanes (SASS: FFMA2)
__device__ __forceinline__ void fma2(pk2 &acc, pk2 a, pk2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
// two 64-bit pairs (one 16-byte vector) from shared memory at a 32-bit shared address + compile-time byte offset
template <int OFF> __device__ __forceinline__ void lds_pairs(uint32_t addr, pk2 &p0, pk2 &p1)
{
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2+%3];" : "=l"(p0), "=l"(p1) : "r"(addr), "n"(OFF));
}
// 16-byte shared-memory load at a 32-bit shared address + compile-time byte offset
template <int OFF> __device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
    return v;
}

// producer-side cursor over the work items of a launch: item u = ((vt * ins + in) * pc + p), vt = tile * SPLIT + half.
// The IR offset advances by one unit per item while p runs; it is recomputed only when (in, vt) changes.
template <int SPLIT>
struct MhCursor
{
    uint32_t vt, in, p;
    uint64_t hoff, xrow;           // vector offsets: this item's IR (half) unit; slot 0 of the FDL ring of (group, in, bin tile)
    __device__ __forceinline__ void locate(const Geom &g)
    {
        const uint32_t tile = vt / SPLIT, half = vt - tile * SPLIT;
        hoff = ((uint64_t(tile) * g.ins + in) * g.Pcap + p) * g.Q + uint64_t(half) * (g.Q / SPLIT);
        const uint32_t bt = tile % g.n_bt, grp = tile / (g.n_bt * g.n_ot);
        xrow = ((uint64_t(grp) * g.ins + in) * g.n_bt + bt) * g.R * g.TBV;
    }
    __device__ __forceinline__ void seek(const Geom &g, const Range &r, uint64_t u)
    {
        vt = (uint32_t) (u / r.upt);
        uint32_t rem = (uint32_t) (u - uint64_t(vt) * r.upt);
        in = rem / r.pc;
        p = r.p0 + (rem - in * r.pc);
        locate(g);
    }
    __device__ __forceinline__ void advance(const Geom &g, const Range &r)
    {
        hoff += g.Q;
        if (++p < r.p0 + r.pc) return;
        p = r.p0;
        if (++in == g.ins) { in = 0; vt++; }
        locate(g);
    }
};

// IPS = work items per step: one CTA-wide barrier per step.  A step has fixed costs (the barrier and its skew, the exposed
// shared-memory latency behind it, the preparation of the next tiles: ~600 cycles measured) against 128 FFMA2 per warp and
// item (>= 512 cycles of the FMA pipe), so the 16 KiB half-unit items of NH = 8 are taken two per step.
// The step barrier of the warp-specialised part: producer, idle and consumer warps arrive from different places in the code, which
// __syncthreads() does not allow formally (compute-sanitizer synccheck: "divergent threads in block") although it is the same
// hardware barrier; a named barrier with an explicit thread count is the form made for this.
__device__ __forceinline__ void step_barrier()
{
    // (barrier.sync without .aligned: the lanes of the producer warp may not have reconverged behind the elected lane's copies)
    __syncwarp();
    asm volatile("barrier.sync 1, %0;" ::"r"(blockDim.x) : "memory");
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// NH = 8: three warpgroups are launched (168 registers each) and the registers are re-dealt with setmaxnreg -- the producer's
// warpgroup keeps 40, the two consumer warpgroups take 232, which lets both operand sides of an item stay in registers
template <int NH, int RB, int TBV, int IPS>
__global__ void __launch_bounds__(NH == 8 ? 384 : 288, 1) k_cmac_mh2(const Geom g, const Range rg, const float4 *__restrict__ H, const float4 *__restrict__ X,
                                                      float4 *__restrict__ S, const int nstages, const uint64_t set_stride)
{
    constexpr int SPLIT = 8 / RB;
    constexpr uint32_t HQ = 256u * RB;                                // vectors of the IR (half) unit of one item = Q / SPLIT
    constexpr uint32_t ITEM_VECS = HQ + (uint32_t) NH * TBV;          // [H][FDL tiles of the NH hops]
    constexpr uint32_t STAGE_VECS = IPS * ITEM_VECS;
    constexpr uint32_t H_BYTES = HQ * 16u, X_BYTES = TBV * 16u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *ring = reinterpret_cast<float4 *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + size_t(nstages) * STAGE_VECS * sizeof(float4));

    const uint32_t tid = threadIdx.x;
    const bool producer = tid >= 256;                                  // warp 8
    const uint32_t tx = tid % TBV;
    trace_mark(g, rg.kind, 0);

    // work items of this launch: rg.U units x SPLIT halves, dealt as contiguous ranges
    const uint64_t items = rg.U * SPLIT;
    const uint64_t u0 = unit_begin(blockIdx.x, items, rg.G), u1 = unit_begin(blockIdx.x + 1, items, rg.G);
    const uint32_t n = (uint32_t) (u1 - u0);
    const uint32_t nsteps = (n + IPS - 1) / IPS;

    if (tid == 0)
    {
        for (int s = 0; s < nstages; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (producer)
    {
        if (NH == 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid >= 288)
        {
            // warps 9-11 only exist to make the producer's warpgroup whole: they keep the step barriers
            for (uint32_t t = 0; t < nsteps; t++) step_barrier();
            return;
        }
        // ---- producer warp: all 32 lanes walk the cursor (warp-uniform values), one elected lane issues ----
        MhCursor<SPLIT> prod;
        prod.seek(g, rg, u0);
        const uint64_t pol_h = l2_policy_evict_first();
        const uint64_t pol_x = rg.pin_s ? l2_policy_evict_last() : l2_policy_evict_first();     // delay line kept in L2 only while it fits (plan_geometry)
        uint32_t issued = 0;                                          // steps whose copies have been issued
        // one item = the IR (half) unit and the FDL tiles of the NH hops that meet it.  Hop j reads slot s - j (mod R): in
        // memory the tiles are one ascending run of slots [s - NH + 1, s] (hop NH - 1 first, which is also the shared-memory
        // order), or two runs when the ring wraps inside them: [R - (NH - 1 - s), R - 1] then [0, s]
        auto issue = [&](uint32_t st)
        {
            const uint32_t cnt = min((uint32_t) IPS, n - issued * IPS);
            if (elect_one()) mbar_expect_tx(&full[st], cnt * (H_BYTES + NH * X_BYTES));
            __syncwarp();
            for (uint32_t q = 0; q < cnt; q++)
            {
                float4 *dst = ring + size_t(st) * STAGE_VECS + q * ITEM_VECS;
                uint32_t s = rg.slot + prod.p;
                if (s >= g.R) s -= g.R;
                const uint32_t hi_tiles = s >= uint32_t(NH - 1) ? 0u : uint32_t(NH - 1) - s;      // tiles of the wrapped run
                if (elect_one())
                {
                    bulk_g2s(dst, H + prod.hoff, H_BYTES, &full[st], pol_h);
                    if (hi_tiles) bulk_g2s(dst + HQ, X + prod.xrow + uint64_t(g.R - hi_tiles) * TBV, hi_tiles * X_BYTES, &full[st], pol_x);
                    bulk_g2s(dst + HQ + hi_tiles * TBV, X + prod.xrow + uint64_t(s + hi_tiles - (NH - 1)) * TBV, (NH - hi_tiles) * X_BYTES, &full[st], pol_x);
                }
                prod.advance(g, rg);
            }
            issued++;
        };
        while (issued < nsteps && issued + 1 < (uint32_t) nstages) issue(issued);
        uint32_t stage = 0;
        for (uint32_t t = 0; t < nsteps; t++)
        {
            // stage (t + nstages - 1) % nstages was drained in step t-1 (barrier at the end of that step)
            if (issued < nsteps) issue(stage ? stage - 1 : nstages - 1);
            if (++stage == (uint32_t) nstages) stage = 0;
            step_barrier();
        }
        return;
    }

    // ---- consumer warps ----
    if (NH == 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // only the virtual tile matters on this side (where the partial segment goes)
    uint32_t c_vt = (uint32_t) (u0 / rg.upt);
    uint32_t c_left = rg.upt - (uint32_t) (u0 - uint64_t(c_vt) * rg.upt);     // items left in this virtual tile
    pk2 acc[NH][RB][2];
#pragma unroll
    for (int j = 0; j < NH; j++)
#pragma unroll
        for (int b = 0; b < RB; b++) acc[j][b][0] = acc[j][b][1] = 0ull;

    const uint32_t ring_addr = smem_u32(ring);
    // RB rows of one item against the delay-line tiles of its NH hops; hs / xs: this thread's column in the item.
    // Per bin and hop two FFMA2 and nothing else: xr (broadcast) x (hr, hi), then (-hi, hr) x xi (broadcast) -- the swap
    // and the one negated half are operand modifiers of the first source (SASS: -R.F32x2.LO_HI.NP), the broadcasts of the
    // other (R.F32); which source carries which matters to ptxas, hence the operand order below.  The side with fewer
    // vectors (RB rows of h or NH tiles of x) stays in registers, the other is streamed through 4 registers.
    auto load_x = [&](uint32_t xs, int j) -> float4
    {
        switch (NH - 1 - j)                                           // hop j sits at shared-memory position NH - 1 - j
        {
            case 0: return lds128<0 * TBV * 16>(xs);
            case 1: return lds128<1 * TBV * 16>(xs);
            case 2: return lds128<2 * TBV * 16>(xs);
            case 3: return lds128<3 * TBV * 16>(xs);
            case 4: return lds128<4 * TBV * 16>(xs);
            case 5: return lds128<5 * TBV * 16>(xs);
            case 6: return lds128<6 * TBV * 16>(xs);
            default: return lds128<7 * TBV * 16>(xs);
        }
    };
    auto load_h = [&](uint32_t hs, int b) -> float4
    {
        switch (b)
        {
            case 0: return lds128<0 * 4096>(hs);
            case 1: return lds128<1 * 4096>(hs);
            case 2: return lds128<2 * 4096>(hs);
            case 3: return lds128<3 * 4096>(hs);
            case 4: return lds128<4 * 4096>(hs);
            case 5: return lds128<5 * 4096>(hs);
            case 6: return lds128<6 * 4096>(hs);
            default: return lds128<7 * 4096>(hs);
        }
    };
    auto cmac2 = [&](pk2 &a0, pk2 &a1, const float4 &x, const float4 &v)
    {
        fma2(a0, pk(x.x, x.x), pk(v.x, v.y));
        fma2(a1, pk(x.z, x.z), pk(v.z, v.w));
        fma2(a0, pk(-v.y, v.x), pk(x.y, x.y));
        fma2(a1, pk(-v.w, v.z), pk(x.w, x.w));
    };
    auto mac_item = [&](uint32_t hs, uint32_t xs)
    {
        if (NH == 8)
        {
            // 232 registers: all operands of the item are loaded up front, the FFMA2 stream never waits for shared memory
            float4 hv[RB], xv[NH];
#pragma unroll
            for (int b = 0; b < RB; b++) hv[b] = load_h(hs, b);
#pragma unroll
            for (int j = 0; j < NH; j++) xv[j] = load_x(xs, j);
#pragma unroll
            for (int j = 0; j < NH; j++)
#pragma unroll
                for (int b = 0; b < RB; b++) cmac2(acc[j][b][0], acc[j][b][1], xv[j], hv[b]);
        }
        else if (RB <= NH)
        {
            float4 hv[RB];
#pragma unroll
            for (int b = 0; b < RB; b++) hv[b] = load_h(hs, b);
#pragma unroll
            for (int j = 0; j < NH; j++)
            {
                const float4 x = load_x(xs, j);
#pragma unroll
                for (int b = 0; b < RB; b++) cmac2(acc[j][b][0], acc[j][b][1], x, hv[b]);
            }
        }
        else
        {
            float4 xv[NH];
#pragma unroll
            for (int j = 0; j < NH; j++) xv[j] = load_x(xs, j);
#pragma unroll
            for (int b = 0; b < RB; b++)
            {
                const float4 v = load_h(hs, b);
#pragma unroll
                for (int j = 0; j < NH; j++) cmac2(acc[j][b][0], acc[j][b][1], xv[j], v);
            }
        }
    };
    // partial segments of the finished virtual tile to memory, accumulators cleared
    auto flush = [&]()
    {
#pragma unroll
        for (int j = 0; j < NH; j++)
        {
            float4 *seg = S + uint64_t(j) * set_stride + (uint64_t(blockIdx.x) + c_vt) * HQ + tid;
#pragma unroll
            for (int b = 0; b < RB; b++)
            {
                float4 o;
                unpk(acc[j][b][0], o.x, o.y);
                unpk(acc[j][b][1], o.z, o.w);
                seg[256 * b] = o;
                acc[j][b][0] = acc[j][b][1] = 0ull;
            }
        }
        c_vt++;
        c_left = rg.upt;
    };

    uint32_t stage = 0, parity = 0, done = 0;                         // done: items consumed so far
    for (uint32_t t = 0; t < nsteps; t++)
    {
        const uint32_t cnt = min((uint32_t) IPS, n - done);
        mbar_wait(&full[stage], parity);
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t) IPS; q++)
        {
            if (q < cnt)
            {
                const uint32_t item = ring_addr + (stage * STAGE_VECS + q * ITEM_VECS) * 16u;
                mac_item(item + tid * 16u, item + (HQ + tx) * 16u);
                done++;
                if (--c_left == 0 || done == n) flush();
            }
        }
        if (++stage == (uint32_t) nstages) { stage = 0; parity ^= 1; }
        step_barrier();                                               // stage t is free for the producer
    }
    trace_mark(g, rg.kind, 1);
}

namespace
{
// opt in to > 48 KB of dynamic shared memory once per (kernel, device, size), with the largest carveout (as hb_conv.cu)
template <class K> int allow_smem(K kernel, size_t bytes)
{
    static std::mutex m;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    auto it = done.find(std::make_pair((const void *) kernel, dev));
    if (it == done.end())
    {
        HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int) cudaSharedmemCarveoutMaxShared));
        it = done.insert(std::make_pair(std::make_pair((const void *) kernel, dev), size_t(48 * 1024))).first;
    }
    if (it->second >= bytes) return HB_OK;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    it->second = bytes;
    return HB_OK;
}

template <int NH, int RB, int TBV>
int launch_inst(const Geom &g, const Range &r, const void *H, const void *X, void *S, uint64_t set_stride, cudaStream_t st)
{
    constexpr int IPS = RB == 4 ? 2 : 1;                     // half-unit items go two per step
    const int nst = mh2_stages(TBV, NH, RB);
    if (nst < 3) { set_error("internal: multi-hop stage does not fit shared memory"); return HB_ERR_UNSUPPORTED; }
    const size_t smem = size_t(nst) * mh2_stage_bytes(TBV, NH, RB) + size_t(nst) * 8;
    int rc = allow_smem(k_cmac_mh2<NH, RB, TBV, IPS>, smem);
    if (rc) return rc;
    k_cmac_mh2<NH, RB, TBV, IPS><<<r.G, NH == 8 ? 384 : 288, smem, st>>>(g, r, (const float4 *) H, (const float4 *) X, (float4 *) S, nst, set_stride);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <int TBV>
int launch_tbv(const Geom &g, const Range &r, const void *H, const void *X, void *S, int nh, uint64_t set_stride, cudaStream_t st)
{
    switch (nh)
    {
        case 2: return launch_inst<2, 8, TBV>(g, r, H, X, S, set_stride, st);
        case 4: return launch_inst<4, 8, TBV>(g, r, H, X, S, set_stride, st);
        case 8: if (TBV <= 64) return launch_inst<8, 4, (TBV <= 64 ? TBV : 32)>(g, r, H, X, S, set_stride, st);
    }
    set_error("internal: no packed multi-hop kernel for NH=%d TBV=%d", nh, TBV);
    return HB_ERR_UNSUPPORTED;
}
} // namespace

// one ring stage: the items of one step (two half-unit items when rb == 4)
size_t mh2_stage_bytes(uint32_t tbv, int nh, int rb) { return (rb == 4 ? 2 : 1) * (size_t(256) * rb + size_t(nh) * tbv) * 16; }

// ring depth: enough copies in flight to cover the HBM latency of the chip (64 KiB and more per SM beyond the step being
// consumed: profiles/r1_ring_depth.txt)
int mh2_stages(uint32_t tbv, int nh, int rb)
{
    const size_t stage = mh2_stage_bytes(tbv, nh, rb) + 8, tma = stage - 8;
    static const char *env = getenv("HB_MH_STAGES");                         // experiments only
    int st = env && atoi(env) >= 3 ? atoi(env) : (int) ((80 * 1024 + tma - 1) / tma) + 1;
    while (st > 2 && size_t(st) * stage > 226 * 1024) st--;
    return st;
}

bool mh2_supported(const Geom &g, int nh)
{
    if (g.XA != 1 || g.OB != 8 || g.TX != g.TBV || g.TX * g.TY != 256) return false;
    if (g.TBV != 32 && g.TBV != 64 && g.TBV != 128 && g.TBV != 256) return false;
    if (nh == 8 && g.TBV > 64) return false;                // delay-line tiles as large as the IR half unit: not worth it
    if (nh != 2 && nh != 4 && nh != 8) return false;
    return mh2_stages(g.TBV, nh, nh == 8 ? 4 : 8) >= 3;
}

int launch_cmac_mh2(const Geom &g, const Range &r, const void *H, const void *X, void *S, int nh, uint64_t set_stride, cudaStream_t st)
{
    switch (g.TBV)
    {
        case 32: return launch_tbv<32>(g, r, H, X, S, nh, set_stride, st);
        case 64: return launch_tbv<64>(g, r, H, X, S, nh, set_stride, st);
        case 128: return launch_tbv<128>(g, r, H, X, S, nh, set_stride, st);
        case 256: return launch_tbv<256>(g, r, H, X, S, nh, set_stride, st);
    }
    set_error("internal: no packed multi-hop kernel for TBV=%u", g.TBV);
    return HB_ERR_UNSUPPORTED;
}

} // namespace hb
