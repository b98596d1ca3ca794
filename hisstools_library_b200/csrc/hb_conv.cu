// hb_conv.cu -- host side of the uniform partitioned convolution engine behind the hb_conv_* entry
// points of include/hisstools_b200.h (kernels: hb_conv_kernels.cuh).
//
// Streaming bookkeeping (what PartitionedConvolve.cpp:243-385 does with its four FFT buffers and the
// RW counter) is kept on linear device staging rows instead of rings:
//   xin row  = [previous hop: B samples][pending samples of the current hop: rw][new samples of this call]
//   yout row = [result block of the last completed hop: B][blocks of the hops completed in this call]
// so hop h of a call transforms xin[h*B .. h*B + 2B) and the n output samples of the call are
// yout[rw .. rw + n): the output is the linear convolution delayed by exactly B for any call sizes
// (SURVEY A.2).  After a call the tail of both rows becomes the head of the other (ping-pong) row set.
#include "hb_common.cuh"
#include "hb_conv_kernels.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

using namespace hb;

namespace
{
constexpr uintptr_t MIN_FFT_LOG2 = 5;       // PartitionedConvolve.h:18
constexpr uintptr_t MAX_FFT_LOG2 = 20;      // PartitionedConvolve.h:19

// reference error codes (ConvolveErrors.h:4-19)
enum
{
    ERR_NONE = 0,
    ERR_IN_CHAN = 1,
    ERR_OUT_CHAN = 2,
    ERR_MEM_UNAVAILABLE = 3,
    ERR_MEM_ALLOC_TOO_SMALL = 4,
    ERR_TIME_IMPULSE_TOO_LONG = 5,
    ERR_TIME_LENGTH_OUT_OF_RANGE = 6,
    ERR_PARTITION_LENGTH_TOO_LARGE = 7,
    ERR_FFT_SIZE_MAX_TOO_SMALL = 8,
    ERR_FFT_SIZE_MAX_TOO_LARGE = 9,
    ERR_FFT_SIZE_MAX_NON_POWER_OF_TWO = 10,
    ERR_FFT_SIZE_OUT_OF_RANGE = 11,
    ERR_FFT_SIZE_NON_POWER_OF_TWO = 12
};

// ceil(log2(value)) as PartitionedConvolve::log2 (PartitionedConvolve.cpp:114-129)
uintptr_t ceil_log2(uintptr_t value)
{
    uintptr_t bits = 0;
    for (uintptr_t v = value; v; v >>= 1) bits++;
    if (!bits) return 0;
    return value == (uintptr_t(1) << (bits - 1)) ? bits - 1 : bits;
}
} // namespace

struct hb_conv
{
    int dtype = HB_F32;
    int device = 0;
    uint32_t groups = 1, ins = 1, outs = 1;
    uintptr_t max_fft_log2 = 0, fft_log2 = 0;
    uintptr_t max_length = 0;       // taps per pair, rounded up to a multiple of max_fft/2 (cpp:77-82)
    uintptr_t offset = 0, length = 0;
    intptr_t reset_offset = -1;
    bool need_reset = true;

    // device memory
    void *d_H = nullptr;            // IR spectra
    void *d_X = nullptr;            // FDL
    void *d_Hnyq = nullptr, *d_Xnyq = nullptr;
    DevBuf d_S;                     // stream-K partial segments
    void *d_tw = nullptr;
    int tw_log2 = 1;
    DevBuf d_xin[2], d_yout[2];     // ping-pong staging rows
    size_t xin_ld = 0, yout_ld = 0;
    int cur = 0;                    // which staging set holds the retained state
    size_t x_tail = 0, y_tail = 0;  // where the retained head starts inside set `cur`
    DevBuf d_io_in, d_io_out, d_ir; // host-call staging on the device
    PinnedBuf h_in, h_out, h_ir;
    cudaStream_t stream = nullptr;

    // streaming state
    uintptr_t rw = 0;               // samples already received of the current hop
    std::vector<uint32_t> nparts;   // partitions loaded per pair [group][out][in]
    uint32_t P = 0;                 // max over pairs = ring length
    Geom g{};
    int sm_count = 148;
    int ctas_per_sm = 1;
    int variant = 1;                // 1 = TMA ring, 0 = direct loads
    int nstages = 0;
    size_t cmac_smem = 0;
    std::mutex lock;

    // optional per-kernel timing (hb_conv_set_profiling): 4 events per hop on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> ev;
    size_t ev_used = 0;
    double prof_ms[3] = {0, 0, 0};  // forward FFTs, multiply-accumulate, inverse FFTs
    uint64_t prof_hops = 0;

    size_t esize() const { return dtype_size(dtype); }
    size_t pairs() const { return size_t(groups) * ins * outs; }
};

namespace
{

// ---- geometry -------------------------------------------------------------------------------------
uint32_t choose_ot(uint32_t outs)
{
    uint32_t best = 1;
    uint64_t best_cost = ~uint64_t(0);
    for (uint32_t ot = 1; ot <= 64; ot <<= 1)
    {
        uint64_t cost = uint64_t((outs + ot - 1) / ot) * (ot + 1);
        if (cost <= best_cost) { best_cost = cost; best = ot; }
    }
    return best;
}

void plan_geometry(hb_conv *c)
{
    Geom &g = c->g;
    const uint32_t cpv = c->dtype == HB_F64 ? 1 : 2;
    g.groups = c->groups; g.ins = c->ins; g.outs = c->outs;
    g.log2n = (uint32_t) c->fft_log2;
    g.B = 1u << (g.log2n - 1);
    g.Pcap = (uint32_t) (c->max_length / g.B);
    g.P = c->P;
    g.OT = choose_ot(c->outs);
    g.n_ot = (c->outs + g.OT - 1) / g.OT;
    const uint32_t xq = g.B / cpv;
    g.TBV = std::min<uint32_t>(xq, std::max<uint32_t>(2048u / g.OT, 1u));
    g.n_bt = xq / g.TBV;
    g.TX = std::min<uint32_t>(g.TBV, 256u);
    g.XA = g.TBV / g.TX;
    g.TY = std::min<uint32_t>(g.OT, 256u / g.TX);
    g.OB = g.OT / g.TY;
    g.Q = g.OT * g.TBV;
    g.upt = g.ins * g.P;
    g.tiles = g.groups * g.n_ot * g.n_bt;
    g.U = uint64_t(g.tiles) * g.upt;
    uint64_t want = uint64_t(c->sm_count) * (c->variant == 1 ? 1 : std::max(1, c->ctas_per_sm));
    g.G = (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(g.U, want));
    // TMA ring depth: as many stages as fit in ~200 KB, between 2 and 6
    const size_t stage = size_t(g.Q + g.TBV) * 16;
    int st = (int) std::min<size_t>(6, (200 * 1024) / stage);
    c->nstages = std::max(2, st);
    c->cmac_smem = size_t(c->nstages) * stage + size_t(c->nstages) * 8;
}

size_t h_vectors(const hb_conv *c)
{
    // groups * (n_ot*OT) padded rows * ins * max_length bins, in 16-byte vectors
    const uint32_t ot = choose_ot(c->outs);
    const uint32_t rows = ((c->outs + ot - 1) / ot) * ot;
    const size_t cpv = c->dtype == HB_F64 ? 1 : 2;
    return size_t(c->groups) * rows * c->ins * c->max_length / cpv;
}

void free_device(hb_conv *c)
{
    cudaFree(c->d_H); cudaFree(c->d_X); cudaFree(c->d_Hnyq); cudaFree(c->d_Xnyq); cudaFree(c->d_tw);
    c->d_H = c->d_X = c->d_Hnyq = c->d_Xnyq = c->d_tw = nullptr;
    c->d_S.release();
    for (int k = 0; k < 2; k++) { c->d_xin[k].release(); c->d_yout[k].release(); }
    c->d_io_in.release(); c->d_io_out.release(); c->d_ir.release();
    c->h_in.release(); c->h_out.release(); c->h_ir.release();
}

// allocate everything whose size depends on max_length (ctor and resize)
int alloc_capacity(hb_conv *c)
{
    cudaFree(c->d_H); cudaFree(c->d_X); cudaFree(c->d_Hnyq); cudaFree(c->d_Xnyq);
    c->d_H = c->d_X = c->d_Hnyq = c->d_Xnyq = nullptr;
    const size_t hbytes = h_vectors(c) * 16;
    const size_t xbytes = size_t(c->groups) * c->ins * c->max_length * 2 * c->esize();
    // Nyquist side arrays are sized for the smallest hop the object may be switched to
    const size_t minB = size_t(1) << (MIN_FFT_LOG2 - 1);
    const size_t pmax = c->max_length / minB;
    if (cudaMalloc(&c->d_H, std::max<size_t>(hbytes, 16)) != cudaSuccess ||
        cudaMalloc(&c->d_X, std::max<size_t>(xbytes, 16)) != cudaSuccess ||
        cudaMalloc(&c->d_Hnyq, std::max<size_t>(c->pairs() * pmax * c->esize(), 16)) != cudaSuccess ||
        cudaMalloc(&c->d_Xnyq, std::max<size_t>(size_t(c->groups) * c->ins * pmax * c->esize(), 16)) != cudaSuccess)
    {
        set_error("device allocation failed for %zu taps per pair (%s)", (size_t) c->max_length, cudaGetErrorString(cudaGetLastError()));
        return ERR_MEM_UNAVAILABLE;
    }
    // padded rows and never-set pairs must read as silence
    HB_CUDA(cudaMemsetAsync(c->d_H, 0, std::max<size_t>(hbytes, 16), c->stream));
    HB_CUDA(cudaMemsetAsync(c->d_Hnyq, 0, std::max<size_t>(c->pairs() * pmax * c->esize(), 16), c->stream));
    std::fill(c->nparts.begin(), c->nparts.end(), 0u);
    c->P = 0;
    c->need_reset = true;
    return ERR_NONE;
}

// opt in to > 48 KB of dynamic shared memory once per (kernel, device, size)
template <class K> int allow_smem(K kernel, size_t bytes)
{
    if (bytes <= 48 * 1024) return HB_OK;
    static std::mutex m;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(m);
    size_t &have = done[std::make_pair((const void *) kernel, dev)];
    if (have >= bytes) return HB_OK;
    HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    have = bytes;
    return HB_OK;
}

template <class T> size_t fft_smem(uint32_t log2m) { return size_t(padded_elems<HB_PADSH>(1u << log2m)) * sizeof(Cx<T>); }

// ---- kernel dispatch ------------------------------------------------------------------------------
template <class T, int XA, int OB>
int launch_cmac_inst(hb_conv *c, cudaStream_t st)
{
    typedef typename VecOf<T>::type V;
    const Geom &g = c->g;
    if (c->variant == 1)
    {
        int rc = allow_smem(k_cmac_tma<T, XA, OB>, c->cmac_smem);
        if (rc) return rc;
        k_cmac_tma<T, XA, OB><<<g.G, 256, c->cmac_smem, st>>>(g, (const V *) c->d_H, (const V *) c->d_X, (V *) c->d_S.p, c->nstages);
    }
    else
        k_cmac_ldg<T, XA, OB><<<g.G, 256, 0, st>>>(g, (const V *) c->d_H, (const V *) c->d_X, (V *) c->d_S.p);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int launch_cmac(hb_conv *c, cudaStream_t st)
{
    const uint32_t key = c->g.XA * 16 + c->g.OB;
    switch (key)
    {
        case 1 * 16 + 1: return launch_cmac_inst<T, 1, 1>(c, st);
        case 1 * 16 + 2: return launch_cmac_inst<T, 1, 2>(c, st);
        case 1 * 16 + 4: return launch_cmac_inst<T, 1, 4>(c, st);
        case 1 * 16 + 8: return launch_cmac_inst<T, 1, 8>(c, st);
        case 2 * 16 + 1: return launch_cmac_inst<T, 2, 1>(c, st);
        case 2 * 16 + 2: return launch_cmac_inst<T, 2, 2>(c, st);
        case 2 * 16 + 4: return launch_cmac_inst<T, 2, 4>(c, st);
        case 4 * 16 + 1: return launch_cmac_inst<T, 4, 1>(c, st);
        case 4 * 16 + 2: return launch_cmac_inst<T, 4, 2>(c, st);
        case 8 * 16 + 1: return launch_cmac_inst<T, 8, 1>(c, st);
    }
    set_error("internal: no multiply-accumulate kernel for XA=%u OB=%u", c->g.XA, c->g.OB);
    return HB_ERR_UNSUPPORTED;
}

template <class T>
int launch_fwd(hb_conv *c, const T *xin, size_t ld, size_t off, cudaStream_t st)
{
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const size_t smem = fft_smem<T>(log2m);
    const unsigned grid = g.groups * g.ins;
    if ((1u << log2m) / 8 > 1024)
    {
        int rc = allow_smem(k_fwd<T, 16>, smem); if (rc) return rc;
        k_fwd<T, 16><<<grid, fft_threads(log2m, 16), smem, st>>>(g, xin, ld, off, (Cx<T> *) c->d_X, (T *) c->d_Xnyq, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    else
    {
        int rc = allow_smem(k_fwd<T, 8>, smem); if (rc) return rc;
        k_fwd<T, 8><<<grid, fft_threads(log2m, 8), smem, st>>>(g, xin, ld, off, (Cx<T> *) c->d_X, (T *) c->d_Xnyq, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int launch_inv(hb_conv *c, T *yout, size_t ld, size_t off, cudaStream_t st)
{
    typedef typename VecOf<T>::type V;
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const size_t smem = fft_smem<T>(log2m);
    const unsigned grid = g.groups * g.outs;
    if ((1u << log2m) / 8 > 1024)
    {
        int rc = allow_smem(k_inv<T, 16>, smem); if (rc) return rc;
        k_inv<T, 16><<<grid, fft_threads(log2m, 16), smem, st>>>(g, (const V *) c->d_S.p, (const T *) c->d_Xnyq, (const T *) c->d_Hnyq, yout, ld, off, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    else
    {
        int rc = allow_smem(k_inv<T, 8>, smem); if (rc) return rc;
        k_inv<T, 8><<<grid, fft_threads(log2m, 8), smem, st>>>(g, (const V *) c->d_S.p, (const T *) c->d_Xnyq, (const T *) c->d_Hnyq, yout, ld, off, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int launch_ir(hb_conv *c, const T *d_ir, size_t taps, uint32_t grp, uint32_t in, uint32_t o, uint32_t nwrite, cudaStream_t st)
{
    const Geom &g = c->g;
    const uint32_t log2m = g.log2n - 1;
    const size_t smem = fft_smem<T>(log2m);
    if (!nwrite) return HB_OK;
    if ((1u << log2m) / 8 > 1024)
    {
        int rc = allow_smem(k_ir<T, 16>, smem); if (rc) return rc;
        k_ir<T, 16><<<nwrite, fft_threads(log2m, 16), smem, st>>>(g, d_ir, taps, grp, in, o, (Cx<T> *) c->d_H, (T *) c->d_Hnyq, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    else
    {
        int rc = allow_smem(k_ir<T, 8>, smem); if (rc) return rc;
        k_ir<T, 8><<<nwrite, fft_threads(log2m, 8), smem, st>>>(g, d_ir, taps, grp, in, o, (Cx<T> *) c->d_H, (T *) c->d_Hnyq, (const Cx<T> *) c->d_tw, c->tw_log2);
    }
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
int launch_rows(T *dst, size_t dld, const T *src, size_t sld, size_t n, size_t rows, int add, cudaStream_t st)
{
    if (!n || !rows) return HB_OK;
    dim3 grid((unsigned) std::min<size_t>((n + 255) / 256, 64), (unsigned) rows);
    k_rows<T><<<grid, 256, 0, st>>>(dst, dld, src, sld, n, add);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

bool fft_supported(const hb_conv *c, uintptr_t log2n)
{
    const int lim = c->dtype == HB_F64 ? SmemFftLimit<double>::max_log2m : SmemFftLimit<float>::max_log2m;
    return (int) log2n - 1 <= lim;
}

// setFFTSize semantics (PartitionedConvolve.cpp:131-154)
int apply_fft_size(hb_conv *c, uintptr_t fft_size)
{
    const uintptr_t l2 = ceil_log2(fft_size);
    int error = ERR_NONE;
    if (l2 < MIN_FFT_LOG2 || l2 > c->max_fft_log2) return ERR_FFT_SIZE_OUT_OF_RANGE;
    if (fft_size != (uintptr_t(1) << l2)) error = ERR_FFT_SIZE_NON_POWER_OF_TWO;
    if (l2 != c->fft_log2)
    {
        std::fill(c->nparts.begin(), c->nparts.end(), 0u);
        c->P = 0;
        c->fft_log2 = l2;
        c->need_reset = true;
    }
    return error;
}

// make staging rows hold at least `need_x` / `need_y` elements, keeping the retained heads
template <class T>
int ensure_staging(hb_conv *c, size_t need_x, size_t need_y, bool preserve, cudaStream_t st)
{
    const size_t B = c->g.B;
    if (need_x > c->xin_ld)
    {
        const size_t ld = ((need_x * 3 / 2) + 63) & ~size_t(63);
        const size_t rows = size_t(c->groups) * c->ins;
        DevBuf n0, n1;
        int rc;
        if ((rc = n0.ensure(rows * ld * sizeof(T))) || (rc = n1.ensure(rows * ld * sizeof(T)))) return rc;
        if (c->xin_ld && preserve)
        {
            rc = launch_rows<T>((T *) n0.p, ld, (const T *) c->d_xin[c->cur].p + c->x_tail, c->xin_ld, B + c->rw, rows, 0, st);
            if (rc) return rc;
            HB_CUDA(cudaStreamSynchronize(st));
        }
        else
            HB_CUDA(cudaMemsetAsync(n0.p, 0, rows * ld * sizeof(T), st));
        c->d_xin[0].release(); c->d_xin[1].release();
        c->d_xin[0] = n0; c->d_xin[1] = n1;
        if (c->cur == 1) std::swap(c->d_xin[0], c->d_xin[1]);
        c->xin_ld = ld;
        c->x_tail = 0;
    }
    if (need_y > c->yout_ld)
    {
        const size_t ld = ((need_y * 3 / 2) + 63) & ~size_t(63);
        const size_t rows = size_t(c->groups) * c->outs;
        DevBuf n0, n1;
        int rc;
        if ((rc = n0.ensure(rows * ld * sizeof(T))) || (rc = n1.ensure(rows * ld * sizeof(T)))) return rc;
        if (c->yout_ld && preserve)
        {
            rc = launch_rows<T>((T *) n0.p, ld, (const T *) c->d_yout[c->cur].p + c->y_tail, c->yout_ld, B, rows, 0, st);
            if (rc) return rc;
            HB_CUDA(cudaStreamSynchronize(st));
        }
        else
            HB_CUDA(cudaMemsetAsync(n0.p, 0, rows * ld * sizeof(T), st));
        c->d_yout[0].release(); c->d_yout[1].release();
        c->d_yout[0] = n0; c->d_yout[1] = n1;
        if (c->cur == 1) std::swap(c->d_yout[0], c->d_yout[1]);
        c->yout_ld = ld;
        c->y_tail = 0;
    }
    return HB_OK;
}

// the reset block of process (PartitionedConvolve.cpp:267-290)
template <class T>
int do_reset(hb_conv *c, cudaStream_t st)
{
    plan_geometry(c);
    const Geom &g = c->g;
    int rc = c->d_S.ensure(std::max<size_t>((size_t(g.G) + g.tiles) * g.Q * 16, 16));
    if (rc) return rc;
    // FDL silence: stale slots are masked in the reference by mValidPartitions (:285,373); zeros do the same
    HB_CUDA(cudaMemsetAsync(c->d_X, 0, size_t(g.groups) * g.ins * g.P * g.B * 2 * sizeof(T), st));
    HB_CUDA(cudaMemsetAsync(c->d_Xnyq, 0, size_t(g.groups) * g.ins * g.P * sizeof(T), st));
    c->rw = c->reset_offset < 0 ? 0 : uintptr_t(c->reset_offset) % g.B;
    c->g.slot = 0;
    // staging rows start as silence: previous hop, pending samples and previous result block
    if (c->xin_ld) HB_CUDA(cudaMemsetAsync(c->d_xin[c->cur].p, 0, size_t(g.groups) * g.ins * c->xin_ld * sizeof(T), st));
    if (c->yout_ld) HB_CUDA(cudaMemsetAsync(c->d_yout[c->cur].p, 0, size_t(g.groups) * g.outs * c->yout_ld * sizeof(T), st));
    c->x_tail = c->y_tail = 0;
    c->need_reset = false;
    return HB_OK;
}

// fold the recorded event intervals into prof_ms (synchronises on the last recorded event)
int drain_profile(hb_conv *c)
{
    if (!c->ev_used) return HB_OK;
    HB_CUDA(cudaEventSynchronize(c->ev[c->ev_used - 1]));
    for (size_t k = 0; k + 3 < c->ev_used; k += 4)
    {
        for (int j = 0; j < 3; j++)
        {
            float ms = 0.f;
            HB_CUDA(cudaEventElapsedTime(&ms, c->ev[k + j], c->ev[k + j + 1]));
            c->prof_ms[j] += ms;
        }
        c->prof_hops++;
    }
    c->ev_used = 0;
    return HB_OK;
}

int profile_mark(hb_conv *c, cudaStream_t st)
{
    if (c->ev_used == c->ev.size())
    {
        if (c->ev.size() >= 4 * 256) { int rc = drain_profile(c); if (rc) return rc; }
        else
            for (int k = 0; k < 4; k++)
            {
                cudaEvent_t e;
                HB_CUDA(cudaEventCreate(&e));
                c->ev.push_back(e);
            }
    }
    HB_CUDA(cudaEventRecord(c->ev[c->ev_used++], st));
    return HB_OK;
}

// core of process: device rows in, device rows out, everything enqueued on st
template <class T>
int process_core(hb_conv *c, const T *d_in, size_t in_ld, T *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    if (!c->P) return HB_ERR_NO_IR;
    if (!n) return HB_OK;
    int rc;
    if (c->need_reset)
    {
        plan_geometry(c);
        if ((rc = ensure_staging<T>(c, 2 * size_t(c->g.B) + n, 2 * size_t(c->g.B) + n, false, st))) return rc;
        if ((rc = do_reset<T>(c, st))) return rc;
    }
    const size_t B = c->g.B;
    const size_t rows_in = size_t(c->groups) * c->ins, rows_out = size_t(c->groups) * c->outs;
    const size_t rw = c->rw;
    const size_t nh = (rw + n) / B;
    if ((rc = ensure_staging<T>(c, B + rw + n, (nh + 1) * B, true, st))) return rc;

    // bring the retained heads to offset 0 of the other row set when they sit further in
    if (c->x_tail || c->y_tail)
    {
        const int nxt = c->cur ^ 1;
        if ((rc = launch_rows<T>((T *) c->d_xin[nxt].p, c->xin_ld, (const T *) c->d_xin[c->cur].p + c->x_tail, c->xin_ld, B + rw, rows_in, 0, st))) return rc;
        if ((rc = launch_rows<T>((T *) c->d_yout[nxt].p, c->yout_ld, (const T *) c->d_yout[c->cur].p + c->y_tail, c->yout_ld, B, rows_out, 0, st))) return rc;
        c->cur = nxt;
        c->x_tail = c->y_tail = 0;
    }
    T *xin = (T *) c->d_xin[c->cur].p;
    T *yout = (T *) c->d_yout[c->cur].p;
    if ((rc = launch_rows<T>(xin + B + rw, c->xin_ld, d_in, in_ld, n, rows_in, 0, st))) return rc;

    for (size_t h = 0; h < nh; h++)
    {
        // newest spectrum goes one slot below the previous one (mInputPosition--, cpp:374)
        c->g.slot = c->g.slot ? c->g.slot - 1 : c->g.P - 1;
        const bool prof = c->profiling;
        if (prof && (rc = profile_mark(c, st))) return rc;
        if ((rc = launch_fwd<T>(c, xin, c->xin_ld, h * B, st))) return rc;
        if (prof && (rc = profile_mark(c, st))) return rc;
        if ((rc = launch_cmac<T>(c, st))) return rc;
        if (prof && (rc = profile_mark(c, st))) return rc;
        if ((rc = launch_inv<T>(c, yout, c->yout_ld, (h + 1) * B, st))) return rc;
        if (prof && (rc = profile_mark(c, st))) return rc;
    }
    if ((rc = launch_rows<T>(d_out, out_ld, yout + rw, c->yout_ld, n, rows_out, accumulate, st))) return rc;

    c->rw = (rw + n) - nh * B;
    c->x_tail = nh * B;
    c->y_tail = nh * B;
    return HB_OK;
}

template <class T>
int set_ir_core(hb_conv *c, uint32_t grp, uint32_t in, uint32_t o, const T *d_ir, uintptr_t length, cudaStream_t st)
{
    // length clipping of PartitionedConvolve::set (cpp:192-199); d_ir already points at tap `offset`
    int error = ERR_NONE;
    if (c->length && c->length < length) length = c->length;
    if (length > c->max_length) { length = c->max_length; error = ERR_MEM_ALLOC_TOO_SMALL; }
    plan_geometry(c);
    const uint32_t B = c->g.B;
    const uint32_t np = (uint32_t) ((length + B - 1) / B);
    const size_t pair = (size_t(grp) * c->outs + o) * c->ins + in;
    const uint32_t nwrite = std::max(np, c->nparts[pair]);
    int rc = launch_ir<T>(c, d_ir, length, grp, in, o, nwrite, st);
    if (rc) return rc;
    c->nparts[pair] = np;
    c->P = *std::max_element(c->nparts.begin(), c->nparts.end());
    c->need_reset = true;
    return error;
}

int check_handle(hb_conv *c)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    return use_device(c->device);
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int hb_conv_create(hb_conv **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs,
                              uintptr_t max_fft_size, uintptr_t max_length, uintptr_t offset, uintptr_t length, int device)
{
    if (!out || (dtype != HB_F32 && dtype != HB_F64) || !groups || !ins || !outs) { set_error("hb_conv_create: bad argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = use_device(device);
    if (rc) return rc;

    // setMaxFFTSize (PartitionedConvolve.cpp:26-50)
    int ctor_error = ERR_NONE;
    uintptr_t l2 = ceil_log2(max_fft_size);
    if (l2 > MAX_FFT_LOG2) { ctor_error = ERR_FFT_SIZE_MAX_TOO_LARGE; l2 = MAX_FFT_LOG2; }
    if (l2 && l2 < MIN_FFT_LOG2) { ctor_error = ERR_FFT_SIZE_MAX_TOO_SMALL; l2 = MIN_FFT_LOG2; }
    if (max_fft_size != (uintptr_t(1) << l2)) ctor_error = ERR_FFT_SIZE_MAX_NON_POWER_OF_TWO;
    if (l2 < MIN_FFT_LOG2) l2 = MIN_FFT_LOG2;

    hb_conv *c = new hb_conv;
    c->dtype = dtype; c->device = device; c->groups = groups; c->ins = ins; c->outs = outs;
    c->max_fft_log2 = l2;
    if (!fft_supported(c, l2))
    {
        set_error("FFT size 2^%d is beyond the shared-memory FFT of this build (max 2^%d for this dtype)", (int) l2,
                  (dtype == HB_F64 ? SmemFftLimit<double>::max_log2m : SmemFftLimit<float>::max_log2m) + 1);
        delete c;
        return HB_ERR_UNSUPPORTED;
    }
    c->fft_log2 = l2;
    c->offset = offset;
    const uintptr_t maxB = (uintptr_t(1) << l2) >> 1;
    uintptr_t ml = max_length ? max_length : maxB;
    if (ml % maxB) ml = (ml / maxB + 1) * maxB;                 // cpp:77-82
    c->max_length = ml;
    c->length = std::min(length, ml);                           // setLength (cpp:156-161)
    c->nparts.assign(c->pairs(), 0u);

    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete c; return HB_ERR_CUDA; }
    c->tw_log2 = (int) l2;
    rc = make_twiddles(dtype, c->tw_log2, &c->d_tw);
    if (rc == HB_OK) rc = alloc_capacity(c);
    if (rc != HB_OK && rc != ERR_NONE)
    {
        free_device(c);
        cudaStreamDestroy(c->stream);
        delete c;
        return rc;
    }
    plan_geometry(c);
    *out = c;
    return ctor_error;
}

extern "C" void hb_conv_destroy(hb_conv *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_device(c);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int hb_conv_set_fft_size(hb_conv *c, uintptr_t fft_size)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    return apply_fft_size(c, fft_size);
}

extern "C" int hb_conv_set_length(hb_conv *c, uintptr_t length)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->length = std::min(length, c->max_length);
    return length > c->max_length ? ERR_PARTITION_LENGTH_TOO_LARGE : ERR_NONE;
}

extern "C" int hb_conv_set_offset(hb_conv *c, uintptr_t offset)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->offset = offset;
    return ERR_NONE;
}

extern "C" int hb_conv_set_reset_offset(hb_conv *c, intptr_t offset)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->reset_offset = offset;
    return ERR_NONE;
}

extern "C" int hb_conv_set_ir(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (group >= c->groups || in >= c->ins || out >= c->outs || (ir_dtype != HB_F32 && ir_dtype != HB_F64)) { set_error("hb_conv_set_ir: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    // cpp:192: nothing to load when the IR ends before the offset
    uintptr_t len = (!ir || length <= c->offset) ? 0 : length - c->offset;
    uintptr_t take = len;
    if (c->length && c->length < take) take = c->length;
    if (take > c->max_length) take = c->max_length;
    const size_t es = c->esize();
    if (take)
    {
        if ((rc = c->h_ir.ensure(take * es)) || (rc = c->d_ir.ensure(take * es))) return rc;
        // the previous upload from the pinned buffer must have landed before it is overwritten
        HB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->dtype == HB_F64)
        {
            double *dst = (double *) c->h_ir.p;
            if (ir_dtype == HB_F64) memcpy(dst, (const double *) ir + c->offset, take * sizeof(double));
            else for (size_t k = 0; k < take; k++) dst[k] = (double) ((const float *) ir)[c->offset + k];
        }
        else
        {
            float *dst = (float *) c->h_ir.p;
            if (ir_dtype == HB_F32) memcpy(dst, (const float *) ir + c->offset, take * sizeof(float));
            else for (size_t k = 0; k < take; k++) dst[k] = (float) ((const double *) ir)[c->offset + k];   // Convolver.cpp:126-134
        }
        HB_CUDA(cudaMemcpyAsync(c->d_ir.p, c->h_ir.p, take * es, cudaMemcpyHostToDevice, c->stream));
    }
    rc = c->dtype == HB_F64 ? set_ir_core<double>(c, group, in, out, (const double *) c->d_ir.p, len, c->stream)
                            : set_ir_core<float>(c, group, in, out, (const float *) c->d_ir.p, len, c->stream);
    if (rc < 0) return rc;
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

extern "C" int hb_conv_set_ir_dev(hb_conv *c, uint32_t group, uint32_t in, uint32_t out, const void *d_ir, uintptr_t length)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if (group >= c->groups || in >= c->ins || out >= c->outs) { set_error("hb_conv_set_ir_dev: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    uintptr_t len = (!d_ir || length <= c->offset) ? 0 : length - c->offset;
    const char *p = (const char *) d_ir + (len ? c->offset * c->esize() : 0);
    return c->dtype == HB_F64 ? set_ir_core<double>(c, group, in, out, (const double *) p, len, c->stream)
                              : set_ir_core<float>(c, group, in, out, (const float *) p, len, c->stream);
}

extern "C" int hb_conv_resize(hb_conv *c, uintptr_t max_length)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    HB_CUDA(cudaStreamSynchronize(c->stream));
    const uintptr_t maxB = (uintptr_t(1) << c->max_fft_log2) >> 1;
    uintptr_t ml = max_length ? max_length : maxB;
    if (ml % maxB) ml = (ml / maxB + 1) * maxB;
    if (ml == c->max_length) return ERR_NONE;

    // keep what is loaded: same tiling, only the partition stride (Pcap) of the layout changes
    plan_geometry(c);
    const Geom old = c->g;
    void *oldH = c->d_H, *oldHn = c->d_Hnyq;
    const std::vector<uint32_t> old_parts = c->nparts;
    c->d_H = c->d_Hnyq = nullptr;
    c->max_length = ml;
    if (c->length > ml) c->length = ml;
    rc = alloc_capacity(c);
    if (rc)
    {
        cudaFree(oldH); cudaFree(oldHn);
        return rc;
    }
    plan_geometry(c);
    const Geom &now = c->g;
    const uint32_t keep = std::min(old.Pcap, now.Pcap);
    if (keep && oldH)
    {
        const size_t unit = size_t(old.Q) * 16;
        HB_CUDA(cudaMemcpy2DAsync(c->d_H, size_t(now.Pcap) * unit, oldH, size_t(old.Pcap) * unit, size_t(keep) * unit,
                                  size_t(old.tiles) * old.ins, cudaMemcpyDeviceToDevice, c->stream));
        HB_CUDA(cudaMemcpy2DAsync(c->d_Hnyq, size_t(now.Pcap) * c->esize(), oldHn, size_t(old.Pcap) * c->esize(), size_t(keep) * c->esize(),
                                  c->pairs(), cudaMemcpyDeviceToDevice, c->stream));
        for (size_t k = 0; k < c->nparts.size(); k++) c->nparts[k] = std::min(old_parts[k], keep);
        c->P = *std::max_element(c->nparts.begin(), c->nparts.end());
    }
    HB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(oldH); cudaFree(oldHn);
    c->need_reset = true;
    return ERR_NONE;
}

extern "C" int hb_conv_reset(hb_conv *c)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->need_reset = true;
    return ERR_NONE;
}

extern "C" uintptr_t hb_conv_partitions(const hb_conv *c) { return c ? c->P : 0; }
extern "C" uintptr_t hb_conv_max_length(const hb_conv *c) { return c ? c->max_length : 0; }
extern "C" uintptr_t hb_conv_fft_size(const hb_conv *c) { return c ? (uintptr_t(1) << c->fft_log2) : 0; }

extern "C" int hb_conv_process_dev(hb_conv *c, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld,
                                   uintptr_t num_samples, int accumulate, void *stream)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if ((!d_in || !d_out) && num_samples) { set_error("hb_conv_process_dev: null buffer"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    cudaStream_t st = stream ? (cudaStream_t) stream : c->stream;
    return c->dtype == HB_F64 ? process_core<double>(c, (const double *) d_in, in_ld, (double *) d_out, out_ld, num_samples, accumulate, st)
                              : process_core<float>(c, (const float *) d_in, in_ld, (float *) d_out, out_ld, num_samples, accumulate, st);
}

extern "C" int hb_conv_process(hb_conv *c, const void *const *ins, void *const *outs, uintptr_t n, int accumulate)
{
    int rc = check_handle(c);
    if (rc) return rc;
    if ((!ins || !outs) && n) { set_error("hb_conv_process: null buffer"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    if (!c->P) return HB_ERR_NO_IR;
    if (!n) return HB_OK;
    const size_t es = c->esize();
    const size_t rows_in = size_t(c->groups) * c->ins, rows_out = size_t(c->groups) * c->outs;
    if ((rc = c->h_in.ensure(rows_in * n * es)) || (rc = c->h_out.ensure(rows_out * n * es)) ||
        (rc = c->d_io_in.ensure(rows_in * n * es)) || (rc = c->d_io_out.ensure(rows_out * n * es))) return rc;
    for (size_t r = 0; r < rows_in; r++)
    {
        if (!ins[r]) { set_error("hb_conv_process: null input row %zu", r); return HB_ERR_BAD_ARG; }
        memcpy((char *) c->h_in.p + r * n * es, ins[r], n * es);
    }
    HB_CUDA(cudaMemcpyAsync(c->d_io_in.p, c->h_in.p, rows_in * n * es, cudaMemcpyHostToDevice, c->stream));
    rc = c->dtype == HB_F64 ? process_core<double>(c, (const double *) c->d_io_in.p, n, (double *) c->d_io_out.p, n, n, 0, c->stream)
                            : process_core<float>(c, (const float *) c->d_io_in.p, n, (float *) c->d_io_out.p, n, n, 0, c->stream);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(c->h_out.p, c->d_io_out.p, rows_out * n * es, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t r = 0; r < rows_out; r++)
    {
        if (!outs[r]) continue;
        if (!accumulate) memcpy(outs[r], (char *) c->h_out.p + r * n * es, n * es);
        else if (c->dtype == HB_F64)
        {
            double *d = (double *) outs[r];
            const double *s = (const double *) c->h_out.p + r * n;
            for (size_t k = 0; k < n; k++) d[k] += s[k];                   // MonoConvolve.cpp:167-177
        }
        else
        {
            float *d = (float *) outs[r];
            const float *s = (const float *) c->h_out.p + r * n;
            for (size_t k = 0; k < n; k++) d[k] += s[k];
        }
    }
    return HB_OK;
}

extern "C" int hb_conv_set_tuning(hb_conv *c, int ctas_per_sm, int variant)
{
    if (!c) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (variant != 0 && variant != 1) { set_error("variant must be 0 (direct loads) or 1 (TMA ring)"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(c->lock);
    c->ctas_per_sm = ctas_per_sm > 0 ? ctas_per_sm : (variant == 1 ? 1 : 2);
    c->variant = variant;
    c->need_reset = true;           // partial-segment geometry depends on the grid
    return HB_OK;
}

extern "C" int hb_conv_set_profiling(hb_conv *c, int enable)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    if ((rc = drain_profile(c))) return rc;
    c->profiling = enable != 0;
    c->prof_ms[0] = c->prof_ms[1] = c->prof_ms[2] = 0;
    c->prof_hops = 0;
    return HB_OK;
}

extern "C" int hb_conv_get_profile(hb_conv *c, double *ms_forward, double *ms_cmac, double *ms_inverse, uint64_t *hops)
{
    int rc = check_handle(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> g(c->lock);
    if ((rc = drain_profile(c))) return rc;
    if (ms_forward) *ms_forward = c->prof_ms[0];
    if (ms_cmac) *ms_cmac = c->prof_ms[1];
    if (ms_inverse) *ms_inverse = c->prof_ms[2];
    if (hops) *hops = c->prof_hops;
    return HB_OK;
}

extern "C" uint64_t hb_conv_bytes_per_hop(const hb_conv *c)
{
    if (!c || !c->P) return 0;
    // 2sB*P*K (IR spectra) + 2sB*P*I (FDL) + sB*(I+O), K = pairs, per group
    const uint64_t s = c->esize(), B = (uint64_t(1) << c->fft_log2) >> 1, P = c->P;
    const uint64_t I = uint64_t(c->groups) * c->ins, O = uint64_t(c->groups) * c->outs, K = c->pairs();
    return 2 * s * B * P * K + 2 * s * B * P * I + s * B * (I + O);
}
