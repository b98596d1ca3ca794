// hb_matrix.cu -- the non-uniform partition scheme of the reference's MonoConvolve lifted to a whole
// channel matrix: the hb_matrix_* entry points of include/hisstools_b200.h.
//
// The reference builds an N x M Convolver out of N*M MonoConvolve objects, each of which owns a
// time-domain head and up to four PartitionedConvolve parts (MonoConvolve.cpp:203-258,
// NToMonoConvolve.cpp:4-9, Convolver.cpp:5-41).  Here one hb_matrix owns ONE uniform engine
// (hb_conv, hb_conv.cu) per part of the scheme, each holding that part of every pair of the matrix,
// plus one direct-form head for the zero-latency modes.  Inputs cross PCIe once per call, every
// part accumulates into the same device-resident output rows, outputs cross back once.
#include "hb_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

using namespace hb;

namespace
{
constexpr uint32_t TD_MAX_TAPS = 2044;      // TimeDomainConvolve.cpp:62-67
constexpr int TD_BLOCK = 256;

// Direct-form zero-latency head (behaviour of TimeDomainConvolve.cpp:69-163 for every pair at once):
//   out[g][o][s] (+)= sum_i sum_{k < T} h[g][o][i][k] * x[g][i][s - k]
// x for negative sample indices comes from `hist` (the last T samples of the previous calls).
// grid = (ceil(n / TD_BLOCK), groups * outs); shared memory: T taps + (T + TD_BLOCK) samples.
template <class T>
__global__ void __launch_bounds__(TD_BLOCK) k_td(const T *__restrict__ h, const T *__restrict__ x, size_t x_ld,
                                                 const T *__restrict__ hist, T *__restrict__ out, size_t out_ld,
                                                 uint32_t ins, uint32_t outs, uint32_t taps, size_t n, int add)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sh = reinterpret_cast<T *>(smem_raw);
    T *sx = sh + taps;
    const uint32_t row = blockIdx.y;
    const uint32_t grp = row / outs;
    const size_t s0 = size_t(blockIdx.x) * TD_BLOCK;
    const size_t s = s0 + threadIdx.x;
    T acc = T(0);
    for (uint32_t i = 0; i < ins; i++)
    {
        const T *hp = h + (size_t(row) * ins + i) * taps;
        const size_t xrow = size_t(grp) * ins + i;
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < taps; k += TD_BLOCK) sh[k] = hp[k];
        // window [s0 - taps, s0 + TD_BLOCK)
        for (uint32_t j = threadIdx.x; j < taps + TD_BLOCK; j += TD_BLOCK)
        {
            const long long idx = (long long) s0 + (long long) j - (long long) taps;
            T v = T(0);
            if (idx < 0) v = hist[xrow * taps + size_t(idx + (long long) taps)];
            else if (size_t(idx) < n) v = x[xrow * x_ld + size_t(idx)];
            sx[j] = v;
        }
        __syncthreads();
        const T *w = sx + threadIdx.x + taps;       // w[-k] = x[s - k]
#pragma unroll 4
        for (uint32_t k = 0; k < taps; k++) acc = fma(sh[k], w[-(int) k], acc);
    }
    if (s < n)
    {
        T *o = out + size_t(row) * out_ld + s;
        *o = add ? *o + acc : acc;
    }
}

// Register-blocked form for long calls: every thread produces TD_R consecutive output samples from a sliding register
// window of the input, four taps per step (one broadcast LDS.128 of taps + one LDS.128 of new samples per 4 * TD_R FMAs), so
// the loop is bound by the FMA pipe instead of shared-memory loads.  grid = (ceil(n / (TDB * TD_R)), groups * outs);
// shared memory: taps + (taps + TDB * TD_R) samples; taps must be a multiple of 4 (it is A / 2 with A >= 32).
// four consecutive values from a 16-byte aligned shared-memory address
__device__ __forceinline__ void ld4(const float *p, float *v)
{
    const float4 q = *reinterpret_cast<const float4 *>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void ld4(const double *p, double *v)
{
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
constexpr int TD_R = 8;
constexpr int TDB = 128;            // threads of the blocked kernel: several small CTAs per SM overlap their staging and arithmetic
template <class T>
__global__ void __launch_bounds__(TDB) k_td_blocked(const T *__restrict__ h, const T *__restrict__ x, size_t x_ld,
                                                         const T *__restrict__ hist, T *__restrict__ out, size_t out_ld,
                                                         uint32_t ins, uint32_t outs, uint32_t taps, size_t n, int add)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sh = reinterpret_cast<T *>(smem_raw);
    T *sx = sh + taps;
    constexpr uint32_t SPAN = TDB * TD_R;
    const uint32_t row = blockIdx.y;
    const uint32_t grp = row / outs;
    const size_t s0 = size_t(blockIdx.x) * SPAN;
    T acc[TD_R];
#pragma unroll
    for (int r = 0; r < TD_R; r++) acc[r] = T(0);
    const uint32_t p = taps + TD_R * threadIdx.x;          // sx index of x[s], s = first sample of this thread
    for (uint32_t i = 0; i < ins; i++)
    {
        const T *hp = h + (size_t(row) * ins + i) * taps;
        const size_t xrow = size_t(grp) * ins + i;
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < taps; k += TDB) sh[k] = hp[k];
        for (uint32_t j = threadIdx.x; j < taps + SPAN; j += TDB)
        {
            const long long idx = (long long) s0 + (long long) j - (long long) taps;
            T v = T(0);
            if (idx < 0) v = hist[xrow * taps + size_t(idx + (long long) taps)];
            else if (size_t(idx) < n) v = x[xrow * x_ld + size_t(idx)];
            sx[j] = v;
        }
        __syncthreads();
        // win[q] = sx[p - k - 4 + q] (q < 12; 16-byte aligned groups of four): tap k + t (t = 0..3) of output r reads win[4 - t + r]
        T win[TD_R + 4], hv[4];
        ld4(sx + p - 4, win); ld4(sx + p, win + 4); ld4(sx + p + 4, win + 8);
        for (uint32_t k = 0; k < taps; k += 4)
        {
            ld4(sh + k, hv);
#pragma unroll
            for (int r = 0; r < TD_R; r++)
            {
                acc[r] = fma(hv[0], win[4 + r], acc[r]);
                acc[r] = fma(hv[1], win[3 + r], acc[r]);
                acc[r] = fma(hv[2], win[2 + r], acc[r]);
                acc[r] = fma(hv[3], win[1 + r], acc[r]);
            }
            // slide the window four samples towards the past
#pragma unroll
            for (int q = TD_R + 3; q >= 4; q--) win[q] = win[q - 4];
            if (k + 4 < taps) ld4(sx + p - k - 8, win);
        }
    }
    const size_t s = s0 + size_t(TD_R) * threadIdx.x;
#pragma unroll
    for (int r = 0; r < TD_R; r++)
        if (s + r < n)
        {
            T *o = out + size_t(row) * out_ld + s + r;
            *o = add ? *o + acc[r] : acc[r];
        }
}

// new history = last `taps` samples of (old history ++ x[0..n))
template <class T>
__global__ void k_td_hist(const T *__restrict__ old_hist, const T *__restrict__ x, size_t x_ld, T *__restrict__ new_hist, uint32_t taps, size_t n)
{
    const size_t row = blockIdx.y;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < taps; j += gridDim.x * blockDim.x)
    {
        const size_t idx = size_t(j) + n;           // position in the concatenation
        new_hist[row * taps + j] = idx < taps ? old_hist[row * taps + idx] : x[row * x_ld + (idx - taps)];
    }
}

} // namespace

struct hb_matrix
{
    int dtype = HB_F32, device = 0;
    uint32_t groups = 1, ins = 1, outs = 1;
    bool zero_latency = false;
    std::vector<uint32_t> sizes;
    std::vector<hb_conv *> parts;            // fixed parts in scheme order, the resizable tail last
    uintptr_t tail_fft = 0, tail_offset = 0;
    std::vector<uintptr_t> pair_size, pair_len;     // per pair [group][out][in]: allocation and IR length (MonoConvolve mLength)
    // zero-latency head
    uint32_t head_taps = 0;
    void *d_head = nullptr;                  // [pairs][head_taps]
    void *d_hist[2] = {nullptr, nullptr};    // [groups*ins][head_taps], ping-pong
    int hist_cur = 0;
    bool head_reset = true;
    std::vector<uint8_t> head_loaded;        // per pair
    size_t head_count = 0;
    // host-call staging
    cudaStream_t stream = nullptr;
    DevBuf d_in, d_out, d_ir;
    PinnedBuf h_in, h_out, h_ir;
    std::mutex lock;

    size_t esize() const { return dtype_size(dtype); }
    size_t pairs() const { return size_t(groups) * ins * outs; }
    size_t pair_index(uint32_t g, uint32_t i, uint32_t o) const { return (size_t(g) * outs + o) * ins + i; }
};

namespace
{
int check(hb_matrix *m)
{
    if (!m) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    return use_device(m->device);
}

void destroy(hb_matrix *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (hb_conv *p : m->parts) hb_conv_destroy(p);
    cudaFree(m->d_head); cudaFree(m->d_hist[0]); cudaFree(m->d_hist[1]);
    m->d_in.release(); m->d_out.release(); m->d_ir.release();
    m->h_in.release(); m->h_out.release(); m->h_ir.release();
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

// staggered reset phases of MonoConvolve::setResetOffset (MonoConvolve.cpp:85-98); a negative
// (random) request selects phase 0 so that runs are reproducible
// The reference staggers its fixed parts by size / 8 samples (MonoConvolve.cpp:92-96) purely to spread the FFTs of the
// parts over different audio callbacks of ONE CPU thread.  On the GPU a phase other than 0 only forces hop-unaligned calls
// through the staging copies and rules out the multi-hop batches, and the output does not depend on the phase beyond
// rounding (SURVEY B): every part takes the caller's offset as it is.
void apply_reset_offset(hb_matrix *m, intptr_t offset)
{
    if (offset < 0) offset = 0;
    for (hb_conv *p : m->parts) hb_conv_set_reset_offset(p, offset);
}

int grow_tail(hb_matrix *m, uintptr_t size)
{
    const uintptr_t need = std::max<uintptr_t>(size, m->tail_fft) - m->tail_offset;
    if (need > hb_conv_max_length(m->parts.back())) return hb_conv_resize(m->parts.back(), need);
    return 0;
}

template <class T>
int head_store(hb_matrix *m, size_t pair, const void *ir, int ir_dtype, uintptr_t length)
{
    // TimeDomainConvolve::set with offset 0 and length head_taps (TimeDomainConvolve.cpp:69-87)
    const uint32_t taps = m->head_taps;
    int rc;
    if ((rc = m->h_ir.ensure(taps * sizeof(T)))) return rc;
    HB_CUDA(cudaStreamSynchronize(m->stream));
    T *dst = (T *) m->h_ir.p;
    const size_t have = ir ? std::min<size_t>(length, taps) : 0;
    for (size_t k = 0; k < have; k++) dst[k] = ir_dtype == HB_F64 ? (T) ((const double *) ir)[k] : (T) ((const float *) ir)[k];
    for (size_t k = have; k < taps; k++) dst[k] = T(0);
    HB_CUDA(cudaMemcpyAsync((T *) m->d_head + pair * taps, dst, taps * sizeof(T), cudaMemcpyHostToDevice, m->stream));
    HB_CUDA(cudaStreamSynchronize(m->stream));
    const uint8_t now = have ? 1 : 0;
    m->head_count += now;
    m->head_count -= m->head_loaded[pair];
    m->head_loaded[pair] = now;
    m->head_reset = true;
    return HB_OK;
}

int set_parts(hb_matrix *m, uint32_t g, uint32_t i, uint32_t o, const void *ir, int ir_dtype, uintptr_t length)
{
    int rc;
    if (m->head_taps)
    {
        rc = m->dtype == HB_F64 ? head_store<double>(m, m->pair_index(g, i, o), ir, ir_dtype, length)
                                : head_store<float>(m, m->pair_index(g, i, o), ir, ir_dtype, length);
        if (rc) return rc;
    }
    for (hb_conv *p : m->parts)
    {
        rc = hb_conv_set_ir(p, g, i, o, ir, ir_dtype, ir ? length : 0);
        if (rc < 0) return rc;
    }
    return HB_OK;
}

// The sum of all parts into device rows: MonoConvolve::process (MonoConvolve.cpp:179-201) for every pair.
// The reference chains its five slots (Time1, Part1, Part2, Part3, part4) with the flag
// "accumulate || the slot just before this one exists" (:195-199): a slot whose predecessor is absent
// OVERWRITES the output.  With the shipped latency modes every gap-free chain sums all parts; with a
// zero-latency custom scheme of fewer than four sizes the first FFT part replaces the head's output.
// That behaviour is kept as it is (results must equal the reference's).  HB_ERR_NO_IR when nothing is loaded.
template <class T>
int process_rows(hb_matrix *m, const T *d_in, size_t in_ld, T *d_out, size_t out_ld, size_t n, int accumulate, cudaStream_t st)
{
    bool any = false;
    // slot k of the reference: 0 = Time1, 1..3 = Part1..Part3 (the fixed parts fill the LAST of these), 4 = part4
    const size_t fixed = m->parts.size() - 1;
    bool prev_exists = false;
    for (int slot = 0; slot < 5; slot++)
    {
        const bool exists = slot == 0 ? m->head_taps != 0 : (slot == 4 || size_t(slot) > 3 - fixed);
        const bool add = accumulate != 0 || prev_exists;
        prev_exists = exists;
        if (!exists) continue;
        if (slot == 0)
        {
            if (!m->head_count) continue;
            const uint32_t taps = m->head_taps;
            const size_t rows_in = size_t(m->groups) * m->ins;
            if (m->head_reset)
            {
                HB_CUDA(cudaMemsetAsync(m->d_hist[m->hist_cur], 0, rows_in * taps * sizeof(T), st));
                m->head_reset = false;
            }
            if (n >= 4 * TD_BLOCK && taps % 4 == 0)
            {
                // long calls: register-blocked kernel (TD_R samples per thread)
                const size_t span = size_t(TDB) * TD_R;
                dim3 grid((unsigned) ((n + span - 1) / span), m->groups * m->outs);
                const size_t smem = (size_t(2) * taps + span) * sizeof(T);
                k_td_blocked<T><<<grid, TDB, smem, st>>>((const T *) m->d_head, d_in, in_ld, (const T *) m->d_hist[m->hist_cur], d_out, out_ld,
                                                             m->ins, m->outs, taps, n, add ? 1 : 0);
            }
            else
            {
                dim3 grid((unsigned) ((n + TD_BLOCK - 1) / TD_BLOCK), m->groups * m->outs);
                const size_t smem = (size_t(2) * taps + TD_BLOCK) * sizeof(T);
                k_td<T><<<grid, TD_BLOCK, smem, st>>>((const T *) m->d_head, d_in, in_ld, (const T *) m->d_hist[m->hist_cur], d_out, out_ld,
                                                     m->ins, m->outs, taps, n, add ? 1 : 0);
            }
            HB_LAUNCH_CHECK();
            dim3 hgrid((taps + 255) / 256, (unsigned) rows_in);
            k_td_hist<T><<<hgrid, 256, 0, st>>>((const T *) m->d_hist[m->hist_cur], d_in, in_ld, (T *) m->d_hist[m->hist_cur ^ 1], taps, n);
            HB_LAUNCH_CHECK();
            m->hist_cur ^= 1;
            any = true;
            continue;
        }
        hb_conv *p = slot == 4 ? m->parts.back() : m->parts[size_t(slot) - (4 - fixed)];
        const int rc = hb_conv_process_dev(p, d_in, in_ld, d_out, out_ld, n, add ? 1 : 0, st);
        if (rc == HB_OK) any = true;
        else if (rc != HB_ERR_NO_IR) return rc;
    }
    return any ? HB_OK : HB_ERR_NO_IR;
}

bool valid_sizes(const uint32_t in[4], std::vector<uint32_t> &sizes)
{
    // MonoConvolve.cpp:207-229: powers of two are not required here (PartitionedConvolve rounds), range 2^5..2^20, strictly increasing
    uint32_t prev = 0;
    for (int k = 0; k < 4; k++)
    {
        if (in[k] >= (1u << 5) && in[k] <= (1u << 20) && in[k] > prev) { sizes.push_back(in[k]); prev = in[k]; }
        else if (in[k]) return false;
    }
    return !sizes.empty();
}
} // namespace

extern "C" int hb_matrix_create(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                                int zero_latency, uint32_t A, uint32_t B, uint32_t C, uint32_t D, int device)
{
    if (!out || (dtype != HB_F32 && dtype != HB_F64) || !groups || !ins || !outs) { set_error("hb_matrix_create: bad argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    const uint32_t in[4] = {A, B, C, D};
    std::vector<uint32_t> sizes;
    if (!valid_sizes(in, sizes)) { set_error("invalid FFT size or order"); return HB_ERR_BAD_ARG; }     // the reference throws (MonoConvolve.cpp:212,229)
    int rc = use_device(device);
    if (rc) return rc;

    hb_matrix *m = new hb_matrix;
    m->dtype = dtype; m->device = device; m->groups = groups; m->ins = ins; m->outs = outs;
    m->zero_latency = zero_latency != 0;
    m->sizes = sizes;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete m; return HB_ERR_CUDA; }

    // part map of setPartitions (MonoConvolve.cpp:231-252)
    const size_t n = sizes.size();
    uintptr_t offset = m->zero_latency ? sizes[0] >> 1 : 0;
    m->head_taps = (uint32_t) std::min<uintptr_t>(offset, TD_MAX_TAPS);
    auto add_fixed = [&](uint32_t size, uint32_t next) -> int
    {
        const uintptr_t taps = (next - size) >> 1;
        hb_conv *p = nullptr;
        int r = hb_conv_create(&p, dtype, groups, ins, outs, size, taps, offset, taps, device);
        if (r < 0) return r;
        m->parts.push_back(p);
        offset += taps;
        return HB_OK;
    };
    rc = HB_OK;
    if (n == 4) rc = add_fixed(sizes[0], sizes[1]);
    if (rc >= 0 && n > 2) rc = add_fixed(sizes[n - 3], sizes[n - 2]);
    if (rc >= 0 && n > 1) rc = add_fixed(sizes[n - 2], sizes[n - 1]);
    if (rc >= 0)
    {
        // tail allocator of MonoConvolve.cpp:247-250: PartitionedConvolve(largest, max(size, largest) - offset, offset, 0)
        m->tail_fft = sizes[n - 1];
        m->tail_offset = offset;
        hb_conv *p = nullptr;
        rc = hb_conv_create(&p, dtype, groups, ins, outs, m->tail_fft, std::max<uintptr_t>(max_length, m->tail_fft) - offset, offset, 0, device);
        if (rc >= 0) m->parts.push_back(p);
    }
    if (rc < 0) { destroy(m); return rc; }
    m->pair_size.assign(m->pairs(), max_length);
    m->pair_len.assign(m->pairs(), 0);
    if (m->head_taps)
    {
        const size_t hb_ = m->pairs() * m->head_taps * m->esize(), xb = size_t(groups) * ins * m->head_taps * m->esize();
        if (cudaMalloc(&m->d_head, hb_) != cudaSuccess || cudaMalloc(&m->d_hist[0], xb) != cudaSuccess || cudaMalloc(&m->d_hist[1], xb) != cudaSuccess)
        {
            set_error("device allocation failed for the zero-latency head");
            destroy(m);
            return HB_ERR_CUDA;
        }
        cudaMemset(m->d_head, 0, hb_);
        m->head_loaded.assign(m->pairs(), 0);
    }
    apply_reset_offset(m, 0);
    *out = m;
    return HB_OK;
}

extern "C" int hb_matrix_create_latency(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                                        int latency_mode, int device)
{
    // MonoConvolve.cpp:26-32
    switch (latency_mode)
    {
        case 0: return hb_matrix_create(out, dtype, groups, ins, outs, max_length, 1, 256, 1024, 4096, 16384, device);
        case 1: return hb_matrix_create(out, dtype, groups, ins, outs, max_length, 0, 256, 1024, 4096, 16384, device);
        case 2: return hb_matrix_create(out, dtype, groups, ins, outs, max_length, 0, 1024, 4096, 16384, 0, device);
    }
    set_error("unknown LatencyMode %d", latency_mode);
    return HB_ERR_BAD_ARG;
}

extern "C" void hb_matrix_destroy(hb_matrix *m) { destroy(m); }

extern "C" int hb_matrix_set_reset_offset(hb_matrix *m, intptr_t offset)
{
    if (!m) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);
    apply_reset_offset(m, offset);
    return HB_OK;
}

extern "C" int hb_matrix_resize(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out, uintptr_t length)
{
    int rc = check(m);
    if (rc) return rc;
    if (group >= m->groups || in >= m->ins || out >= m->outs) { set_error("hb_matrix_resize: pair out of range"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);                  // blocking, as MemorySwap::equal (MonoConvolve.cpp:100-110)
    const size_t pair = m->pair_index(group, in, out);
    m->pair_len[pair] = 0;
    const int grown = grow_tail(m, length);
    rc = set_parts(m, group, in, out, nullptr, HB_F32, 0);
    if (rc < 0) return rc;
    if (grown) return 3;                                     // CONVOLVE_ERR_MEM_UNAVAILABLE
    m->pair_size[pair] = length;
    return 0;
}

extern "C" int hb_matrix_set(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out, const void *ir, int ir_dtype, uintptr_t length, int request_resize)
{
    int rc = check(m);
    if (rc) return rc;
    if (group >= m->groups || in >= m->ins || out >= m->outs || (ir_dtype != HB_F32 && ir_dtype != HB_F64)) { set_error("hb_matrix_set: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);                  // MonoConvolve::set (MonoConvolve.cpp:118-140)
    if (!ir) length = 0;
    const size_t pair = m->pair_index(group, in, out);
    m->pair_len[pair] = 0;
    if (request_resize && length != m->pair_size[pair])
    {
        if (grow_tail(m, length))
        {
            rc = set_parts(m, group, in, out, nullptr, HB_F32, 0);
            if (rc < 0) return rc;
            return length ? 3 : 0;
        }
        m->pair_size[pair] = length;
    }
    const uintptr_t size = m->pair_size[pair];
    // process() ignores a pair whose IR is longer than its allocation (MonoConvolve.cpp:183)
    const bool active = length && length <= size;
    rc = set_parts(m, group, in, out, active ? ir : nullptr, ir_dtype, active ? length : 0);
    if (rc < 0) return rc;
    m->pair_len[pair] = length;
    return length > size ? 4 : 0;                            // CONVOLVE_ERR_MEM_ALLOC_TOO_SMALL
}

extern "C" int hb_matrix_reset(hb_matrix *m)
{
    if (!m) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);
    for (hb_conv *p : m->parts) hb_conv_reset(p);
    m->head_reset = true;
    return 0;
}

extern "C" uint32_t hb_matrix_parts(const hb_matrix *m) { return m ? (uint32_t) m->parts.size() : 0; }
extern "C" hb_conv *hb_matrix_part(hb_matrix *m, uint32_t index) { return (m && index < m->parts.size()) ? m->parts[index] : nullptr; }
extern "C" uint32_t hb_matrix_head_taps(const hb_matrix *m) { return m ? m->head_taps : 0; }

extern "C" int hb_matrix_process_dev(hb_matrix *m, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld, uintptr_t n,
                                     int accumulate, void *stream)
{
    int rc = check(m);
    if (rc) return rc;
    if ((!d_in || !d_out) && n) { set_error("hb_matrix_process_dev: null buffer"); return HB_ERR_BAD_ARG; }
    // the audio thread never waits for set / resize: the block is skipped (MonoConvolve.cpp:181-183, MemorySwap.h:182-185)
    std::unique_lock<std::mutex> g(m->lock, std::try_to_lock);
    if (!g.owns_lock()) return HB_ERR_BUSY;
    if (!n) return HB_OK;
    cudaStream_t st = stream ? (cudaStream_t) stream : m->stream;
    return m->dtype == HB_F64 ? process_rows<double>(m, (const double *) d_in, in_ld, (double *) d_out, out_ld, n, accumulate, st)
                              : process_rows<float>(m, (const float *) d_in, in_ld, (float *) d_out, out_ld, n, accumulate, st);
}

extern "C" int hb_matrix_process(hb_matrix *m, const void *const *ins, void *const *outs, uintptr_t n, int accumulate)
{
    int rc = check(m);
    if (rc) return rc;
    if ((!ins || !outs) && n) { set_error("hb_matrix_process: null buffer"); return HB_ERR_BAD_ARG; }
    std::unique_lock<std::mutex> g(m->lock, std::try_to_lock);
    if (!g.owns_lock()) return HB_ERR_BUSY;
    if (!n) return HB_OK;
    // a single uniform part (no head, one FFT size) is exactly one engine: its host path defers the device
    // work of hop-aligned calls behind the API's own one-hop latency (hb_conv_process)
    if (!m->head_taps && m->parts.size() == 1) return hb_conv_process(m->parts[0], ins, outs, n, accumulate);
    bool loaded = m->head_count != 0;
    for (hb_conv *p : m->parts) loaded = loaded || hb_conv_partitions(p) != 0;
    if (!loaded) return HB_ERR_NO_IR;
    const size_t es = m->esize();
    const size_t rows_in = size_t(m->groups) * m->ins, rows_out = size_t(m->groups) * m->outs;
    if ((rc = m->h_in.ensure(rows_in * n * es)) || (rc = m->h_out.ensure(rows_out * n * es)) ||
        (rc = m->d_in.ensure(rows_in * n * es)) || (rc = m->d_out.ensure(rows_out * n * es))) return rc;
    for (size_t r = 0; r < rows_in; r++)
    {
        // a null input row is an inactive channel: silence (NToMonoConvolve.cpp:41 stops at activeInChans)
        if (ins[r]) memcpy((char *) m->h_in.p + r * n * es, ins[r], n * es);
        else memset((char *) m->h_in.p + r * n * es, 0, n * es);
    }
    HB_CUDA(cudaMemcpyAsync(m->d_in.p, m->h_in.p, rows_in * n * es, cudaMemcpyHostToDevice, m->stream));
    // parts sum into a zeroed device block; `accumulate` keeps its reference meaning for the slot chain
    HB_CUDA(cudaMemsetAsync(m->d_out.p, 0, rows_out * n * es, m->stream));
    rc = m->dtype == HB_F64 ? process_rows<double>(m, (const double *) m->d_in.p, n, (double *) m->d_out.p, n, n, accumulate, m->stream)
                            : process_rows<float>(m, (const float *) m->d_in.p, n, (float *) m->d_out.p, n, n, accumulate, m->stream);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(m->h_out.p, m->d_out.p, rows_out * n * es, cudaMemcpyDeviceToHost, m->stream));
    HB_CUDA(cudaStreamSynchronize(m->stream));
    for (size_t r = 0; r < rows_out; r++)
    {
        if (!outs[r]) continue;
        if (!accumulate) memcpy(outs[r], (char *) m->h_out.p + r * n * es, n * es);
        else if (m->dtype == HB_F64)
        {
            double *d = (double *) outs[r];
            const double *s = (const double *) m->h_out.p + r * n;
            for (size_t k = 0; k < n; k++) d[k] += s[k];                   // MonoConvolve.cpp:167-177
        }
        else
        {
            float *d = (float *) outs[r];
            const float *s = (const float *) m->h_out.p + r * n;
            for (size_t k = 0; k < n; k++) d[k] += s[k];
        }
    }
    return HB_OK;
}
