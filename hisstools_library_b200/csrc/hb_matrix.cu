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
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
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
// age (optional): samples each pair has seen since it was restarted alone (hb_matrix_set on a running matrix, hb_matrix_reset_pair),
// saturating at `taps`: a pair only reaches that far back into the input, as a freshly reset TimeDomainConvolve would.
template <class T>
__global__ void __launch_bounds__(TD_BLOCK) k_td(const T *__restrict__ h, const T *__restrict__ x, size_t x_ld,
                                                 const T *__restrict__ hist, T *__restrict__ out, size_t out_ld,
                                                 uint32_t ins, uint32_t outs, uint32_t taps, size_t n, int add, const uint32_t *__restrict__ age)
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
        uint32_t reach = taps;                      // taps k < reach: x[s - k] is input from after the pair's restart
        if (age) reach = (uint32_t) min((unsigned long long) taps, (unsigned long long) s + age[size_t(row) * ins + i] + 1ull);
#pragma unroll 4
        for (uint32_t k = 0; k < reach; k++) acc = fma(sh[k], w[-(int) k], acc);
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

struct MultiFront;

struct hb_matrix
{
    MultiFront *multi = nullptr;             // non-null: this handle is the front of a matrix dealt to several GPUs (below)
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
    std::vector<uint32_t> head_age;          // per pair: samples since the pair was restarted alone, saturating at head_taps
    size_t head_young = 0;                   // pairs whose age is below head_taps
    void *d_head_age = nullptr;
    bool head_age_dirty = false;
    bool running = false;                    // a block has been processed since the last whole reset
    int hop_overlap = 1;                     // hb_matrix_set_hop_overlap: 0 never, 1 on the matrix's own stream, 2 on any stream
    int parts_overlap = -1;                  // what the parts were last told (hb_conv_set_hop_overlap: 0 or 2)
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
    if (m->multi) return HB_OK;               // a multi-device front leaves the caller's device alone: its shards select theirs
    return use_device(m->device);
}

void destroy_front(hb_matrix *m);

void destroy(hb_matrix *m)
{
    if (!m) return;
    if (m->multi) { destroy_front(m); return; }
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (hb_conv *p : m->parts) hb_conv_destroy(p);
    cudaFree(m->d_head); cudaFree(m->d_hist[0]); cudaFree(m->d_hist[1]); cudaFree(m->d_head_age);
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

void head_restart_pair(hb_matrix *m, size_t pair)
{
    if (!m->head_taps) return;
    if (m->head_age[pair] >= m->head_taps) m->head_young++;
    m->head_age[pair] = 0;
    m->head_age_dirty = true;
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
    // a matrix that is running restarts this pair alone (MonoConvolve::set resets one object, MonoConvolve.cpp:118-140): the pair
    // forgets the input it has seen, the other pairs keep their history
    if (m->running && !m->head_reset) head_restart_pair(m, pair);
    else m->head_reset = true;
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
        // on a running engine the pair restarts alone where that is possible in place (hb_conv_set_ir_live)
        rc = hb_conv_set_ir_live(p, g, i, o, ir, ir_dtype, ir ? length : 0);
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
                // a whole reset: every pair starts from silence together, no pair is younger than the history
                std::fill(m->head_age.begin(), m->head_age.end(), taps);
                m->head_young = 0;
                m->head_age_dirty = false;
            }
            const uint32_t *d_age = nullptr;
            if (m->head_young)
            {
                if (m->head_age_dirty)
                {
                    HB_CUDA(cudaMemcpyAsync(m->d_head_age, m->head_age.data(), m->head_age.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                    HB_CUDA(cudaStreamSynchronize(st));                  // the host copy changes below
                    m->head_age_dirty = false;
                }
                d_age = (const uint32_t *) m->d_head_age;
            }
            if (!d_age && n >= 4 * TD_BLOCK && taps % 4 == 0)
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
                                                     m->ins, m->outs, taps, n, add ? 1 : 0, d_age);
            }
            if (m->head_young)
            {
                // the restarted pairs have now seen n more samples
                for (uint32_t &a : m->head_age)
                    if (a < taps)
                    {
                        a = (uint32_t) std::min<size_t>(taps, size_t(a) + n);
                        if (a >= taps) m->head_young--;
                    }
                m->head_age_dirty = true;
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
    if (any) m->running = true;
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

// =====================================================================================================
// One matrix on several GPUs of this process (hb_matrix_create_multi).
//
// The front handle owns one ordinary single-device hb_matrix per device ("shard") and a host worker thread per
// device; every hb_matrix_* entry point on the front fans out to them.
//   N x M matrix (groups == 1): the INPUT channels are dealt to the devices -- shard d holds inputs
//     [d * ins / n, (d + 1) * ins / n) against all outputs (1 / n of the spectra and of the HBM traffic) and owns the
//     sums of outputs [d * outs / n, ...).  The sum over inputs of NToMonoConvolve.cpp:39-42 then crosses devices:
//       fused   (one uniform FFT size, no head): the engines of the shards form a fused exchange (hb_conv_shard_*):
//               every inverse-FFT kernel stores its partial blocks straight into the owner's memory over NVLink and
//               the host call is hb_conv_process on every engine, pipelined as on one device;
//       generic (any partition scheme, any call size): every shard adds its parts into a local block of partial
//               outputs, and each owner sums its rows out of all shards' blocks with one kernel that reads peer memory.
//   parallel banks (groups > 1): the banks are dealt to the devices; nothing crosses.
// =====================================================================================================
namespace
{
// calls on a multi-device front visit several devices: the caller's current device is put back on return
struct DeviceRestore
{
    int dev = -1;
    DeviceRestore() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
    ~DeviceRestore() { if (dev >= 0) cudaSetDevice(dev); }
};

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

// one worker thread per device: run(fn) calls fn(d) on every worker and returns the first failure
class DevicePool
{
public:
    void start(const std::vector<int> &devices)
    {
        n_ = (uint32_t) devices.size();
        rc_.assign(n_, 0);
        err_.assign(n_, std::string());
        for (uint32_t d = 0; d < n_; d++)
            threads_.emplace_back([this, d, dev = devices[d]]() { cudaSetDevice(dev); loop(d); });
    }
    void stop()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_.store(true, std::memory_order_release);
        }
        cv_.notify_all();
        for (std::thread &t : threads_) t.join();
        threads_.clear();
    }
    int run(const std::function<int(uint32_t)> &fn)
    {
        job_ = &fn;
        pending_.store(n_, std::memory_order_release);
        {
            std::lock_guard<std::mutex> l(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (uint32_t spins = 0; pending_.load(std::memory_order_acquire) != 0; spins++)
            if (spins < 200000) cpu_relax(); else std::this_thread::yield();
        for (uint32_t d = 0; d < n_; d++)
            if (rc_[d] < 0 && rc_[d] != HB_ERR_NO_IR && rc_[d] != HB_ERR_BUSY) { set_error("device %u: %s", d, err_[d].c_str()); return rc_[d]; }
        return HB_OK;
    }
    int code(uint32_t d) const { return rc_[d]; }

private:
    void loop(uint32_t d)
    {
        uint64_t seen = 0;
        for (;;)
        {
            // a streaming caller comes back within microseconds: spin first, sleep only when it stays away
            uint32_t spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen && !quit_.load(std::memory_order_acquire))
            {
                if (++spins < 100000) { cpu_relax(); continue; }
                std::unique_lock<std::mutex> l(m_);
                cv_.wait_for(l, std::chrono::milliseconds(50), [&]() { return gen_.load(std::memory_order_acquire) != seen || quit_.load(std::memory_order_acquire); });
            }
            if (quit_.load(std::memory_order_acquire)) return;
            seen = gen_.load(std::memory_order_acquire);
            const int r = (*job_)(d);
            rc_[d] = r;
            if (r < 0) err_[d] = hb_last_error();
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    uint32_t n_ = 0;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<uint32_t> pending_{0};
    std::atomic<bool> quit_{false};
    const std::function<int(uint32_t)> *job_ = nullptr;
    std::vector<int> rc_;
    std::vector<std::string> err_;
};

constexpr int MULTI_MAX = 16;
struct PeerParts { const void *p[MULTI_MAX]; };

// owner-side sum of the generic path: out[r][s] = sum over shards of part_e[row0 + r][s]; the blocks of the other shards are
// read straight out of their memory (peer access over NVLink)
template <class T>
__global__ void k_peer_sum(PeerParts parts, uint32_t world, size_t ld, uint32_t row0, size_t n, T *__restrict__ out, size_t out_ld)
{
    const size_t r = blockIdx.y;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    {
        T sum = T(0);
        for (uint32_t e = 0; e < world; e++) sum += reinterpret_cast<const T *>(parts.p[e])[(row0 + r) * ld + i];
        out[r * out_ld + i] = sum;
    }
}
} // namespace

struct MultiShard
{
    hb_matrix *m = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_part = nullptr;
    DevBuf d_in, d_part, d_sum;
    PinnedBuf h_in, h_out;
    int loaded = 0;                          // result of the last local process: did this shard have anything loaded
};

struct MultiFront
{
    std::vector<MultiShard> sh;
    bool by_inputs = true;                   // N x M matrix dealt by input channel; false: parallel banks dealt by bank
    bool fused = false;                      // the shards' engines form a fused exchange
    uint32_t l_ins = 0, l_outs = 0, l_groups = 0;
    DevicePool pool;
};

namespace
{
void destroy_front(hb_matrix *m)
{
    DeviceRestore keep;
    MultiFront *f = m->multi;
    f->pool.stop();
    // the engines write into each other's memory: every device must be idle before any shard goes
    for (MultiShard &s : f->sh)
        if (cudaSetDevice(s.device) == cudaSuccess) cudaDeviceSynchronize();
    for (MultiShard &s : f->sh)
    {
        cudaSetDevice(s.device);
        s.d_in.release(); s.d_part.release(); s.d_sum.release(); s.h_in.release(); s.h_out.release();
        if (s.ev_part) cudaEventDestroy(s.ev_part);
        if (s.stream) cudaStreamDestroy(s.stream);
        destroy(s.m);
    }
    delete f;
    delete m;
}

// which shard holds pair (group, in, out) and under which local indices
uint32_t front_locate(const hb_matrix *m, uint32_t &group, uint32_t &in)
{
    const MultiFront *f = m->multi;
    if (f->by_inputs) { const uint32_t d = in / f->l_ins; in -= d * f->l_ins; return d; }
    const uint32_t d = group / f->l_groups;
    group -= d * f->l_groups;
    return d;
}

int front_process(hb_matrix *m, const void *const *ins, void *const *outs, size_t n, int accumulate)
{
    MultiFront *f = m->multi;
    const uint32_t nd = (uint32_t) f->sh.size();
    const size_t es = m->esize();
    if (f->fused)
    {
        // every engine takes its inputs and returns the outputs it owns: hb_conv_process pipelines the call as on one device
        bool any = false;
        for (MultiShard &s : f->sh) any = any || hb_conv_partitions(s.m->parts[0]) != 0;
        if (!any) return HB_ERR_NO_IR;
        return f->pool.run([&](uint32_t d) -> int
        {
            return hb_conv_process(f->sh[d].m->parts[0], ins + size_t(d) * f->l_ins, outs + size_t(d) * f->l_outs, n, accumulate);
        });
    }
    // generic path, phase A: every shard adds its parts into a zeroed local block (all outputs of its banks)
    const size_t rows_in = size_t(f->l_groups) * f->l_ins;                       // per shard
    const size_t rows_part = size_t(f->l_groups) * m->outs;
    int rc = f->pool.run([&](uint32_t d) -> int
    {
        MultiShard &s = f->sh[d];
        int r;
        if ((r = use_device(s.device))) return r;
        if ((r = s.h_in.ensure(rows_in * n * es)) || (r = s.d_in.ensure(rows_in * n * es)) || (r = s.d_part.ensure(rows_part * n * es))) return r;
        const void *const *my = ins + size_t(d) * rows_in;                      // inputs (or banks) are dealt in contiguous runs
        for (size_t q = 0; q < rows_in; q++)
        {
            if (my[q]) memcpy((char *) s.h_in.p + q * n * es, my[q], n * es);
            else memset((char *) s.h_in.p + q * n * es, 0, n * es);
        }
        HB_CUDA(cudaMemcpyAsync(s.d_in.p, s.h_in.p, rows_in * n * es, cudaMemcpyHostToDevice, s.stream));
        HB_CUDA(cudaMemsetAsync(s.d_part.p, 0, rows_part * n * es, s.stream));
        r = hb_matrix_process_dev(s.m, s.d_in.p, n, s.d_part.p, n, n, accumulate, s.stream);
        s.loaded = r == HB_OK;
        if (r != HB_OK && r != HB_ERR_NO_IR) return r;
        HB_CUDA(cudaEventRecord(s.ev_part, s.stream));
        return HB_OK;
    });
    if (rc) return rc;
    bool any = false;
    for (MultiShard &s : f->sh) any = any || s.loaded;
    // phase B: every owner sums its rows out of all shards' blocks (or, for banks, just takes its own) and hands them over
    const size_t rows_own = f->by_inputs ? f->l_outs : rows_part;
    rc = f->pool.run([&](uint32_t d) -> int
    {
        MultiShard &s = f->sh[d];
        int r;
        if ((r = use_device(s.device))) return r;
        const void *src = s.d_part.p;
        if (f->by_inputs)
        {
            if ((r = s.d_sum.ensure(rows_own * n * es))) return r;
            PeerParts parts;
            for (uint32_t e = 0; e < nd; e++)
            {
                parts.p[e] = f->sh[e].d_part.p;
                if (e != d) HB_CUDA(cudaStreamWaitEvent(s.stream, f->sh[e].ev_part, 0));
            }
            dim3 grid((unsigned) std::min<size_t>((n + 255) / 256, 64), (unsigned) rows_own);
            if (m->dtype == HB_F64) k_peer_sum<double><<<grid, 256, 0, s.stream>>>(parts, nd, n, d * f->l_outs, n, (double *) s.d_sum.p, n);
            else k_peer_sum<float><<<grid, 256, 0, s.stream>>>(parts, nd, n, d * f->l_outs, n, (float *) s.d_sum.p, n);
            HB_LAUNCH_CHECK();
            src = s.d_sum.p;
        }
        if (any)
        {
            if ((r = s.h_out.ensure(rows_own * n * es))) return r;
            HB_CUDA(cudaMemcpyAsync(s.h_out.p, src, rows_own * n * es, cudaMemcpyDeviceToHost, s.stream));
        }
        HB_CUDA(cudaStreamSynchronize(s.stream));                               // also: the peers may overwrite their blocks again
        if (!any) return HB_OK;
        void *const *mine = outs + size_t(d) * rows_own;
        for (size_t q = 0; q < rows_own; q++)
        {
            if (!mine[q]) continue;
            const char *sp = (const char *) s.h_out.p + q * n * es;
            if (!accumulate) memcpy(mine[q], sp, n * es);
            else if (m->dtype == HB_F64) { double *dd = (double *) mine[q]; const double *ss = (const double *) sp; for (size_t k = 0; k < n; k++) dd[k] += ss[k]; }
            else { float *dd = (float *) mine[q]; const float *ss = (const float *) sp; for (size_t k = 0; k < n; k++) dd[k] += ss[k]; }
        }
        return HB_OK;
    });
    if (rc) return rc;
    return any ? HB_OK : HB_ERR_NO_IR;
}

int create_front(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                 const std::function<int(hb_matrix **, uint32_t, uint32_t, int)> &make_shard, const int *devices, uint32_t n_devices)
{
    if (!out || !devices || n_devices < 1 || n_devices > (uint32_t) MULTI_MAX) { set_error("hb_matrix_create_multi: 1 to %d devices", MULTI_MAX); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    DeviceRestore keep;
    for (uint32_t a = 0; a < n_devices; a++)
        for (uint32_t b = a + 1; b < n_devices; b++)
            if (devices[a] == devices[b]) { set_error("hb_matrix_create_multi: device %d listed twice", devices[a]); return HB_ERR_BAD_ARG; }
    const bool by_inputs = groups == 1;
    if (by_inputs ? (ins % n_devices || outs % n_devices) : (groups % n_devices) != 0)
    {
        set_error("hb_matrix_create_multi: %s must be a multiple of the device count %u", by_inputs ? "inputs and outputs" : "the bank count", n_devices);
        return HB_ERR_BAD_ARG;
    }
    hb_matrix *m = new hb_matrix;
    MultiFront *f = new MultiFront;
    m->multi = f;
    m->dtype = dtype; m->device = devices[0]; m->groups = groups; m->ins = ins; m->outs = outs;
    f->by_inputs = by_inputs;
    f->l_ins = by_inputs ? ins / n_devices : ins;
    f->l_outs = by_inputs ? outs / n_devices : outs;
    f->l_groups = by_inputs ? 1 : groups / n_devices;
    f->sh.resize(n_devices);
    int rc = HB_OK;
    for (uint32_t d = 0; d < n_devices && rc >= 0; d++)
    {
        MultiShard &s = f->sh[d];
        s.device = devices[d];
        if ((rc = use_device(s.device))) break;
        rc = make_shard(&s.m, f->l_groups, f->l_ins, s.device);
        if (rc < 0) break;
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&s.ev_part, cudaEventDisableTiming) != cudaSuccess)
        {
            set_error("stream / event creation failed on device %d", s.device);
            rc = HB_ERR_CUDA;
        }
    }
    // every owner reads (generic) or writes (fused) the other shards' memory
    for (uint32_t a = 0; a < n_devices && rc >= 0 && by_inputs; a++)
        for (uint32_t b = 0; b < n_devices && rc >= 0; b++)
        {
            if (a == b) continue;
            int can = 0;
            cudaSetDevice(devices[a]);
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can)
            {
                set_error("device %d cannot map the memory of device %d (peer access over NVLink / PCIe is required)", devices[a], devices[b]);
                rc = HB_ERR_UNSUPPORTED;
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d) -> %s", devices[b], cudaGetErrorString(e)); cudaGetLastError(); rc = HB_ERR_CUDA; }
        }
    // one uniform part and no head: the engines exchange inside their inverse-FFT kernels
    if (rc >= 0 && by_inputs && n_devices > 1 && f->sh[0].m->parts.size() == 1 && !f->sh[0].m->head_taps)
    {
        static const char *env = getenv("HB_MULTI_GENERIC");                  // experiments only: force the generic path
        if (!(env && atoi(env)))
        {
            std::vector<hb_conv *> engines;
            for (MultiShard &s : f->sh) engines.push_back(s.m->parts[0]);
            for (uint32_t d = 0; d < n_devices && rc >= 0; d++) rc = hb_conv_shard_export(engines[d], n_devices, d, nullptr);
            for (uint32_t d = 0; d < n_devices && rc >= 0; d++) rc = hb_conv_shard_attach_local(engines[d], engines.data());
            f->fused = rc >= 0;
        }
    }
    if (rc < 0)
    {
        const std::string keep = hb_last_error();
        f->sh.erase(std::remove_if(f->sh.begin(), f->sh.end(), [](const MultiShard &s) { return s.m == nullptr && s.stream == nullptr; }), f->sh.end());
        destroy_front(m);
        set_error("%s", keep.c_str());
        return rc;
    }
    (void) max_length;
    std::vector<int> devs(devices, devices + n_devices);
    f->pool.start(devs);
    *out = m;
    return HB_OK;
}
} // namespace

extern "C" int hb_matrix_create_multi(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                                      int zero_latency, uint32_t A, uint32_t B, uint32_t C, uint32_t D, const int *devices, uint32_t n_devices)
{
    if ((dtype != HB_F32 && dtype != HB_F64) || !groups || !ins || !outs) { set_error("hb_matrix_create_multi: bad argument"); return HB_ERR_BAD_ARG; }
    return create_front(out, dtype, groups, ins, outs, max_length, [&](hb_matrix **shard, uint32_t l_groups, uint32_t l_ins, int device)
    {
        return hb_matrix_create(shard, dtype, l_groups, l_ins, outs, max_length, zero_latency, A, B, C, D, device);
    }, devices, n_devices);
}

extern "C" int hb_matrix_create_latency_multi(hb_matrix **out, int dtype, uint32_t groups, uint32_t ins, uint32_t outs, uintptr_t max_length,
                                              int latency_mode, const int *devices, uint32_t n_devices)
{
    if ((dtype != HB_F32 && dtype != HB_F64) || !groups || !ins || !outs) { set_error("hb_matrix_create_latency_multi: bad argument"); return HB_ERR_BAD_ARG; }
    return create_front(out, dtype, groups, ins, outs, max_length, [&](hb_matrix **shard, uint32_t l_groups, uint32_t l_ins, int device)
    {
        return hb_matrix_create_latency(shard, dtype, l_groups, l_ins, outs, max_length, latency_mode, device);
    }, devices, n_devices);
}

extern "C" uint32_t hb_matrix_shards(const hb_matrix *m) { return !m ? 0 : (m->multi ? (uint32_t) m->multi->sh.size() : 1u); }
extern "C" hb_matrix *hb_matrix_shard(hb_matrix *m, uint32_t index)
{
    if (!m) return nullptr;
    if (!m->multi) return index == 0 ? m : nullptr;
    return index < m->multi->sh.size() ? m->multi->sh[index].m : nullptr;
}
extern "C" int hb_matrix_exchange(const hb_matrix *m) { return !m || !m->multi ? 0 : (!m->multi->by_inputs ? 0 : (m->multi->fused ? 2 : 1)); }

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
        m->head_age.assign(m->pairs(), m->head_taps);
        if (cudaMalloc(&m->d_head_age, m->pairs() * sizeof(uint32_t)) != cudaSuccess)
        {
            set_error("device allocation failed for the zero-latency head");
            destroy(m);
            return HB_ERR_CUDA;
        }
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
    if (m->multi)
    {
        DeviceRestore keep;
        for (MultiShard &s : m->multi->sh) hb_matrix_set_reset_offset(s.m, offset);
        return HB_OK;
    }
    apply_reset_offset(m, offset);
    return HB_OK;
}

extern "C" int hb_matrix_resize(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out, uintptr_t length)
{
    int rc = check(m);
    if (rc) return rc;
    if (group >= m->groups || in >= m->ins || out >= m->outs) { set_error("hb_matrix_resize: pair out of range"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);                  // blocking, as MemorySwap::equal (MonoConvolve.cpp:100-110)
    if (m->multi)
    {
        DeviceRestore keep;
        const uint32_t d = front_locate(m, group, in);
        return hb_matrix_resize(m->multi->sh[d].m, group, in, out, length);
    }
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
    if (m->multi)
    {
        DeviceRestore keep;
        const uint32_t d = front_locate(m, group, in);
        return hb_matrix_set(m->multi->sh[d].m, group, in, out, ir, ir_dtype, length, request_resize);
    }
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
    if (m->multi)
    {
        for (MultiShard &s : m->multi->sh) hb_matrix_reset(s.m);
        return 0;
    }
    for (hb_conv *p : m->parts) hb_conv_reset(p);
    m->head_reset = true;
    m->running = false;
    return 0;
}

extern "C" int hb_matrix_reset_pair(hb_matrix *m, uint32_t group, uint32_t in, uint32_t out)
{
    int rc = check(m);
    if (rc) return rc;
    if (group >= m->groups || in >= m->ins || out >= m->outs) { set_error("hb_matrix_reset_pair: pair out of range"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);
    if (m->multi)
    {
        DeviceRestore keep;
        const uint32_t d = front_locate(m, group, in);
        return hb_matrix_reset_pair(m->multi->sh[d].m, group, in, out);
    }
    // MonoConvolve::reset of one object (Convolver.cpp:88-97): every part of the pair and its head restart from silence
    for (hb_conv *p : m->parts)
    {
        rc = hb_conv_reset_pair(p, group, in, out);
        if (rc < 0) return rc;
    }
    if (m->head_taps)
    {
        if (m->running && !m->head_reset) head_restart_pair(m, m->pair_index(group, in, out));
        else m->head_reset = true;
    }
    return 0;
}

// (on a multi-device front: those of the first shard -- every shard has the same scheme)
extern "C" uint32_t hb_matrix_parts(const hb_matrix *m) { return !m ? 0 : (m->multi ? hb_matrix_parts(m->multi->sh[0].m) : (uint32_t) m->parts.size()); }
extern "C" hb_conv *hb_matrix_part(hb_matrix *m, uint32_t index)
{
    if (m && m->multi) return hb_matrix_part(m->multi->sh[0].m, index);
    return (m && index < m->parts.size()) ? m->parts[index] : nullptr;
}
extern "C" uint32_t hb_matrix_head_taps(const hb_matrix *m) { return !m ? 0 : (m->multi ? m->multi->sh[0].m->head_taps : m->head_taps); }

extern "C" int hb_matrix_process_dev(hb_matrix *m, const void *d_in, uintptr_t in_ld, void *d_out, uintptr_t out_ld, uintptr_t n,
                                     int accumulate, void *stream)
{
    int rc = check(m);
    if (rc) return rc;
    if ((!d_in || !d_out) && n) { set_error("hb_matrix_process_dev: null buffer"); return HB_ERR_BAD_ARG; }
    if (m->multi) { set_error("hb_matrix_process_dev: a multi-device matrix takes host pointers (hb_matrix_process); its shards (hb_matrix_shard) take device pointers"); return HB_ERR_UNSUPPORTED; }
    // the audio thread never waits for set / resize: the block is skipped (MonoConvolve.cpp:181-183, MemorySwap.h:182-185)
    std::unique_lock<std::mutex> g(m->lock, std::try_to_lock);
    if (!g.owns_lock()) return HB_ERR_BUSY;
    if (!n) return HB_OK;
    cudaStream_t st = stream ? (cudaStream_t) stream : m->stream;
    // consecutive fused hops of a part may overlap where the rows of a call are complete when it is made: the matrix's own
    // stream (nothing can order them behind the caller's work there), or any stream if the caller says so
    const int overlap = (m->hop_overlap == 2 || (m->hop_overlap == 1 && !stream)) ? 2 : 0;
    if (overlap != m->parts_overlap)
    {
        for (hb_conv *p : m->parts) hb_conv_set_hop_overlap(p, overlap);
        m->parts_overlap = overlap;
    }
    return m->dtype == HB_F64 ? process_rows<double>(m, (const double *) d_in, in_ld, (double *) d_out, out_ld, n, accumulate, st)
                              : process_rows<float>(m, (const float *) d_in, in_ld, (float *) d_out, out_ld, n, accumulate, st);
}

extern "C" int hb_matrix_set_hop_overlap(hb_matrix *m, int mode)
{
    if (!m) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    if (mode < 0 || mode > 2) { set_error("hop overlap mode must be 0, 1 or 2"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(m->lock);
    m->hop_overlap = mode;
    if (m->multi)
        for (MultiShard &s : m->multi->sh) hb_matrix_set_hop_overlap(s.m, mode);
    return HB_OK;
}

extern "C" int hb_matrix_process(hb_matrix *m, const void *const *ins, void *const *outs, uintptr_t n, int accumulate)
{
    int rc = check(m);
    if (rc) return rc;
    if ((!ins || !outs) && n) { set_error("hb_matrix_process: null buffer"); return HB_ERR_BAD_ARG; }
    std::unique_lock<std::mutex> g(m->lock, std::try_to_lock);
    if (!g.owns_lock()) return HB_ERR_BUSY;
    if (!n) return HB_OK;
    if (m->multi) return front_process(m, ins, outs, n, accumulate);
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
