// hb_fft.cu -- the FFT family of the C ABI (include/hisstools_b200.h) and its kernels.
//
// Replaces HISSTools_FFT (reference: HISSTools_FFT/HISSTools_FFT.h:87-369, .cpp:111-248, Core:1293-1374)
// with one CTA per transform: the whole complex array lives in shared memory, a Stockham radix-8/4/2
// FFT runs in place (hb_fft_block.cuh), and the even/odd (un)zip, zero padding and the real split
// pass are folded into the load / store stages, so every transform is one pass over global memory.
#include "hb_common.cuh"
#include "hb_fft_block.cuh"
#include "hb_fft_big.cuh"

#include <cmath>
#include <mutex>
#include <vector>

namespace hb
{

// ---------------------------------------------------------------------------------------------
// process-wide bookkeeping
// ---------------------------------------------------------------------------------------------
static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int use_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
        set_error("no CUDA device available (%s); libhisstools_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return HB_ERR_CUDA;
    }
    if (device < 0 || device >= count)
    {
        set_error("device %d out of range (%d devices)", device, count);
        return HB_ERR_BAD_ARG;
    }
    HB_CUDA(cudaSetDevice(device));
    return HB_OK;
}

int make_twiddles(int dtype, int log2, void **d_out)
{
    if (log2 < 1) log2 = 1;
    const size_t half = size_t(1) << (log2 - 1);
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double n = (long double) (size_t(1) << log2);
    void *d = nullptr;
    if (dtype == HB_F64)
    {
        std::vector<Cx<double>> h(half);
        for (size_t q = 0; q < half; q++)
        {
            long double a = -2.0L * pi * (long double) q / n;
            h[q].x = (double) cosl(a);
            h[q].y = (double) sinl(a);
        }
        HB_CUDA(cudaMalloc(&d, half * sizeof(Cx<double>)));
        HB_CUDA(cudaMemcpy(d, h.data(), half * sizeof(Cx<double>), cudaMemcpyHostToDevice));
    }
    else
    {
        std::vector<Cx<float>> h(half);
        for (size_t q = 0; q < half; q++)
        {
            long double a = -2.0L * pi * (long double) q / n;
            h[q].x = (float) cosl(a);
            h[q].y = (float) sinl(a);
        }
        HB_CUDA(cudaMalloc(&d, half * sizeof(Cx<float>)));
        HB_CUDA(cudaMemcpy(d, h.data(), half * sizeof(Cx<float>), cudaMemcpyHostToDevice));
    }
    *d_out = d;
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// kernels (one CTA per transform; blockIdx.x = batch index)
// ---------------------------------------------------------------------------------------------

// complex forward (swap = 0) / planes-exchanged forward = unscaled inverse (swap = 1), in place or not
template <class T, int EPT>
__global__ void __launch_bounds__(512) k_cfft(const T *re_in, const T *im_in,
                                               T *re_out, T *im_out, int log2m, int swap,
                                               size_t stride, const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t M = 1u << log2m;
    const size_t base = size_t(blockIdx.x) * stride;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
    {
        T a = re_in[base + i], b = im_in[base + i];
        s[sidx<HB_PADSH>(i)] = swap ? cx<T>(b, a) : cx<T>(a, b);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, log2m, tw, tw_log2);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(i)];
        re_out[base + i] = swap ? v.y : v.x;
        im_out[base + i] = swap ? v.x : v.y;
    }
}

// real forward.  Source is either split planes already holding the de-interleaved signal (x == nullptr;
// HISSTools_FFT.h:154,166) or a real array of in_length samples that is de-interleaved and zero padded on
// the way in (HISSTools_FFT.h:180-208 = unzip_zero, Core:1258-1287, + rfft).  TI = element type of x.
template <class T, class TI, int EPT>
__global__ void __launch_bounds__(512) k_rfft(const TI *__restrict__ x, size_t in_length, size_t x_stride,
                                               const T *re_in, const T *im_in,
                                               T *re_out, T *im_out, size_t stride, int log2n,
                                               const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const int log2m = log2n - 1;
    const uint32_t M = 1u << log2m;
    const size_t base = size_t(blockIdx.x) * stride;
    if (x)
    {
        const TI *xb = x + size_t(blockIdx.x) * x_stride;
        const size_t n = size_t(1) << log2n;
        const size_t len = in_length < n ? in_length : n;
        const size_t pairs = len >> 1;
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
        {
            T a = 0, b = 0;
            if (i < pairs) { a = (T) xb[2 * size_t(i)]; b = (T) xb[2 * size_t(i) + 1]; }
            else if (i == pairs && (len & 1)) a = (T) xb[len - 1];
            s[sidx<HB_PADSH>(i)] = cx<T>(a, b);
        }
    }
    else
    {
        for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) s[sidx<HB_PADSH>(i)] = cx<T>(re_in[base + i], im_in[base + i]);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, log2m, tw, tw_log2);
    block_real_split<T, EPT, HB_PADSH>(s, M, log2n, false, tw, tw_log2);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(i)];
        re_out[base + i] = v.x;
        im_out[base + i] = v.y;
    }
}

// real inverse.  Result is left de-interleaved in the planes (even samples in re, odd in im:
// HISSTools_FFT.h:244,256) and, when y != nullptr, also interleaved into the real array y
// (HISSTools_FFT.h:269,282 = rifft + zip, Core:1228-1254).
template <class T, int EPT>
__global__ void __launch_bounds__(512) k_rifft(const T *re_in, const T *im_in,
                                                T *re_out, T *im_out, size_t stride,
                                                T *__restrict__ y, size_t y_stride, int log2n,
                                                const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const int log2m = log2n - 1;
    const uint32_t M = 1u << log2m;
    const size_t base = size_t(blockIdx.x) * stride;
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x) s[sidx<HB_PADSH>(i)] = cx<T>(re_in[base + i], im_in[base + i]);
    __syncthreads();
    block_real_split<T, EPT, HB_PADSH>(s, M, log2n, true, tw, tw_log2);
    __syncthreads();
    // hisstools_ifft = forward transform on exchanged planes (Core:1341-1346)
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(i)];
        s[sidx<HB_PADSH>(i)] = cx<T>(v.y, v.x);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, log2m, tw, tw_log2);
    for (uint32_t i = threadIdx.x; i < M; i += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(i)];
        if (re_out) { re_out[base + i] = v.y; im_out[base + i] = v.x; }
        if (y)
        {
            T *yb = y + size_t(blockIdx.x) * y_stride;
            yb[2 * size_t(i)] = v.y;
            yb[2 * size_t(i) + 1] = v.x;
        }
    }
}

// ---- sizes above the single-CTA limit: planes / real arrays <-> the interleaved array of hb_fft_big.cuh ----

// planes -> Z (exchange = 1 swaps the planes: the unscaled inverse is a forward transform of them, Core:1341-1346)
template <class T>
__global__ void k_big_pack_planes(const T *__restrict__ re, const T *__restrict__ im, size_t stride, Cx<T> *__restrict__ z, int m, int exchange)
{
    const size_t M = size_t(1) << m;
    const size_t b = blockIdx.y;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < M; i += size_t(gridDim.x) * blockDim.x)
    {
        const T a = re[b * stride + i], c = im[b * stride + i];
        z[b * M + i] = exchange ? cx<T>(c, a) : cx<T>(a, c);
    }
}

// real array -> Z with de-interleave and zero padding (unzip_zero, Core:1258-1287)
template <class T, class TI>
__global__ void k_big_pack_real(const TI *__restrict__ x, size_t in_length, size_t x_stride, Cx<T> *__restrict__ z, int m)
{
    const size_t M = size_t(1) << m;
    const size_t b = blockIdx.y;
    const TI *xb = x + b * x_stride;
    const size_t len = in_length < 2 * M ? in_length : 2 * M, pairs = len >> 1;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < M; i += size_t(gridDim.x) * blockDim.x)
    {
        T a = 0, c = 0;
        if (i < pairs) { a = (T) xb[2 * i]; c = (T) xb[2 * i + 1]; }
        else if (i == pairs && (len & 1)) a = (T) xb[len - 1];
        z[b * M + i] = cx<T>(a, c);
    }
}

// Z -> planes and / or an interleaved real array (exchange = 1: the result of a planes-exchanged transform)
template <class T>
__global__ void k_big_unpack(const Cx<T> *__restrict__ z, int m, int exchange, T *__restrict__ re, T *__restrict__ im, size_t stride,
                             T *__restrict__ y, size_t y_stride)
{
    const size_t M = size_t(1) << m;
    const size_t b = blockIdx.y;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < M; i += size_t(gridDim.x) * blockDim.x)
    {
        const Cx<T> v = z[b * M + i];
        const T a = exchange ? v.y : v.x, c = exchange ? v.x : v.y;
        if (re) { re[b * stride + i] = a; im[b * stride + i] = c; }
        if (y) { y[b * y_stride + 2 * i] = a; y[b * y_stride + 2 * i + 1] = c; }
    }
}

template <class T>
static int big_cfft_planes(BigScratch *bs, const T *re_in, const T *im_in, T *re_out, T *im_out, int log2m, int swap, size_t batch, size_t stride,
                           const Cx<T> *tw, int tw_log2, cudaStream_t st)
{
    if (!bs) { set_error("complex FFT of 2^%d points needs scratch memory (use a setup-based entry point)", log2m); return HB_ERR_UNSUPPORTED; }
    int rc = bs->ensure<T>(log2m, batch);
    if (rc) return rc;
    Cx<T> *z1 = (Cx<T> *) bs->z1.p, *z2 = (Cx<T> *) bs->z2.p;
    k_big_pack_planes<T><<<big_grid(log2m, batch), 256, 0, st>>>(re_in, im_in, stride, z1, log2m, swap);
    HB_LAUNCH_CHECK();
    if ((rc = big_cfft<T>(z1, z2, z1, log2m, batch, tw, tw_log2, st))) return rc;
    k_big_unpack<T><<<big_grid(log2m, batch), 256, 0, st>>>(z1, log2m, swap, re_out, im_out, stride, (T *) nullptr, 0);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T, class TI>
static int big_rfft(BigScratch *bs, const TI *x, size_t in_length, size_t x_stride, const T *re_in, const T *im_in, T *re_out, T *im_out,
                    size_t stride, int log2n, size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st)
{
    const int m = log2n - 1;
    if (!bs) { set_error("real FFT of 2^%d points needs scratch memory", log2n); return HB_ERR_UNSUPPORTED; }
    int rc = bs->ensure<T>(m, batch);
    if (rc) return rc;
    Cx<T> *z1 = (Cx<T> *) bs->z1.p, *z2 = (Cx<T> *) bs->z2.p;
    if (x) k_big_pack_real<T, TI><<<big_grid(m, batch), 256, 0, st>>>(x, in_length, x_stride, z1, m);
    else k_big_pack_planes<T><<<big_grid(m, batch), 256, 0, st>>>(re_in, im_in, stride, z1, m, 0);
    HB_LAUNCH_CHECK();
    if ((rc = big_cfft<T>(z1, z2, z1, m, batch, tw, tw_log2, st))) return rc;
    if ((rc = big_split<T>(z1, m, 0, batch, tw, tw_log2, st))) return rc;
    k_big_unpack<T><<<big_grid(m, batch), 256, 0, st>>>(z1, m, 0, re_out, im_out, stride, (T *) nullptr, 0);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
static int big_rifft(BigScratch *bs, const T *re_in, const T *im_in, T *re_out, T *im_out, size_t stride, T *y, size_t y_stride, int log2n,
                     size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st)
{
    const int m = log2n - 1;
    if (!bs) { set_error("real FFT of 2^%d points needs scratch memory", log2n); return HB_ERR_UNSUPPORTED; }
    int rc = bs->ensure<T>(m, batch);
    if (rc) return rc;
    Cx<T> *z1 = (Cx<T> *) bs->z1.p, *z2 = (Cx<T> *) bs->z2.p;
    k_big_pack_planes<T><<<big_grid(m, batch), 256, 0, st>>>(re_in, im_in, stride, z1, m, 0);
    HB_LAUNCH_CHECK();
    if ((rc = big_split<T>(z1, m, 1, batch, tw, tw_log2, st))) return rc;
    k_big_exchange<T><<<big_grid(m, batch).x, 256, 0, st>>>(z1, batch << m);
    HB_LAUNCH_CHECK();
    if ((rc = big_cfft<T>(z1, z2, z1, m, batch, tw, tw_log2, st))) return rc;
    k_big_unpack<T><<<big_grid(m, batch), 256, 0, st>>>(z1, m, 1, re_out, im_out, stride, y, y_stride);
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
template <class T> static size_t fft_smem_bytes(int log2m) { return size_t(padded_elems<HB_PADSH>(1u << log2m)) * sizeof(Cx<T>); }

template <class K> static int allow_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return HB_OK;
}

template <class T>
static int launch_cfft(const T *re_in, const T *im_in, T *re_out, T *im_out, int log2m, int swap, size_t batch, size_t stride,
                       const Cx<T> *tw, int tw_log2, cudaStream_t st, BigScratch *bs = nullptr)
{
    if (log2m > SmemFftLimit<T>::max_log2m) return big_cfft_planes<T>(bs, re_in, im_in, re_out, im_out, log2m, swap, batch, stride, tw, tw_log2, st);
    const size_t smem = fft_smem_bytes<T>(log2m);
    HB_EPT_DISPATCH(log2m,
        int rc = allow_smem(k_cfft<T, EPT>, smem); if (rc) return rc;
        k_cfft<T, EPT><<<(unsigned) batch, fft_threads(log2m, EPT), smem, st>>>(re_in, im_in, re_out, im_out, log2m, swap, stride, tw, tw_log2));
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// complex transform of `batch` split-plane arrays in device memory (forward, or the unscaled inverse when swap != 0:
// a forward transform of the exchanged planes, Core:1341-1346) -- the entry other translation units use
int cfft_planes(int dtype, const void *re_in, const void *im_in, void *re_out, void *im_out, int log2n, int swap, size_t batch, size_t stride,
                const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs)
{
    if (dtype == HB_F64)
        return launch_cfft<double>((const double *) re_in, (const double *) im_in, (double *) re_out, (double *) im_out, log2n, swap, batch, stride,
                                   (const Cx<double> *) tw, tw_log2, st, bs);
    return launch_cfft<float>((const float *) re_in, (const float *) im_in, (float *) re_out, (float *) im_out, log2n, swap, batch, stride,
                              (const Cx<float> *) tw, tw_log2, st, bs);
}

template <class T, class TI>
static int launch_rfft(const TI *x, size_t in_length, size_t x_stride, const T *re_in, const T *im_in, T *re_out, T *im_out,
                       size_t stride, int log2n, size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st, BigScratch *bs = nullptr)
{
    const int log2m = log2n - 1;
    if (log2m > SmemFftLimit<T>::max_log2m) return big_rfft<T, TI>(bs, x, in_length, x_stride, re_in, im_in, re_out, im_out, stride, log2n, batch, tw, tw_log2, st);
    const size_t smem = fft_smem_bytes<T>(log2m);
    HB_EPT_DISPATCH(log2m,
        int rc = allow_smem(k_rfft<T, TI, EPT>, smem); if (rc) return rc;
        k_rfft<T, TI, EPT><<<(unsigned) batch, fft_threads(log2m, EPT), smem, st>>>(x, in_length, x_stride, re_in, im_in, re_out, im_out, stride, log2n, tw, tw_log2));
    HB_LAUNCH_CHECK();
    return HB_OK;
}

template <class T>
static int launch_rifft(const T *re_in, const T *im_in, T *re_out, T *im_out, size_t stride, T *y, size_t y_stride, int log2n,
                        size_t batch, const Cx<T> *tw, int tw_log2, cudaStream_t st, BigScratch *bs = nullptr)
{
    const int log2m = log2n - 1;
    if (log2m > SmemFftLimit<T>::max_log2m) return big_rifft<T>(bs, re_in, im_in, re_out, im_out, stride, y, y_stride, log2n, batch, tw, tw_log2, st);
    const size_t smem = fft_smem_bytes<T>(log2m);
    HB_EPT_DISPATCH(log2m,
        int rc = allow_smem(k_rifft<T, EPT>, smem); if (rc) return rc;
        k_rifft<T, EPT><<<(unsigned) batch, fft_threads(log2m, EPT), smem, st>>>(re_in, im_in, re_out, im_out, stride, y, y_stride, log2n, tw, tw_log2));
    HB_LAUNCH_CHECK();
    return HB_OK;
}

// real transforms on split planes in device memory for other translation units (hb_spectral.cu): forward from a real
// array x (zero-padded from in_length; x == nullptr: in place on the planes), inverse to the planes and / or an
// interleaved real array y
int rfft_planes(int dtype, const void *x, size_t in_length, void *re, void *im, int log2n, const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs)
{
    if (dtype == HB_F64)
        return launch_rfft<double, double>((const double *) x, in_length, 0, (const double *) re, (const double *) im, (double *) re, (double *) im, 0, log2n, 1,
                                           (const Cx<double> *) tw, tw_log2, st, bs);
    return launch_rfft<float, float>((const float *) x, in_length, 0, (const float *) re, (const float *) im, (float *) re, (float *) im, 0, log2n, 1,
                                     (const Cx<float> *) tw, tw_log2, st, bs);
}
int rifft_planes(int dtype, void *re, void *im, void *y, int log2n, const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs)
{
    if (dtype == HB_F64)
        return launch_rifft<double>((const double *) re, (const double *) im, (double *) re, (double *) im, 0, (double *) y, 0, log2n, 1,
                                    (const Cx<double> *) tw, tw_log2, st, bs);
    return launch_rifft<float>((const float *) re, (const float *) im, (float *) re, (float *) im, 0, (float *) y, 0, log2n, 1,
                               (const Cx<float> *) tw, tw_log2, st, bs);
}

} // namespace hb

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace hb;

struct hb_fft_setup
{
    int dtype = HB_F32;
    int device = 0;
    int max_log2 = 0;
    int tw_log2 = 1;
    void *tw = nullptr;
    cudaStream_t stream = nullptr;
    DevBuf d_a, d_b, d_c;       // planes / real array scratch
    BigScratch big;             // interleaved arrays of the four-step path
    std::mutex lock;
};

extern "C" const char *hb_last_error(void) { return g_error; }
extern "C" uint64_t hb_launch_count(void) { return g_launches.load(); }
extern "C" const char *hb_version(void) { return "hisstools_b200 0.1 (sm_100a, CUDA " HB_STR(CUDART_VERSION) ")"; }

extern "C" int hb_fft_setup_create(hb_fft_setup **out, int dtype, uintptr_t max_fft_log2, int device)
{
    if (!out || (dtype != HB_F32 && dtype != HB_F64) || max_fft_log2 > 30) { set_error("hb_fft_setup_create: bad argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = use_device(device);
    if (rc) return rc;
    hb_fft_setup *s = new hb_fft_setup;
    s->dtype = dtype; s->device = device; s->max_log2 = (int) max_fft_log2;
    // the table serves complex transforms of 2^max and real ones of 2^max (which need order-2^max roots)
    // (sizes above the shared-memory limit take the four-step path: complex up to 2^22, real up to 2^23 points)
    int lim = BIG_MAX_LOG2 + 1;
    s->tw_log2 = s->max_log2 < 1 ? 1 : (s->max_log2 > lim ? lim : s->max_log2);
    rc = make_twiddles(dtype, s->tw_log2, &s->tw);
    if (rc) { delete s; return rc; }
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); cudaFree(s->tw); delete s; return HB_ERR_CUDA; }
    *out = s;
    return HB_OK;
}

extern "C" void hb_fft_setup_destroy(hb_fft_setup *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    s->d_a.release(); s->d_b.release(); s->d_c.release();
    s->big.release();
    cudaFree(s->tw);
    cudaStreamDestroy(s->stream);
    delete s;
}

namespace
{
enum Op { OP_FFT, OP_IFFT, OP_RFFT, OP_RIFFT };

template <class T>
int inplace_op(hb_fft_setup *s, Op op, T *re, T *im, uintptr_t log2n)
{
    const bool real = (op == OP_RFFT || op == OP_RIFFT);
    if (log2n == 0) return HB_OK;                       // hisstools_fft of one point is a no-op (Core:1328-1336)
    if ((int) log2n > s->max_log2) { set_error("log2n %d exceeds the setup's maximum %d", (int) log2n, s->max_log2); return HB_ERR_BAD_ARG; }
    if ((int) log2n - (real ? 1 : 0) > BIG_MAX_LOG2) { set_error("transforms above 2^%d complex points are not implemented", BIG_MAX_LOG2); return HB_ERR_UNSUPPORTED; }
    const size_t planes = real ? (size_t(1) << (log2n - 1)) : (size_t(1) << log2n);
    const size_t bytes = planes * sizeof(T);
    int rc;
    if ((rc = s->d_a.ensure(bytes)) || (rc = s->d_b.ensure(bytes))) return rc;
    T *d_re = (T *) s->d_a.p, *d_im = (T *) s->d_b.p;
    const Cx<T> *tw = (const Cx<T> *) s->tw;
    HB_CUDA(cudaMemcpyAsync(d_re, re, bytes, cudaMemcpyHostToDevice, s->stream));
    HB_CUDA(cudaMemcpyAsync(d_im, im, bytes, cudaMemcpyHostToDevice, s->stream));
    switch (op)
    {
        case OP_FFT:   rc = launch_cfft<T>(d_re, d_im, d_re, d_im, (int) log2n, 0, 1, 0, tw, s->tw_log2, s->stream, &s->big); break;
        case OP_IFFT:  rc = launch_cfft<T>(d_re, d_im, d_re, d_im, (int) log2n, 1, 1, 0, tw, s->tw_log2, s->stream, &s->big); break;
        case OP_RFFT:  rc = launch_rfft<T, T>(nullptr, 0, 0, d_re, d_im, d_re, d_im, 0, (int) log2n, 1, tw, s->tw_log2, s->stream, &s->big); break;
        case OP_RIFFT: rc = launch_rifft<T>(d_re, d_im, d_re, d_im, 0, nullptr, 0, (int) log2n, 1, tw, s->tw_log2, s->stream, &s->big); break;
    }
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(re, d_re, bytes, cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaMemcpyAsync(im, d_im, bytes, cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    return HB_OK;
}

int inplace_dispatch(hb_fft_setup *s, Op op, void *re, void *im, uintptr_t log2n)
{
    if (!s || !re || !im) { set_error("null argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? inplace_op<double>(s, op, (double *) re, (double *) im, log2n)
                              : inplace_op<float>(s, op, (float *) re, (float *) im, log2n);
}
} // namespace

extern "C" int hb_fft(hb_fft_setup *s, void *re, void *im, uintptr_t log2n) { return inplace_dispatch(s, OP_FFT, re, im, log2n); }
extern "C" int hb_ifft(hb_fft_setup *s, void *re, void *im, uintptr_t log2n) { return inplace_dispatch(s, OP_IFFT, re, im, log2n); }
extern "C" int hb_rfft(hb_fft_setup *s, void *re, void *im, uintptr_t log2n) { return inplace_dispatch(s, OP_RFFT, re, im, log2n); }
extern "C" int hb_rifft(hb_fft_setup *s, void *re, void *im, uintptr_t log2n) { return inplace_dispatch(s, OP_RIFFT, re, im, log2n); }

namespace
{
template <class T, class TI>
int rfft_real_host(hb_fft_setup *s, const TI *input, T *re, T *im, uintptr_t in_length, uintptr_t log2n)
{
    const size_t n = size_t(1) << log2n, half = n >> 1;
    const size_t len = in_length < n ? in_length : n;
    int rc;
    if ((rc = s->d_a.ensure(half * sizeof(T))) || (rc = s->d_b.ensure(half * sizeof(T))) || (rc = s->d_c.ensure((len ? len : 1) * sizeof(TI)))) return rc;
    if (len) HB_CUDA(cudaMemcpyAsync(s->d_c.p, input, len * sizeof(TI), cudaMemcpyHostToDevice, s->stream));
    rc = launch_rfft<T, TI>((const TI *) s->d_c.p, len, 0, nullptr, nullptr, (T *) s->d_a.p, (T *) s->d_b.p, 0, (int) log2n, 1,
                            (const Cx<T> *) s->tw, s->tw_log2, s->stream, &s->big);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(re, s->d_a.p, half * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaMemcpyAsync(im, s->d_b.p, half * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    return HB_OK;
}

template <class T>
int rifft_real_host(hb_fft_setup *s, T *re, T *im, T *output, uintptr_t log2n)
{
    const size_t n = size_t(1) << log2n, half = n >> 1;
    int rc;
    if ((rc = s->d_a.ensure(half * sizeof(T))) || (rc = s->d_b.ensure(half * sizeof(T))) || (rc = s->d_c.ensure(n * sizeof(T)))) return rc;
    HB_CUDA(cudaMemcpyAsync(s->d_a.p, re, half * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    HB_CUDA(cudaMemcpyAsync(s->d_b.p, im, half * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    rc = launch_rifft<T>((const T *) s->d_a.p, (const T *) s->d_b.p, (T *) s->d_a.p, (T *) s->d_b.p, 0, (T *) s->d_c.p, 0, (int) log2n, 1,
                         (const Cx<T> *) s->tw, s->tw_log2, s->stream, &s->big);
    if (rc) return rc;
    HB_CUDA(cudaMemcpyAsync(re, s->d_a.p, half * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaMemcpyAsync(im, s->d_b.p, half * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaMemcpyAsync(output, s->d_c.p, n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    return HB_OK;
}
} // namespace

extern "C" int hb_rfft_real(hb_fft_setup *s, const void *input, int in_dtype, void *re, void *im, uintptr_t in_length, uintptr_t log2n)
{
    if (!s || !re || !im || (!input && in_length) || log2n < 1) { set_error("hb_rfft_real: bad argument"); return HB_ERR_BAD_ARG; }
    if ((int) log2n > s->max_log2) { set_error("log2n %d exceeds the setup's maximum %d", (int) log2n, s->max_log2); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    if (s->dtype == HB_F64)
        return in_dtype == HB_F32 ? rfft_real_host<double, float>(s, (const float *) input, (double *) re, (double *) im, in_length, log2n)
                                  : rfft_real_host<double, double>(s, (const double *) input, (double *) re, (double *) im, in_length, log2n);
    if (in_dtype != HB_F32) { set_error("double input needs a double setup"); return HB_ERR_BAD_ARG; }
    return rfft_real_host<float, float>(s, (const float *) input, (float *) re, (float *) im, in_length, log2n);
}

extern "C" int hb_rifft_real(hb_fft_setup *s, void *re, void *im, void *output, uintptr_t log2n)
{
    if (!s || !re || !im || !output || log2n < 1) { set_error("hb_rifft_real: bad argument"); return HB_ERR_BAD_ARG; }
    if ((int) log2n > s->max_log2) { set_error("log2n %d exceeds the setup's maximum %d", (int) log2n, s->max_log2); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? rifft_real_host<double>(s, (double *) re, (double *) im, (double *) output, log2n)
                              : rifft_real_host<float>(s, (float *) re, (float *) im, (float *) output, log2n);
}

extern "C" int hb_rfft_real_batched_dev(hb_fft_setup *s, const void *d_input, void *d_re, void *d_im, uintptr_t log2n, uintptr_t batch,
                                        uintptr_t in_stride, uintptr_t out_stride, void *stream)
{
    if (!s || !d_input || !d_re || !d_im || log2n < 1 || (int) log2n > s->max_log2) { set_error("hb_rfft_real_batched_dev: bad argument"); return HB_ERR_BAD_ARG; }
    int rc = use_device(s->device);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t) stream : s->stream;
    const size_t n = size_t(1) << log2n;
    if (s->dtype == HB_F64)
        return launch_rfft<double, double>((const double *) d_input, n, in_stride, nullptr, nullptr, (double *) d_re, (double *) d_im, out_stride,
                                           (int) log2n, batch, (const Cx<double> *) s->tw, s->tw_log2, st, &s->big);
    return launch_rfft<float, float>((const float *) d_input, n, in_stride, nullptr, nullptr, (float *) d_re, (float *) d_im, out_stride,
                                     (int) log2n, batch, (const Cx<float> *) s->tw, s->tw_log2, st, &s->big);
}

extern "C" int hb_rifft_real_batched_dev(hb_fft_setup *s, const void *d_re, const void *d_im, void *d_output, uintptr_t log2n, uintptr_t batch,
                                         uintptr_t in_stride, uintptr_t out_stride, void *stream)
{
    if (!s || !d_output || !d_re || !d_im || log2n < 1 || (int) log2n > s->max_log2) { set_error("hb_rifft_real_batched_dev: bad argument"); return HB_ERR_BAD_ARG; }
    int rc = use_device(s->device);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t) stream : s->stream;
    if (s->dtype == HB_F64)
        return launch_rifft<double>((const double *) d_re, (const double *) d_im, nullptr, nullptr, in_stride, (double *) d_output, out_stride,
                                    (int) log2n, batch, (const Cx<double> *) s->tw, s->tw_log2, st, &s->big);
    return launch_rifft<float>((const float *) d_re, (const float *) d_im, nullptr, nullptr, in_stride, (float *) d_output, out_stride,
                               (int) log2n, batch, (const Cx<float> *) s->tw, s->tw_log2, st, &s->big);
}
