// hb_spectral.cu -- one-shot FFT convolution of two real buffers: the hb_spectral_* entry points of
// include/hisstools_b200.h.
//
// Replaces spectral_processor<T>::convolve(T*, in_ptr, in_ptr, EdgeMode) of the reference
// (SpectralProcessor.hpp:169-172 -> binary_op :616-674, op_sizes :323-354, arrange_convolve :445-481)
// and the per-bin product it uses (SpectralFunctions.hpp: ir_convolve_real :420-424, real_operation
// :63-84, impl::convolve :274-281).  Two launches:
//   k_spec_fwd  2 CTAs: zero-padded (or, for the fold modes, mirror-extended) inputs -> real FFT -> packed spectra
//   k_spec_inv  1 CTA : scaled per-bin product (DC and Nyquist as two real products) -> inverse real FFT
//                       -> the edge-mode arrangement (copy / wrap-add / offset copy) straight into the output
#include "hb_common.cuh"
#include "hb_fft_block.cuh"
#include "hb_fft_big.cuh"

#include <algorithm>
#include <mutex>

using namespace hb;

namespace
{
enum { MODE_LINEAR = 0, MODE_WRAP = 1, MODE_WRAP_CENTRE = 2, MODE_FOLD = 3, MODE_FOLD_REPEAT = 4 };

struct SpecGeom
{
    uint32_t log2n;            // FFT size
    uint32_t n1, n2;           // input lengths
    uint32_t mn, mx;           // min / max of them
    uint32_t fold_size;        // mirrored samples each side of the longer input (fold modes)
    int mode;
    int first_is_big;          // n1 >= n2
};

// sample j of the FFT input of operand `which` (0: first / extended operand, 1: second)
template <class T>
__device__ __forceinline__ T spec_fetch(const SpecGeom &g, const T *__restrict__ in1, const T *__restrict__ in2, int which, uint32_t j)
{
    const bool fold = g.mode == MODE_FOLD || g.mode == MODE_FOLD_REPEAT;
    if (!fold)
    {
        if (which == 0) return j < g.n1 ? in1[j] : T(0);
        return j < g.n2 ? in2[j] : T(0);
    }
    // SpectralProcessor.hpp:624-641: the longer input is mirror-extended by fold_size on both sides
    const T *big = g.first_is_big ? in1 : in2, *small = g.first_is_big ? in2 : in1;
    if (which == 1) return j < g.mn ? small[j] : T(0);
    const uint32_t off = g.mode == MODE_FOLD_REPEAT ? 0 : 1;
    if (j < g.fold_size) return big[off + g.fold_size - 1 - j];
    if (j < g.fold_size + g.mx) return big[j - g.fold_size];
    if (j < g.mx + 2 * g.fold_size) return big[g.mx - off - 1 - (j - g.fold_size - g.mx)];
    return T(0);
}

template <class T, int EPT>
__global__ void __launch_bounds__(512) k_spec_fwd(const SpecGeom g, const T *__restrict__ in1, const T *__restrict__ in2,
                                                   Cx<T> *__restrict__ spectra, const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const int which = blockIdx.x;
    const uint32_t M = 1u << (g.log2n - 1);
    for (uint32_t k = threadIdx.x; k < M; k += blockDim.x)
        s[sidx<HB_PADSH>(k)] = cx<T>(spec_fetch<T>(g, in1, in2, which, 2 * k), spec_fetch<T>(g, in1, in2, which, 2 * k + 1));
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    block_real_split<T, EPT, HB_PADSH>(s, M, (int) g.log2n, false, tw, tw_log2);
    __syncthreads();
    Cx<T> *dst = spectra + size_t(which) * M;
    for (uint32_t k = threadIdx.x; k < M; k += blockDim.x) dst[k] = s[sidx<HB_PADSH>(k)];
}

template <class T, int EPT>
__global__ void __launch_bounds__(512) k_spec_inv(const SpecGeom g, const Cx<T> *__restrict__ spectra, T *__restrict__ out,
                                                   const Cx<T> *__restrict__ tw, int tw_log2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *s = reinterpret_cast<Cx<T> *>(smem_raw);
    const uint32_t M = 1u << (g.log2n - 1);
    const T scale = T(0.25) / T(size_t(1) << g.log2n);                 // SpectralProcessor.hpp:643
    const Cx<T> *A = spectra, *B = spectra + M;
    for (uint32_t k = threadIdx.x; k < M; k += blockDim.x)
    {
        const Cx<T> a = A[k], b = B[k];
        // SpectralFunctions.hpp:63-84: bin 0 carries DC and Nyquist, both real; :274-281 for the rest
        Cx<T> r = k ? cx<T>(scale * (a.x * b.x - a.y * b.y), scale * (a.y * b.x + a.x * b.y))
                    : cx<T>(scale * (a.x * b.x), scale * (a.y * b.y));
        s[sidx<HB_PADSH>(k)] = r;
    }
    __syncthreads();
    block_real_split<T, EPT, HB_PADSH>(s, M, (int) g.log2n, true, tw, tw_log2);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < M; k += blockDim.x)
    {
        Cx<T> v = s[sidx<HB_PADSH>(k)];
        s[sidx<HB_PADSH>(k)] = cx<T>(v.y, v.x);
    }
    __syncthreads();
    block_fft<T, EPT, HB_PADSH>(s, (int) g.log2n - 1, tw, tw_log2);
    // time sample i of the linear result: even samples sit in .y, odd ones in .x (planes exchanged)
    auto lin = [&](uint32_t i) -> T { const Cx<T> v = s[sidx<HB_PADSH>(i >> 1)]; return (i & 1) ? v.x : v.y; };
    const uint32_t min_m1 = g.mn - 1;
    const uint32_t result = g.mode == MODE_LINEAR ? g.n1 + g.n2 - 1 : g.mx;
    for (uint32_t i = threadIdx.x; i < result; i += blockDim.x)
    {
        T v;
        switch (g.mode)
        {
            case MODE_LINEAR:
                v = lin(i);
                break;
            case MODE_WRAP:                                             // copy + wrap (SpectralProcessor.hpp:421-435,445-481)
                v = lin(i);
                if (i < min_m1) v += lin(g.mx + i);
                break;
            case MODE_WRAP_CENTRE:
            {
                const uint32_t wrapped = min_m1 >> 1;
                v = lin(wrapped + i);
                if (i < min_m1 - wrapped) v += lin(g.mx + wrapped + i);
                if (i >= g.mx - wrapped) v += lin(i - (g.mx - wrapped));
                break;
            }
            default:                                                    // Fold / FoldRepeat: offset copy
                v = lin(min_m1 + i);
                break;
        }
        out[i] = v;
    }
}

// ---- sizes above the single-CTA limit: the same three stages over global memory (hb_fft_big.cuh) ----
template <class T>
__global__ void k_spec_big_pack(const SpecGeom g, const T *__restrict__ in1, const T *__restrict__ in2, Cx<T> *__restrict__ z)
{
    const uint32_t M = 1u << (g.log2n - 1);
    const int which = blockIdx.y;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x)
        z[size_t(which) * M + k] = cx<T>(spec_fetch<T>(g, in1, in2, which, 2 * k), spec_fetch<T>(g, in1, in2, which, 2 * k + 1));
}

template <class T>
__global__ void k_spec_big_mul(const SpecGeom g, Cx<T> *__restrict__ z)
{
    const uint32_t M = 1u << (g.log2n - 1);
    const T scale = T(0.25) / T(size_t(1) << g.log2n);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x)
    {
        const Cx<T> a = z[k], b = z[size_t(M) + k];
        z[k] = k ? cx<T>(scale * (a.x * b.x - a.y * b.y), scale * (a.y * b.x + a.x * b.y))
                 : cx<T>(scale * (a.x * b.x), scale * (a.y * b.y));
    }
}

// the edge-mode arrangement of k_spec_inv, reading the planes-exchanged transform from global memory
template <class T>
__global__ void k_spec_big_arrange(const SpecGeom g, const Cx<T> *__restrict__ z, T *__restrict__ out)
{
    auto lin = [&](uint32_t i) -> T { const Cx<T> v = z[i >> 1]; return (i & 1) ? v.x : v.y; };
    const uint32_t min_m1 = g.mn - 1;
    const uint32_t result = g.mode == MODE_LINEAR ? g.n1 + g.n2 - 1 : g.mx;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < result; i += gridDim.x * blockDim.x)
    {
        T v;
        switch (g.mode)
        {
            case MODE_LINEAR:
                v = lin(i);
                break;
            case MODE_WRAP:
                v = lin(i);
                if (i < min_m1) v += lin(g.mx + i);
                break;
            case MODE_WRAP_CENTRE:
            {
                const uint32_t wrapped = min_m1 >> 1;
                v = lin(wrapped + i);
                if (i < min_m1 - wrapped) v += lin(g.mx + wrapped + i);
                if (i >= g.mx - wrapped) v += lin(i - (g.mx - wrapped));
                break;
            }
            default:
                v = lin(min_m1 + i);
                break;
        }
        out[i] = v;
    }
}

uint32_t ceil_log2(uintptr_t value)
{
    uint32_t bits = 0;
    while ((uintptr_t(1) << bits) < value) bits++;
    return bits;
}

template <class K> int allow_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) HB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    return HB_OK;
}
} // namespace

struct hb_spectral
{
    int dtype = HB_F32, device = 0;
    uint32_t max_log2 = 0;          // calc_fft_size_log2 of the requested maximum (SpectralProcessor.hpp:96-108)
    void *tw = nullptr;
    int tw_log2 = 1;
    cudaStream_t stream = nullptr;
    DevBuf d_in1, d_in2, d_spec, d_out;
    BigScratch big;
    std::mutex lock;
};

namespace
{
// op_sizes (SpectralProcessor.hpp:323-354): lengths and FFT size of one operation; 0 when it cannot run
struct Sizes { uintptr_t mn, mx, linear, fold_size, need, result; uint32_t log2n; };

bool plan_sizes(uintptr_t n1, uintptr_t n2, int mode, uint32_t max_log2, Sizes &z)
{
    if (!n1 || !n2) return false;
    z.mn = n1 < n2 ? n1 : n2;
    z.mx = n1 < n2 ? n2 : n1;
    z.linear = n1 + n2 - 1;
    const bool fold = mode == MODE_FOLD || mode == MODE_FOLD_REPEAT;
    z.fold_size = fold ? z.mn >> 1 : 0;
    z.need = fold ? z.mx + 2 * z.fold_size + (z.mn - 1) : z.linear;
    z.log2n = ceil_log2(z.need);
    z.result = mode == MODE_LINEAR ? z.linear : z.mx;
    return z.log2n <= max_log2;
}

int set_max(hb_spectral *s, uintptr_t max_fft_size)
{
    const uint32_t l2 = ceil_log2(max_fft_size ? max_fft_size : 1);
    const int lim = BIG_MAX_LOG2 + 1;
    if ((int) l2 > lim)
    {
        set_error("maximum FFT size 2^%u is beyond what this build implements (2^%d)", l2, lim);
        return HB_ERR_UNSUPPORTED;
    }
    if (s->tw && (int) l2 <= s->tw_log2) { s->max_log2 = l2; return HB_OK; }
    cudaFree(s->tw);
    s->tw = nullptr;
    s->tw_log2 = l2 < 1 ? 1 : (int) l2;
    int rc = make_twiddles(s->dtype, s->tw_log2, &s->tw);
    if (rc) return rc;
    s->max_log2 = l2;
    return HB_OK;
}

template <class T>
int convolve(hb_spectral *s, T *output, const T *in1, uintptr_t n1, const T *in2, uintptr_t n2, int mode, uintptr_t *written)
{
    Sizes z;
    if (written) *written = 0;
    if (!plan_sizes(n1, n2, mode, s->max_log2, z)) return HB_OK;               // silently does nothing (SpectralProcessor.hpp:651-652)
    if (n1 == 1 && n2 == 1)                                                     // :656-660
    {
        output[0] = in1[0] * in2[0];
        if (written) *written = 1;
        return HB_OK;
    }
    const uint32_t log2n = z.log2n < 2 ? 2 : z.log2n;                           // smallest transform the kernels run: 4 points
    const uint32_t log2m = log2n - 1;
    const size_t M = size_t(1) << log2m;
    int rc;
    if ((rc = s->d_in1.ensure(n1 * sizeof(T))) || (rc = s->d_in2.ensure(n2 * sizeof(T))) ||
        (rc = s->d_spec.ensure(2 * M * sizeof(Cx<T>))) || (rc = s->d_out.ensure(z.result * sizeof(T)))) return rc;
    HB_CUDA(cudaMemcpyAsync(s->d_in1.p, in1, n1 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    HB_CUDA(cudaMemcpyAsync(s->d_in2.p, in2, n2 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    SpecGeom g;
    g.log2n = log2n; g.n1 = (uint32_t) n1; g.n2 = (uint32_t) n2; g.mn = (uint32_t) z.mn; g.mx = (uint32_t) z.mx;
    g.fold_size = (uint32_t) z.fold_size; g.mode = mode; g.first_is_big = n1 >= n2;
    const size_t smem = size_t(padded_elems<HB_PADSH>((uint32_t) M)) * sizeof(Cx<T>);
    const Cx<T> *tw = (const Cx<T> *) s->tw;
    if ((int) log2m > SmemFftLimit<T>::max_log2m)
    {
        if ((rc = s->big.ensure<T>((int) log2m, 2))) return rc;
        Cx<T> *z1 = (Cx<T> *) s->big.z1.p, *z2 = (Cx<T> *) s->big.z2.p;
        const unsigned blocks = (unsigned) std::min<size_t>((M + 255) / 256, 1024);
        k_spec_big_pack<T><<<dim3(blocks, 2), 256, 0, s->stream>>>(g, (const T *) s->d_in1.p, (const T *) s->d_in2.p, z1);
        HB_LAUNCH_CHECK();
        if ((rc = big_cfft<T>(z1, z2, z1, (int) log2m, 2, tw, s->tw_log2, s->stream))) return rc;
        if ((rc = big_split<T>(z1, (int) log2m, 0, 2, tw, s->tw_log2, s->stream))) return rc;
        k_spec_big_mul<T><<<blocks, 256, 0, s->stream>>>(g, z1);
        HB_LAUNCH_CHECK();
        if ((rc = big_split<T>(z1, (int) log2m, 1, 1, tw, s->tw_log2, s->stream))) return rc;
        k_big_exchange<T><<<blocks, 256, 0, s->stream>>>(z1, M);
        HB_LAUNCH_CHECK();
        if ((rc = big_cfft<T>(z1, z2, z1, (int) log2m, 1, tw, s->tw_log2, s->stream))) return rc;
        k_spec_big_arrange<T><<<(unsigned) std::min<size_t>((z.result + 255) / 256, 1024), 256, 0, s->stream>>>(g, z1, (T *) s->d_out.p);
        HB_LAUNCH_CHECK();
    }
    else
    HB_EPT_DISPATCH(log2m,
        if ((rc = allow_smem(k_spec_fwd<T, EPT>, smem)) || (rc = allow_smem(k_spec_inv<T, EPT>, smem))) return rc;
        k_spec_fwd<T, EPT><<<2, fft_threads(log2m, EPT), smem, s->stream>>>(g, (const T *) s->d_in1.p, (const T *) s->d_in2.p, (Cx<T> *) s->d_spec.p, tw, s->tw_log2);
        count_launch();
        k_spec_inv<T, EPT><<<1, fft_threads(log2m, EPT), smem, s->stream>>>(g, (const Cx<T> *) s->d_spec.p, (T *) s->d_out.p, tw, s->tw_log2));
    HB_LAUNCH_CHECK();
    HB_CUDA(cudaMemcpyAsync(output, s->d_out.p, z.result * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    if (written) *written = z.result;
    return HB_OK;
}
} // namespace

extern "C" int hb_spectral_create(hb_spectral **out, int dtype, uintptr_t max_fft_size, int device)
{
    if (!out || (dtype != HB_F32 && dtype != HB_F64)) { set_error("hb_spectral_create: bad argument"); return HB_ERR_BAD_ARG; }
    *out = nullptr;
    int rc = use_device(device);
    if (rc) return rc;
    hb_spectral *s = new hb_spectral;
    s->dtype = dtype; s->device = device;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete s; return HB_ERR_CUDA; }
    rc = set_max(s, max_fft_size);
    if (rc) { cudaStreamDestroy(s->stream); delete s; return rc; }
    *out = s;
    return HB_OK;
}

extern "C" void hb_spectral_destroy(hb_spectral *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    s->d_in1.release(); s->d_in2.release(); s->d_spec.release(); s->d_out.release();
    s->big.release();
    cudaFree(s->tw);
    cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int hb_spectral_set_max_fft_size(hb_spectral *s, uintptr_t max_fft_size)
{
    if (!s) { set_error("null handle"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return set_max(s, max_fft_size);
}

extern "C" uintptr_t hb_spectral_max_fft_size(const hb_spectral *s) { return s ? uintptr_t(1) << s->max_log2 : 0; }

extern "C" uintptr_t hb_spectral_convolved_size(const hb_spectral *s, uintptr_t n1, uintptr_t n2, int mode)
{
    Sizes z;
    if (!s || mode < MODE_LINEAR || mode > MODE_FOLD_REPEAT || !plan_sizes(n1, n2, mode, s->max_log2, z)) return 0;
    return z.result;
}

extern "C" int hb_spectral_convolve(hb_spectral *s, void *output, const void *in1, uintptr_t n1, const void *in2, uintptr_t n2, int mode, uintptr_t *written)
{
    if (!s || mode < MODE_LINEAR || mode > MODE_FOLD_REPEAT || ((!output || !in1 || !in2) && n1 && n2))
    {
        set_error("hb_spectral_convolve: bad argument");
        return HB_ERR_BAD_ARG;
    }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? convolve<double>(s, (double *) output, (const double *) in1, n1, (const double *) in2, n2, mode, written)
                              : convolve<float>(s, (float *) output, (const float *) in1, n1, (const float *) in2, n2, mode, written);
}
