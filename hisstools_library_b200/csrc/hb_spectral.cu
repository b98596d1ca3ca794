// hb_spectral.cu -- one-shot FFT convolution and correlation of two buffers: the hb_spectral_* entry points of
// include/hisstools_b200.h.
//
// Replaces spectral_processor<T>::convolve / correlate of the reference, real inputs (SpectralProcessor.hpp:169-172,
// 181-184 -> binary_op :616-674) and complex inputs (:164-167, 176-179 -> binary_op :559-614), with op_sizes :323-354,
// arrange_convolve :445-481, arrange_correlate :483-538 and the per-bin products they use (SpectralFunctions.hpp:
// real_operation :63-84, complex_operation :49-61, impl::correlate :265-272, impl::convolve :274-281).
// The edge-mode arrangement is evaluated as a list of "terms" (Arrange below) built on the host from the reference's
// copy / wrap / zero calls.  Real inputs, two launches:
//   k_spec_fwd  2 CTAs: zero-padded (or, for the fold modes, mirror-extended) inputs -> real FFT -> packed spectra
//   k_spec_inv  1 CTA : scaled per-bin product (DC and Nyquist as two real products) -> inverse real FFT
//                       -> the edge-mode arrangement (copy / wrap-add / offset copy) straight into the output
#include "hb_common.cuh"
#include "hb_fft_block.cuh"
#include "hb_fft_big.cuh"

#include <algorithm>
#include <mutex>

using namespace hb;

namespace hb
{
int cfft_planes(int dtype, const void *re_in, const void *im_in, void *re_out, void *im_out, int log2n, int swap, size_t batch, size_t stride,
                const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs);          // hb_fft.cu
int rfft_planes(int dtype, const void *x, size_t in_length, void *re, void *im, int log2n, const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs);
int rifft_planes(int dtype, void *re, void *im, void *y, int log2n, const void *tw, int tw_log2, cudaStream_t st, BigScratch *bs);
}

namespace
{
enum { MODE_LINEAR = 0, MODE_WRAP = 1, MODE_WRAP_CENTRE = 2, MODE_FOLD = 3, MODE_FOLD_REPEAT = 4 };

// out[i] = sum over terms t with lo <= i < hi of L(src + i - lo), L = the result of the inverse transform (0 at or past
// the FFT size).  The reference's copy() calls never overlap and its zero() ranges are the gaps, so "assign" and "add"
// coincide and a gap reads as 0.
struct Arrange
{
    uint32_t n;
    uint32_t lo[5], hi[5], src[5];
    uint32_t end_rel[5];       // the source index counts back from the end of the transform (negative lags)
    uint32_t result;           // samples written
    uint32_t fft_ref;          // the reference's transform size: a read at or past it contributes 0
    uint32_t shift;            // this library's transform size minus fft_ref (non-zero only below 4 points)
};

struct SpecGeom
{
    uint32_t log2n;            // FFT size
    uint32_t n1, n2;           // input lengths
    uint32_t mn, mx;           // min / max of them
    uint32_t fold_size;        // mirrored samples each side of the longer input (fold modes)
    int mode;
    int first_is_big;          // n1 >= n2
    int op;                    // 0 convolve, 1 correlate
    Arrange arr;
};

template <class F>
__device__ __forceinline__ auto arrange_eval(const Arrange &a, uint32_t i, F lin) -> decltype(lin(0u))
{
    decltype(lin(0u)) v = 0;
    for (uint32_t t = 0; t < a.n; t++)
        if (i >= a.lo[t] && i < a.hi[t])
        {
            const uint32_t j = a.src[t] + (i - a.lo[t]);
            if (j < a.fft_ref) v += lin(a.end_rel[t] ? j + a.shift : j);
        }
    return v;
}

// sample j of the FFT input of input `which` (0: in1, 1: in2): zero-padded, and for the fold modes the longer input
// mirror-extended by fold_size on both sides (SpectralProcessor.hpp:356-372, 624-641)
template <class T>
__device__ __forceinline__ T fetch_padded(const T *__restrict__ in, uint32_t n, uint32_t size, uint32_t fold_size, uint32_t off, uint32_t j)
{
    // plane of `n` samples padded with zeros to `size`, then extended: [mirror | plane | mirror | zeros]
    uint32_t k;
    if (j < fold_size) k = off + fold_size - 1 - j;
    else if (j < fold_size + size) k = j - fold_size;
    else if (j < size + 2 * fold_size) k = size - off - 1 - (j - fold_size - size);
    else return T(0);
    return k < n ? in[k] : T(0);
}

template <class T>
__device__ __forceinline__ T spec_fetch(const SpecGeom &g, const T *__restrict__ in1, const T *__restrict__ in2, int which, uint32_t j)
{
    const bool fold = g.mode == MODE_FOLD || g.mode == MODE_FOLD_REPEAT;
    const bool is_big = (which == 0) == (g.first_is_big != 0);
    const uint32_t n = which ? g.n2 : g.n1;
    return fetch_padded<T>(which ? in2 : in1, n, n, (fold && is_big) ? g.fold_size : 0u, g.mode == MODE_FOLD_REPEAT ? 0u : 1u, j);
}

// scaled per-bin product of SpectralFunctions.hpp:265-281: a * b (convolve) or a * conj(b) (correlate)
template <class T>
__device__ __forceinline__ Cx<T> spec_op(int op, T scale, const Cx<T> a, const Cx<T> b)
{
    return op ? cx<T>(scale * (a.x * b.x + a.y * b.y), scale * (a.y * b.x - a.x * b.y))
              : cx<T>(scale * (a.x * b.x - a.y * b.y), scale * (a.y * b.x + a.x * b.y));
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
        // SpectralFunctions.hpp:63-84: bin 0 carries DC and Nyquist, both real (for either operation); :265-281 for the rest
        s[sidx<HB_PADSH>(k)] = k ? spec_op<T>(g.op, scale, a, b) : cx<T>(scale * (a.x * b.x), scale * (a.y * b.y));
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
    for (uint32_t i = threadIdx.x; i < g.arr.result; i += blockDim.x) out[i] = arrange_eval(g.arr, i, lin);
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
        z[k] = k ? spec_op<T>(g.op, scale, a, b) : cx<T>(scale * (a.x * b.x), scale * (a.y * b.y));
    }
}

// the edge-mode arrangement of k_spec_inv, reading the planes-exchanged transform from global memory
template <class T>
__global__ void k_spec_big_arrange(const SpecGeom g, const Cx<T> *__restrict__ z, T *__restrict__ out)
{
    auto lin = [&](uint32_t i) -> T { const Cx<T> v = z[i >> 1]; return (i & 1) ? v.x : v.y; };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.arr.result; i += gridDim.x * blockDim.x)
        out[i] = arrange_eval(g.arr, i, lin);
}

// ---- complex inputs: split planes in global memory, the transforms through hb::cfft_planes (single CTA or four-step) ----
// planes of operand `which` (blockIdx.y): [re | im], each padded to the FFT size; input planes may be shorter than the
// operand (or absent), copy_fold_zero of SpectralProcessor.hpp:386-399
template <class T>
__global__ void k_cspec_pack(const SpecGeom g, const T *__restrict__ r1, uint32_t nr1, const T *__restrict__ i1, uint32_t ni1,
                             const T *__restrict__ r2, uint32_t nr2, const T *__restrict__ i2, uint32_t ni2, T *__restrict__ planes)
{
    const uint32_t N = 1u << g.log2n;
    const int which = blockIdx.y;
    const bool fold = g.mode == MODE_FOLD || g.mode == MODE_FOLD_REPEAT;
    const bool is_big = (which == 0) == (g.first_is_big != 0);
    const uint32_t fs = (fold && is_big) ? g.fold_size : 0u, off = g.mode == MODE_FOLD_REPEAT ? 0u : 1u;
    const uint32_t size = which ? g.n2 : g.n1;
    const T *re = which ? r2 : r1, *im = which ? i2 : i1;
    const uint32_t nre = which ? nr2 : nr1, nim = which ? ni2 : ni1;
    T *pr = planes + size_t(which) * N, *pi = planes + size_t(2 + which) * N;      // layout: [re0 | re1 | im0 | im1]
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x)
    {
        pr[j] = fetch_padded<T>(re, nre, size, fs, off, j);
        pi[j] = fetch_padded<T>(im, nim, size, fs, off, j);
    }
}

// operand 0 <- scale * op(operand 0, operand 1) over all N bins (complex_operation, SpectralFunctions.hpp:49-61)
template <class T>
__global__ void k_cspec_mul(const SpecGeom g, T *__restrict__ planes)
{
    const uint32_t N = 1u << g.log2n;
    const T scale = T(1) / T(size_t(1) << g.log2n);                               // SpectralProcessor.hpp:572
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x)
    {
        const Cx<T> a = cx<T>(planes[k], planes[size_t(2) * N + k]), b = cx<T>(planes[size_t(N) + k], planes[size_t(3) * N + k]);
        const Cx<T> r = spec_op<T>(g.op, scale, a, b);
        planes[k] = r.x;
        planes[size_t(2) * N + k] = r.y;
    }
}

template <class T>
__global__ void k_cspec_arrange(const SpecGeom g, const T *__restrict__ planes, T *__restrict__ out)
{
    const uint32_t N = 1u << g.log2n;
    const T *src = planes + size_t(blockIdx.y) * 2 * N;                            // plane 0 = re, plane 1 = im of operand 0
    T *dst = out + size_t(blockIdx.y) * g.arr.result;
    auto lin = [&](uint32_t i) -> T { return src[i]; };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.arr.result; i += gridDim.x * blockDim.x) dst[i] = arrange_eval(g.arr, i, lin);
}

// ---- change_phase: per-bin stages of ir_phase (SpectralFunctions.hpp:283-336, 405-418) on the packed half spectrum ----
// planes re / im of fft/2 bins; bin 0 carries DC in re[0] and Nyquist in im[0], both handled as real values with a zero
// imaginary part (real_operation, :86-108).  stage: 0 log of the power spectrum (:177-185), 1 causal window of the real
// cepstrum (:307-327), 2 exponential (:187-195), 3 interpolated phase (:209-230), 4 linear-phase amplitude (:158-165).
struct PhaseArgs { uint32_t half; int stage; double scale, min_factor, lin_factor; };

template <class T>
__global__ void k_phase_stage(const PhaseArgs a, T *__restrict__ re, T *__restrict__ im)
{
    const uint32_t fft = a.half * 2;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.half; i += gridDim.x * blockDim.x)
    {
        const T r = re[i], m = im[i];
        T ro, mo;
        switch (a.stage)
        {
            case 0:
            {
                const T min_power = (T) 1e-30;                                     // pow(10, -300 / 10)
                if (i == 0) { ro = T(0.5) * log(max(r * r, min_power)); mo = T(0.5) * log(max(m * m, min_power)); }
                else { ro = T(0.5) * log(max(r * r + m * m, min_power)); mo = T(0); }
                break;
            }
            case 1:
            {
                // sequential semantics of :312-327 (for fft = 2 the first and the third rule both hit bin 0)
                const uint32_t q = fft >> 2;
                ro = r; mo = m;
                if (i == 0) { ro = ro * T(0.5 * a.scale); mo = mo * T(a.scale); }
                if (i >= 1 && i < q) { ro = ro * T(a.scale); mo = mo * T(a.scale); }
                if (i == q) { ro = ro * T(0.5 * a.scale); mo = T(0); }
                if (i > q) { ro = T(0); mo = T(0); }
                break;
            }
            case 2:
                if (i == 0) { ro = exp(r); mo = exp(m); }
                else { const T e = exp(r); T sn, cs; sincos(m, &sn, &cs); ro = e * cs; mo = e * sn; }
                break;
            case 3:
                if (i == 0)
                {
                    ro = T(exp(double(r)));                                        // phase = lin * 0 + min * 0
                    mo = T(exp(double(m)) * cos(a.lin_factor * double(fft >> 1)));
                }
                else
                {
                    const double amp = exp(double(r)), ph = a.lin_factor * double(i) + a.min_factor * double(m);
                    ro = T(amp * cos(ph)); mo = T(amp * sin(ph));
                }
                break;
            default:
                if (i == 0) { ro = sqrt(r * r); mo = sqrt(m * m) * (((fft >> 1) & 1u) ? T(-1) : T(1)); }
                else { ro = sqrt(r * r + m * m) * ((i & 1u) ? T(-1) : T(1)); mo = T(0); }
                break;
        }
        re[i] = ro; im[i] = mo;
    }
}

template <class T>
__global__ void k_scale(T *__restrict__ y, size_t n, T scale)
{
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) y[i] *= scale;
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
    DevBuf d_in3, d_in4, d_planes;      // complex inputs: imaginary planes and the [re0 | re1 | im0 | im1] work array
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

// The reference's arrange_convolve (SpectralProcessor.hpp:445-481) / arrange_correlate (:483-538) as terms.  `copy(out,
// spec, o, src, size)` is a term as is; `wrap(out, spec, o, third, size)` reads from `third - size` in the real
// overloads (:437-443, an end) and from `third` in the Split overloads (:421-427, an offset): complex_index selects.
Arrange build_arrange(uintptr_t n1, uintptr_t n2, uintptr_t fft, uintptr_t fft_actual, int mode, int op, bool complex_index)
{
    Arrange a;
    memset(&a, 0, sizeof(a));
    a.fft_ref = (uint32_t) fft;
    a.shift = (uint32_t) (fft_actual - fft);
    const uintptr_t mn = std::min(n1, n2), mx = std::max(n1, n2), linear = n1 + n2 - 1, min_m1 = mn - 1, size2_m1 = n2 - 1;
    // from_end: `src` was derived from the transform size (the negative lags of a correlation sit at its end)
    auto copy = [&](uintptr_t o, uintptr_t src, uintptr_t size, bool from_end)
    {
        if (!size) return;
        a.lo[a.n] = (uint32_t) o; a.hi[a.n] = (uint32_t) (o + size); a.src[a.n] = (uint32_t) src; a.end_rel[a.n] = from_end ? 1u : 0u; a.n++;
    };
    auto wrap = [&](uintptr_t o, uintptr_t third, uintptr_t size, bool from_end) { copy(o, complex_index ? third : third - size, size, from_end); };
    a.result = (uint32_t) (mode == MODE_LINEAR ? linear : mx);
    if (op == 0)
    {
        switch (mode)
        {
            case MODE_LINEAR: copy(0, 0, linear, false); break;
            case MODE_WRAP: copy(0, 0, mx, false); wrap(0, linear, min_m1, false); break;
            case MODE_WRAP_CENTRE:
            {
                const uintptr_t wrapped = min_m1 >> 1;
                copy(0, wrapped, mx, false);
                wrap(0, linear, min_m1 - wrapped, false);
                wrap(mx - wrapped, wrapped, wrapped, false);
                break;
            }
            default: copy(0, min_m1, mx, false); break;
        }
        return a;
    }
    switch (mode)
    {
        case MODE_LINEAR: copy(0, 0, n1, false); copy(n1, fft - size2_m1, size2_m1, true); break;
        case MODE_WRAP: copy(0, 0, n1, false); wrap(mx - size2_m1, fft, size2_m1, true); break;       // zero(n1, n2) = the gap
        case MODE_WRAP_CENTRE:
        {
            const uintptr_t wrapped1 = min_m1 >> 1, wrapped2 = std::min(size2_m1, mx - wrapped1), wrapped3 = size2_m1 - wrapped2;
            const uintptr_t offset = wrapped3 ? 0 : mx - (size2_m1 + wrapped1);
            copy(0, wrapped1, n1 - wrapped1, false);
            copy(mx - wrapped1, 0, wrapped1, false);
            wrap(offset, fft, wrapped2, true);
            wrap(mx - wrapped3, fft - wrapped2, wrapped3, true);
            break;
        }
        default:
            if (n1 >= n2) copy(0, 0, mx, false);
            else { copy(0, 0, 1, false); copy(1, fft - (mx - 1), mx - 1, true); }
            break;
    }
    return a;
}

template <class T>
int binary_real(hb_spectral *s, T *output, const T *in1, uintptr_t n1, const T *in2, uintptr_t n2, int mode, int op, uintptr_t *written)
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
    g.op = op;
    g.arr = build_arrange(n1, n2, uintptr_t(1) << z.log2n, uintptr_t(1) << log2n, mode, op, false);
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
// complex inputs (SpectralProcessor.hpp:559-614): every plane has its own length, a length of 0 = absent
template <class T>
int binary_complex(hb_spectral *s, T *r_out, T *i_out, const T *r1, uintptr_t nr1, const T *i1, uintptr_t ni1,
                   const T *r2, uintptr_t nr2, const T *i2, uintptr_t ni2, int mode, int op, uintptr_t *written)
{
    const uintptr_t n1 = std::max(nr1, ni1), n2 = std::max(nr2, ni2);
    Sizes z;
    if (written) *written = 0;
    if (!plan_sizes(n1, n2, mode, s->max_log2, z)) return HB_OK;
    if (n1 == 1 && n2 == 1)                                                     // :585-590 (the plain product for both operations)
    {
        const T a = nr1 ? r1[0] : T(0), b = ni1 ? i1[0] : T(0), c = nr2 ? r2[0] : T(0), d = ni2 ? i2[0] : T(0);
        r_out[0] = a * c - b * d;
        i_out[0] = a * d + b * c;
        if (written) *written = 1;
        return HB_OK;
    }
    const uint32_t log2n = z.log2n < 1 ? 1 : z.log2n;
    const size_t N = size_t(1) << log2n;
    int rc;
    if ((rc = s->d_in1.ensure(std::max<size_t>(nr1, 1) * sizeof(T))) || (rc = s->d_in3.ensure(std::max<size_t>(ni1, 1) * sizeof(T))) ||
        (rc = s->d_in2.ensure(std::max<size_t>(nr2, 1) * sizeof(T))) || (rc = s->d_in4.ensure(std::max<size_t>(ni2, 1) * sizeof(T))) ||
        (rc = s->d_planes.ensure(4 * N * sizeof(T))) || (rc = s->d_out.ensure(2 * z.result * sizeof(T)))) return rc;
    if (nr1) HB_CUDA(cudaMemcpyAsync(s->d_in1.p, r1, nr1 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    if (ni1) HB_CUDA(cudaMemcpyAsync(s->d_in3.p, i1, ni1 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    if (nr2) HB_CUDA(cudaMemcpyAsync(s->d_in2.p, r2, nr2 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    if (ni2) HB_CUDA(cudaMemcpyAsync(s->d_in4.p, i2, ni2 * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    SpecGeom g;
    g.log2n = log2n; g.n1 = (uint32_t) n1; g.n2 = (uint32_t) n2; g.mn = (uint32_t) z.mn; g.mx = (uint32_t) z.mx;
    g.fold_size = (uint32_t) z.fold_size; g.mode = mode; g.first_is_big = n1 >= n2;
    g.op = op;
    g.arr = build_arrange(n1, n2, uintptr_t(1) << z.log2n, uintptr_t(1) << log2n, mode, op, true);
    T *planes = (T *) s->d_planes.p;
    const unsigned blocks = (unsigned) std::min<size_t>((N + 255) / 256, 1024);
    k_cspec_pack<T><<<dim3(blocks, 2), 256, 0, s->stream>>>(g, (const T *) s->d_in1.p, (uint32_t) nr1, (const T *) s->d_in3.p, (uint32_t) ni1,
                                                            (const T *) s->d_in2.p, (uint32_t) nr2, (const T *) s->d_in4.p, (uint32_t) ni2, planes);
    HB_LAUNCH_CHECK();
    // both operands forward in one batched launch (planes N apart), product into operand 0, unscaled inverse of operand 0
    if ((rc = cfft_planes(s->dtype, planes, planes + 2 * N, planes, planes + 2 * N, (int) log2n, 0, 2, N, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
    k_cspec_mul<T><<<blocks, 256, 0, s->stream>>>(g, planes);
    HB_LAUNCH_CHECK();
    if ((rc = cfft_planes(s->dtype, planes, planes + 2 * N, planes, planes + 2 * N, (int) log2n, 1, 1, N, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
    k_cspec_arrange<T><<<dim3((unsigned) std::min<size_t>((z.result + 255) / 256, 1024), 2), 256, 0, s->stream>>>(g, planes, (T *) s->d_out.p);
    HB_LAUNCH_CHECK();
    HB_CUDA(cudaMemcpyAsync(r_out, s->d_out.p, z.result * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaMemcpyAsync(i_out, (const T *) s->d_out.p + z.result, z.result * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    if (written) *written = z.result;
    return HB_OK;
}

// change_phase (SpectralProcessor.hpp:186-208): output receives fft_size samples
template <class T>
int change_phase(hb_spectral *s, T *output, const T *input, uintptr_t size, double phase, double time_multiplier, uintptr_t *written)
{
    if (written) *written = 0;
    if (!size) return HB_OK;
    if (size == 1) { output[0] = input[0]; if (written) *written = 1; return HB_OK; }
    // calc_fft_size_log2(round(size * time_multiplier)) (:189, 229-240)
    const uintptr_t target = (uintptr_t) std::llround(double(size) * time_multiplier);
    uint32_t log2n = 0;
    while (target >> log2n) log2n++;
    if (log2n && target == (uintptr_t(1) << (log2n - 1))) log2n--;
    if (log2n < 1 || log2n > s->max_log2)
    {
        set_error("hb_spectral_change_phase: FFT size 2^%u outside 2 .. the processor's maximum 2^%u", log2n, s->max_log2);
        return HB_ERR_BAD_ARG;
    }
    const size_t n = size_t(1) << log2n, half = n >> 1, len = std::min<size_t>(size, n);
    int rc;
    if ((rc = s->d_in1.ensure(len * sizeof(T))) || (rc = s->d_planes.ensure(2 * half * sizeof(T))) || (rc = s->d_out.ensure(n * sizeof(T)))) return rc;
    HB_CUDA(cudaMemcpyAsync(s->d_in1.p, input, len * sizeof(T), cudaMemcpyHostToDevice, s->stream));
    T *re = (T *) s->d_planes.p, *im = re + half;
    if ((rc = rfft_planes(s->dtype, s->d_in1.p, len, re, im, (int) log2n, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
    PhaseArgs a;
    a.half = (uint32_t) half; a.scale = 1.0 / double(n); a.min_factor = 0; a.lin_factor = 0;
    const unsigned blocks = (unsigned) std::min<size_t>((half + 255) / 256, 1024);
    auto stage = [&](int st) -> int
    {
        a.stage = st;
        k_phase_stage<T><<<blocks, 256, 0, s->stream>>>(a, re, im);
        HB_LAUNCH_CHECK();
        return HB_OK;
    };
    if (phase == 0.5)
    {
        if ((rc = stage(4))) return rc;                                          // ir_phase :407-413, zero_center = false
    }
    else
    {
        // minimum_phase_components (:283-336): log power -> real cepstrum -> causal window -> back to the spectrum
        if ((rc = stage(0))) return rc;
        if ((rc = rifft_planes(s->dtype, re, im, nullptr, (int) log2n, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
        if ((rc = stage(1))) return rc;
        if ((rc = rfft_planes(s->dtype, nullptr, 0, re, im, (int) log2n, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
        if (phase == 0.0)
        {
            if ((rc = stage(2))) return rc;
        }
        else
        {
            // phase_interpolate (:199-230), zero_center = false
            const double delay_factor = (phase <= 0.5) ? 0.0 : 1.0 / double(n);
            const double ph = std::max(0.0, std::min(1.0, phase));
            a.min_factor = 1.0 - 2.0 * ph;
            a.lin_factor = -2.0 * M_PI * (ph - delay_factor);
            if ((rc = stage(3))) return rc;
        }
    }
    if ((rc = rifft_planes(s->dtype, re, im, s->d_out.p, (int) log2n, s->tw, s->tw_log2, s->stream, &s->big))) return rc;
    k_scale<T><<<(unsigned) std::min<size_t>((n + 255) / 256, 1024), 256, 0, s->stream>>>((T *) s->d_out.p, n, T(0.5) / T(n));
    HB_LAUNCH_CHECK();
    HB_CUDA(cudaMemcpyAsync(output, s->d_out.p, n * sizeof(T), cudaMemcpyDeviceToHost, s->stream));
    HB_CUDA(cudaStreamSynchronize(s->stream));
    if (written) *written = n;
    return HB_OK;
}

int real_entry(hb_spectral *s, void *output, const void *in1, uintptr_t n1, const void *in2, uintptr_t n2, int mode, int op, uintptr_t *written, const char *who)
{
    if (!s || mode < MODE_LINEAR || mode > MODE_FOLD_REPEAT || ((!output || !in1 || !in2) && n1 && n2))
    {
        set_error("%s: bad argument", who);
        return HB_ERR_BAD_ARG;
    }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? binary_real<double>(s, (double *) output, (const double *) in1, n1, (const double *) in2, n2, mode, op, written)
                              : binary_real<float>(s, (float *) output, (const float *) in1, n1, (const float *) in2, n2, mode, op, written);
}

int complex_entry(hb_spectral *s, void *r_out, void *i_out, const void *r1, uintptr_t nr1, const void *i1, uintptr_t ni1,
                  const void *r2, uintptr_t nr2, const void *i2, uintptr_t ni2, int mode, int op, uintptr_t *written, const char *who)
{
    const bool some = (nr1 || ni1) && (nr2 || ni2);
    if (!s || mode < MODE_LINEAR || mode > MODE_FOLD_REPEAT || (some && (!r_out || !i_out)) || (nr1 && !r1) || (ni1 && !i1) || (nr2 && !r2) || (ni2 && !i2))
    {
        set_error("%s: bad argument", who);
        return HB_ERR_BAD_ARG;
    }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? binary_complex<double>(s, (double *) r_out, (double *) i_out, (const double *) r1, nr1, (const double *) i1, ni1,
                                                       (const double *) r2, nr2, (const double *) i2, ni2, mode, op, written)
                              : binary_complex<float>(s, (float *) r_out, (float *) i_out, (const float *) r1, nr1, (const float *) i1, ni1,
                                                      (const float *) r2, nr2, (const float *) i2, ni2, mode, op, written);
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
    s->d_in3.release(); s->d_in4.release(); s->d_planes.release();
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
    return real_entry(s, output, in1, n1, in2, n2, mode, 0, written, "hb_spectral_convolve");
}

extern "C" int hb_spectral_correlate(hb_spectral *s, void *output, const void *in1, uintptr_t n1, const void *in2, uintptr_t n2, int mode, uintptr_t *written)
{
    return real_entry(s, output, in1, n1, in2, n2, mode, 1, written, "hb_spectral_correlate");
}

extern "C" int hb_spectral_convolve_complex(hb_spectral *s, void *r_out, void *i_out, const void *r_in1, uintptr_t nr1, const void *i_in1, uintptr_t ni1,
                                            const void *r_in2, uintptr_t nr2, const void *i_in2, uintptr_t ni2, int mode, uintptr_t *written)
{
    return complex_entry(s, r_out, i_out, r_in1, nr1, i_in1, ni1, r_in2, nr2, i_in2, ni2, mode, 0, written, "hb_spectral_convolve_complex");
}

extern "C" int hb_spectral_correlate_complex(hb_spectral *s, void *r_out, void *i_out, const void *r_in1, uintptr_t nr1, const void *i_in1, uintptr_t ni1,
                                             const void *r_in2, uintptr_t nr2, const void *i_in2, uintptr_t ni2, int mode, uintptr_t *written)
{
    return complex_entry(s, r_out, i_out, r_in1, nr1, i_in1, ni1, r_in2, nr2, i_in2, ni2, mode, 1, written, "hb_spectral_correlate_complex");
}

extern "C" int hb_spectral_change_phase(hb_spectral *s, void *output, const void *input, uintptr_t size, double phase, double time_multiplier, uintptr_t *written)
{
    if (!s || (size && (!output || !input))) { set_error("hb_spectral_change_phase: bad argument"); return HB_ERR_BAD_ARG; }
    std::lock_guard<std::mutex> g(s->lock);
    int rc = use_device(s->device);
    if (rc) return rc;
    return s->dtype == HB_F64 ? change_phase<double>(s, (double *) output, (const double *) input, size, phase, time_multiplier, written)
                              : change_phase<float>(s, (float *) output, (const float *) input, size, phase, time_multiplier, written);
}
