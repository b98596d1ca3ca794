// coreside_probe.cu -- does a second kernel's CTA get placed on an SM beside a resident CTA of a persistent kernel?
// A: 148 CTAs x 256 threads, spins `dur_us`, dynamic smem smemA, NA live accumulators (register pressure)
// B: 64 CTAs x 512 threads, spins 20 us, dynamic smem smemB, NB accumulators
// Reports B's duration (events on its stream) launched while A is resident.  Build: nvcc -arch=sm_100a -O3 -o probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int N, int THREADS>
__global__ void __launch_bounds__(THREADS) spin(float *out, unsigned long long dur_ns, float seed)
{
    extern __shared__ float sm[];
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; i++) acc[i] = seed * (i + 1) + threadIdx.x;
    sm[threadIdx.x] = seed;
    __syncthreads();
    const unsigned long long t0 = gtime();
    while (gtime() - t0 < dur_ns)
    {
#pragma unroll
        for (int i = 0; i < N; i++) acc[i] = fmaf(acc[i], 1.0001f, sm[(threadIdx.x + i) & 255]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < N; i++) s += acc[i];
    if (s == 12345.678f) out[blockIdx.x] = s;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NA, int NB>
void trial(size_t smemA, size_t smemB, int carve, int gridB)
{
    if (smemA < 2048) smemA = 2048;
    if (smemB < 2048) smemB = 2048;
    auto A = spin<NA, 256>;
    auto B = spin<NB, 512>;
    cudaFuncAttributes fa, fb;
    CK(cudaFuncGetAttributes(&fa, A));
    CK(cudaFuncGetAttributes(&fb, B));
    CK(cudaFuncSetAttribute(A, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemA));
    CK(cudaFuncSetAttribute(B, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemB));
    CK(cudaFuncSetAttribute(A, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    CK(cudaFuncSetAttribute(B, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    float *out;
    CK(cudaMalloc(&out, 4096));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float alone = 0, beside = 0;
    for (int rep = 0; rep < 2; rep++)
    {
        CK(cudaEventRecord(e0, s2));
        B<<<gridB, 512, smemB, s2>>>(out, 20000ull, 1.f);
        CK(cudaEventRecord(e1, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&alone, e0, e1));
        A<<<148, 256, smemA, s1>>>(out, 1000000ull, 1.f);
        // give A time to become resident
        B<<<1, 512, smemB, s2>>>(out, 50000ull, 1.f);
        CK(cudaEventRecord(e0, s2));
        B<<<gridB, 512, smemB, s2>>>(out, 20000ull, 1.f);
        CK(cudaEventRecord(e1, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&beside, e0, e1));
    }
    printf("A regs %3d smem %6zu | B regs %3d smem %6zu grid %3d | carveout %3d | B alone %.3f ms, beside A %.3f ms  %s\n", fa.numRegs, smemA,
           fb.numRegs, smemB, gridB, carve, alone, beside, beside < 0.5f ? "CO-RESIDENT" : "serialised");
    cudaFree(out);
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

// a persistent bulk-copy (TMA) streaming kernel: one thread keeps a ring of `stages` x 32 KB copies in flight,
// everybody waits on the mbarrier of the oldest stage and reads one value of it
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256, 1) tma_stream(const float4 *src, size_t chunks_per_cta, int stages, float *out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    float4 *ring = reinterpret_cast<float4 *>(raw);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(raw + size_t(stages) * 32768);
    const unsigned CH = 32768;
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const float4 *base = src + size_t(blockIdx.x) * chunks_per_cta * (CH / 16);
    size_t issued = 0;
    auto issue = [&](int s)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(CH) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(ring + size_t(s) * (CH / 16))), "l"(base + issued * (CH / 16)), "r"(CH), "r"(smem_u32(&bar[s])) : "memory");
        issued++;
    };
    if (threadIdx.x == 0)
        for (int s = 0; s + 1 < stages && issued < chunks_per_cta; s++) issue(s);
    float acc = 0.f;
    int stage = 0;
    unsigned parity = 0;
    for (size_t k = 0; k < chunks_per_cta; k++)
    {
        if (threadIdx.x == 0 && issued < chunks_per_cta) issue(stage ? stage - 1 : stages - 1);
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[stage])), "r"(parity) : "memory");
        acc += ring[size_t(stage) * (CH / 16) + threadIdx.x].x;
        if (++stage == stages) { stage = 0; parity ^= 1; }
        __syncthreads();
    }
    if (acc == 12345.678f) out[blockIdx.x] = acc;
}

void tma_trial(int stages, size_t smemB, int gridB)
{
    auto B = spin<28, 512>;
    const size_t smemA = size_t(stages) * 32768 + 64;
    CK(cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemA));
    CK(cudaFuncSetAttribute(B, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smemB));
    const size_t chunks = 1800;                      // 148 x 1800 x 32 KB = 8.7 GB ~ 1.2 ms
    float4 *src;
    float *out;
    CK(cudaMalloc(&src, size_t(148) * chunks * 32768));
    CK(cudaMemset(src, 0, size_t(148) * chunks * 32768));
    CK(cudaMalloc(&out, 4096));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, a0, a1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
    float alone = 0, beside = 0, ta = 0;
    for (int rep = 0; rep < 2; rep++)
    {
        CK(cudaEventRecord(e0, s2));
        B<<<gridB, 512, smemB, s2>>>(out, 20000ull, 1.f);
        CK(cudaEventRecord(e1, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&alone, e0, e1));
        CK(cudaEventRecord(a0, s1));
        tma_stream<<<148, 256, smemA, s1>>>(src, chunks, stages, out);
        CK(cudaEventRecord(a1, s1));
        B<<<1, 512, smemB, s2>>>(out, 50000ull, 1.f);
        CK(cudaEventRecord(e0, s2));
        B<<<gridB, 512, smemB, s2>>>(out, 20000ull, 1.f);
        CK(cudaEventRecord(e1, s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventElapsedTime(&beside, e0, e1));
        CK(cudaEventElapsedTime(&ta, a0, a1));
    }
    printf("TMA ring %d x 32 KB (%.3f ms, %.0f GB/s) | B smem %6zu grid %3d | B alone %.3f ms, beside %.3f ms  %s\n", stages, ta,
           148.0 * chunks * 32768 / (ta * 1e-3) / 1e9, smemB, gridB, alone, beside, beside < 0.5f ? "CO-RESIDENT" : "serialised");
    cudaFree(src); cudaFree(out);
    cudaStreamDestroy(s1); cudaStreamDestroy(s2);
}

int main()
{
    for (int st : {2, 3, 4, 6})
    {
        tma_trial(st, 67 * 1024, 64);
        tma_trial(st, 2048, 64);
    }
    tma_trial(3, 67 * 1024, 148);
    return 0;
}
