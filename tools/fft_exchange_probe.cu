// tools/fft_exchange_probe.cu -- the exchange between two radix-8 Stockham passes done through warp shuffles against the same
// exchange done through shared memory (what block_fft does), at 512 and 4096 complex points per CTA.
//
// north_star names "warp-shuffle butterflies"; DESIGN.md 4 argues for shared memory on instruction counts.  This probe measures it:
// every thread holds 8 complex values, runs the radix-8 butterfly with twiddles (the arithmetic of one pass), and hands them on
//   SMEM : 8 x STS.64 (padded), barrier, 8 x LDS.64 -- any thread can be the partner (the production path);
//   SHFL : an 8 x 8 transpose among groups of 8 lanes in three butterfly stages (24 SHFL.BFLY + selects per thread) -- only
//          partners inside one warp can be reached, i.e. one of the log8(M) exchanges of a transform of M >= 512 points.
// Both variants compute the same values (checked).  Prints ns per pass and CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fft_exchange_probe tools/fft_exchange_probe.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

struct cxf { float x, y; };
__device__ __forceinline__ cxf cadd(cxf a, cxf b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cxf csub(cxf a, cxf b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cxf cmul(cxf a, cxf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cxf mul_mi(cxf a) { return {a.y, -a.x}; }
__device__ __forceinline__ void dft4(cxf &v0, cxf &v1, cxf &v2, cxf &v3)
{
    cxf a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi(csub(v1, v3));
    v0 = cadd(a0, a2); v2 = csub(a0, a2); v1 = cadd(a1, a3); v3 = csub(a1, a3);
}
__device__ __forceinline__ void dft8(cxf *v)
{
    const float h = 0.70710678118654752f;
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    cxf o1 = {(v[3].x + v[3].y) * h, (v[3].y - v[3].x) * h}, o2 = mul_mi(v[5]), o3 = {(v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h};
    cxf e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0); v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2); v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// thread t (lane l = t % 8 inside its group of 8) ends up with element l of the 8 threads of its group: v'[k] = v_of_lane_k[l]
template <bool SHFL>
__global__ void __launch_bounds__(512) k_probe(const cxf *__restrict__ in, cxf *__restrict__ out, const cxf *__restrict__ tw, int passes)
{
    extern __shared__ cxf s[];
    const uint32_t t = threadIdx.x, nthr = blockDim.x, lane8 = t & 7u, grp = t >> 3;
    cxf v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = in[(size_t(blockIdx.x) * nthr + t) * 8 + k];
    for (int p = 0; p < passes; p++)
    {
#pragma unroll
        for (int k = 1; k < 8; k++) v[k] = cmul(v[k], tw[(t * k + p) & 4095]);
        dft8(v);
        if (SHFL)
        {
#pragma unroll
            for (int bit = 1; bit < 8; bit <<= 1)
            {
                const bool up = (lane8 & bit) != 0;
#pragma unroll
                for (int k = 0; k < 8; k++)
                {
                    if (k & bit) continue;
                    // the lane with the bit clear keeps v[k] and gives v[k | bit]; its partner the other way round
                    cxf send = up ? v[k] : v[k | bit];
                    cxf recv;
                    recv.x = __shfl_xor_sync(0xffffffffu, send.x, bit);
                    recv.y = __shfl_xor_sync(0xffffffffu, send.y, bit);
                    if (up) v[k] = recv; else v[k | bit] = recv;
                }
            }
        }
        else
        {
            // padded like block_fft (one slot per 32 elements): element k of thread t at (grp * 8 + k) * 8 + lane8 after the transpose
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint32_t i = (grp * 8 + lane8) * 8 + k; s[i + (i >> 5)] = v[k]; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint32_t i = (grp * 8 + k) * 8 + lane8; v[k] = s[i + (i >> 5)]; }
            __syncthreads();
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[(size_t(blockIdx.x) * nthr + t) * 8 + k] = v[k];
}

int main()
{
    const int passes = 64, ctas = 148 * 4;
    std::vector<cxf> h_tw(4096);
    for (int q = 0; q < 4096; q++) h_tw[q] = {(float) cos(-2 * M_PI * q / 8192.0), (float) sin(-2 * M_PI * q / 8192.0)};
    cxf *d_tw, *d_in, *d_a, *d_b;
    const size_t n = size_t(ctas) * 512 * 8;
    cudaMalloc(&d_tw, 4096 * sizeof(cxf)); cudaMalloc(&d_in, n * sizeof(cxf)); cudaMalloc(&d_a, n * sizeof(cxf)); cudaMalloc(&d_b, n * sizeof(cxf));
    cudaMemcpy(d_tw, h_tw.data(), 4096 * sizeof(cxf), cudaMemcpyHostToDevice);
    std::vector<cxf> h_in(n);
    for (size_t i = 0; i < n; i++) h_in[i] = {float((i * 2654435761u) % 1000) / 1000.f - 0.5f, float((i * 40503u) % 1000) / 1000.f - 0.5f};
    cudaMemcpy(d_in, h_in.data(), n * sizeof(cxf), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {64, 512})
    {
        const size_t smem = size_t(threads) * 8 * sizeof(cxf) * 33 / 32 + 64;
        cudaFuncSetAttribute(k_probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        float ms[2];
        for (int v = 0; v < 2; v++)
        {
            for (int rep = 0; rep < 3; rep++)
            {
                cudaEventRecord(e0);
                if (v) k_probe<true><<<ctas, threads, 0>>>(d_in, d_b, d_tw, passes); else k_probe<false><<<ctas, threads, smem>>>(d_in, d_a, d_tw, passes);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[v], e0, e1);
            }
        }
        std::vector<cxf> a(size_t(ctas) * threads * 8), b(a.size());
        cudaMemcpy(a.data(), d_a, a.size() * sizeof(cxf), cudaMemcpyDeviceToHost); cudaMemcpy(b.data(), d_b, b.size() * sizeof(cxf), cudaMemcpyDeviceToHost);
        double diff = 0, ref = 0;
        for (size_t i = 0; i < a.size(); i++) { if (std::isfinite(a[i].x)) { diff += fabs(a[i].x - b[i].x) + fabs(a[i].y - b[i].y); ref += fabs(a[i].x) + fabs(a[i].y); } }
        // CTAs run 4 per SM slot-wise: time per pass of one CTA = total / passes / (waves of CTAs per SM)
        printf("%4d points per CTA (%3d threads): shared-memory exchange %.1f ns per pass and CTA, shuffle exchange %.1f ns (%d CTAs, %d passes; mismatch %.2g)\n",
               threads * 8, threads, ms[0] * 1e6 / passes / (ctas / 148.0), ms[1] * 1e6 / passes / (ctas / 148.0), ctas, passes, ref > 0 ? diff / ref : 0.0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
