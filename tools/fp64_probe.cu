// FP64 issue rate of one SM and of the chip: independent DFMA chains, 8 per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int iters, double a, double b)
{
    double v[8];
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fma(v[i], a, b);
    double s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    const int grids[] = {1, 148, 148 * 2};
    const int threads[] = {128, 256, 512, 1024};
    for (int g : grids)
        for (int t : threads)
        {
            k<<<g, t>>>(out, 16, 1.0000001, 1e-9);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k<<<g, t>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double fma = double(g) * t * 8.0 * iters;
            printf("grid %4d x %4d threads: %8.3f ms  %8.2f GDFMA/s  (%.2f DFMA/clk/SM at 1.965 GHz over %d SMs)\n", g, t, ms, fma / ms * 1e-6,
                   fma / (ms * 1e-3) / 1.965e9 / (g < 148 ? g : 148), g < 148 ? g : 148);
        }
    return 0;
}
