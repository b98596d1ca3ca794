// tools/hop_rate.cu -- back-to-back single-block device calls from a C++ caller (no Python in the loop): microseconds per
// hb_conv_process_dev call on the fused engines of BASELINE configs 1-3, by hop-overlap mode (hb_conv_set_hop_overlap).
//   nvcc -O2 -o tools/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "hisstools_b200.h"

static double run(int ins, size_t taps, size_t B, int mode, bool own, int calls)
{
    hb_conv *c = nullptr;
    if (hb_conv_create(&c, HB_F32, 1, ins, 1, 2 * B, taps, 0, 0, 0) < 0) { printf("create failed: %s\n", hb_last_error()); exit(1); }
    hb_conv_set_reset_offset(c, 0);
    hb_conv_set_hop_overlap(c, mode);
    std::vector<float> ir(taps);
    for (int i = 0; i < ins; i++)
    {
        for (size_t k = 0; k < taps; k++) ir[k] = (float) (std::exp(-6.9 * k / taps) * ((rand() & 1023) / 512.0 - 1.0));
        if (hb_conv_set_ir(c, 0, i, 0, ir.data(), HB_F32, taps)) { printf("set_ir failed\n"); exit(1); }
    }
    const int pool = 4;
    float *x = nullptr, *y = nullptr;
    cudaMalloc(&x, sizeof(float) * pool * ins * B);
    cudaMalloc(&y, sizeof(float) * pool * B);
    cudaMemset(x, 0, sizeof(float) * pool * ins * B);
    cudaStream_t st = nullptr;
    if (!own) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++)
    {
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < calls; k++)
        {
            const int q = k % pool;
            const int rc = hb_conv_process_dev(c, x + size_t(q) * ins * B, B, y + size_t(q) * B, B, B, 0, st);
            if (rc) { printf("process failed %d: %s\n", rc, hb_last_error()); exit(1); }
        }
        cudaDeviceSynchronize();
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / calls;
        if (rep && us < best) best = us;
    }
    if (st) cudaStreamDestroy(st);
    cudaFree(x); cudaFree(y);
    hb_conv_destroy(c);
    return best;
}

// the same engine through hb_conv_process: host rows in, host rows out (pipelined behind the one-hop latency of the API)
static double run_host(int ins, size_t taps, size_t B, int calls)
{
    hb_conv *c = nullptr;
    if (hb_conv_create(&c, HB_F32, 1, ins, 1, 2 * B, taps, 0, 0, 0) < 0) { printf("create failed: %s\n", hb_last_error()); exit(1); }
    hb_conv_set_reset_offset(c, 0);
    std::vector<float> ir(taps);
    for (int i = 0; i < ins; i++)
    {
        for (size_t k = 0; k < taps; k++) ir[k] = (float) (std::exp(-6.9 * k / taps) * ((rand() & 1023) / 512.0 - 1.0));
        if (hb_conv_set_ir(c, 0, i, 0, ir.data(), HB_F32, taps)) { printf("set_ir failed\n"); exit(1); }
    }
    std::vector<std::vector<float>> x(ins, std::vector<float>(B, 0.25f));
    std::vector<float> y(B);
    std::vector<const void *> ip(ins);
    for (int i = 0; i < ins; i++) ip[i] = x[i].data();
    void *op[1] = {y.data()};
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++)
    {
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < calls; k++)
        {
            const int rc = hb_conv_process(c, ip.data(), op, B, 0);
            if (rc) { printf("process failed %d: %s\n", rc, hb_last_error()); exit(1); }
        }
        cudaDeviceSynchronize();
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / calls;
        if (rep && us < best) best = us;
    }
    hb_conv_destroy(c);
    return best;
}

int main()
{
    struct { const char *name; int ins; size_t taps, B; } cfg[] = {{"c1", 1, 4096, 512}, {"c2", 1, 65536, 1024}, {"c3", 8, 131072, 2048}};
    printf("# microseconds per hb_conv_process_dev call of one block (C++ caller, best of 3 runs of 20000 calls, wall clock incl. final sync)\n");
    printf("# config   mode 0 (strict)   mode 2 (caller stream)   mode 1 (engine's own stream)   M samples/s at the last   |   hb_conv_process (host rows): us per call, M samples/s\n");
    for (auto &f : cfg)
    {
        const double a = run(f.ins, f.taps, f.B, 0, false, 20000), b = run(f.ins, f.taps, f.B, 2, false, 20000), d = run(f.ins, f.taps, f.B, 1, true, 20000);
        const double h = run_host(f.ins, f.taps, f.B, 20000);
        printf("%s   %8.2f   %8.2f   %8.2f   %8.1f   |   %8.2f   %8.1f\n", f.name, a, b, d, f.B / d, h, f.B / h);
    }
    return 0;
}
