// hb_common.cuh -- error reporting, launch accounting and small RAII helpers shared by the
// translation units of libhisstools_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/hisstools_b200.h"

#define HB_STR2(x) #x
#define HB_STR(x) HB_STR2(x)

namespace hb
{

void set_error(const char *fmt, ...);
void count_launch(uint64_t n = 1);

#define HB_CUDA(call)                                                                                   \
    do                                                                                                  \
    {                                                                                                   \
        cudaError_t hb_e_ = (call);                                                                     \
        if (hb_e_ != cudaSuccess)                                                                       \
        {                                                                                               \
            hb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(hb_e_));     \
            return HB_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

#define HB_LAUNCH_CHECK()                                                                               \
    do                                                                                                  \
    {                                                                                                   \
        hb::count_launch();                                                                             \
        cudaError_t hb_e_ = cudaGetLastError();                                                         \
        if (hb_e_ != cudaSuccess)                                                                       \
        {                                                                                               \
            hb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(hb_e_)); \
            return HB_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

// growable device / pinned-host scratch
struct DevBuf
{
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return HB_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        HB_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
        return HB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PinnedBuf
{
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return HB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        HB_CUDA(cudaMallocHost(&p, bytes));
        cap = bytes;
        return HB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline size_t dtype_size(int dtype) { return dtype == HB_F64 ? 8 : 4; }

// device twiddle table (half circle of order 2^log2) built in long double on the host
int make_twiddles(int dtype, int log2, void **d_out);

// select the device and make sure it is usable; HB_ERR_CUDA otherwise (no CPU fallback exists)
int use_device(int device);

} // namespace hb
