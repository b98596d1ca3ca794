// hb_conv_mh.h -- host interface of the packed-FFMA2 multi-hop kernel (hb_conv_mh.cu)
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace hb
{
struct Geom;
struct Range;

// can a launch of `nh` (2, 4, 8) hops per pass run on the packed kernel for this tiling? (float engines only)
bool mh2_supported(const Geom &g, int nh);
// nh hops over one pass of the IR spectra of range r; S holds nh partial-segment sets set_stride vectors apart.
// nh == 8 works on half units: segments of Q / 2 vectors, work items r.U * 2 (SegSet::split = 2)
int launch_cmac_mh2(const Geom &g, const Range &r, const void *H, const void *X, void *S, int nh, uint64_t set_stride, cudaStream_t st);
size_t mh2_stage_bytes(uint32_t tbv, int nh, int rb);
int mh2_stages(uint32_t tbv, int nh, int rb);
} // namespace hb
