#!/bin/bash
# round 2: overlapping fused hops -- cluster-size sweep of the call rate (C++ caller)
mkdir -p gpurun_out
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets || exit 1
{
echo "## default"; timeout 300 /tmp/hop_rate
for cs in 1 2 4 8; do
  echo "## HB_FUSED_MAX_CS=$cs HB_FUSED_MAX_KB=20000"; HB_FUSED_MAX_CS=$cs HB_FUSED_MAX_KB=20000 timeout 300 /tmp/hop_rate
done
} > gpurun_out/r2v_hop_rate_sweep.txt 2>&1
cat gpurun_out/r2v_hop_rate_sweep.txt
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused or overlap or bench_parity" 2>&1 | tail -3
