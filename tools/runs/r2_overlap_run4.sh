#!/bin/bash
mkdir -p gpurun_out
{
for wl in c1 c2 c3; do for m in 0 2; do timeout 120 python tools/fused_chain_trace.py $wl $m; done; done
} > gpurun_out/r2w_chain_trace.txt 2>&1
cat gpurun_out/r2w_chain_trace.txt
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "bench_parity" 2>&1 | tail -30
