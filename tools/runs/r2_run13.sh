#!/bin/bash
# round 2: latency-mode matrices end to end, call rates after the single-CTA reduction change, whole suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_pytest.log 2>&1; tail -5 gpurun_out/r2aa_pytest.log
nvcc -O2 -o /tmp/hop_rate tools/hop_rate.cu -Iinclude -Lhisstools_library_b200/lib -lhisstools_b200 -Xlinker -rpath=$PWD/hisstools_library_b200/lib -Wno-deprecated-gpu-targets \
  && timeout 300 /tmp/hop_rate > gpurun_out/r2aa_hop_rate.txt 2>&1
cat gpurun_out/r2aa_hop_rate.txt
timeout 900 python tools/latency_mode_probe.py > gpurun_out/r2aa_latency_modes.txt 2>&1
cat gpurun_out/r2aa_latency_modes.txt
