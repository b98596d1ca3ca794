#!/bin/bash
mkdir -p gpurun_out
{
echo "# compute-sanitizer on a B200 (gpurun), round 2, kernels added this round"
for k in 'multi_hop_reuse and float32-None' 'overlap_between_calls and 1-1-1-512-4096-float32' 'overlap_between_calls and 3-1-2-256-5000-float32' 'overlap_between_calls and 20-1-1-128-1024-float64'; do
echo "compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k '$k'"
timeout 200 compute-sanitizer --tool synccheck --print-limit 3 --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "$k" 2>&1 | grep -E "========= (Barrier|Error|ERROR)|     at |passed|failed" | head -8
done
for k in 'overlap_between_calls and 1-1-1-64-64-float32' 'overlap_between_calls and 1-1-1-128-256-float64' 'overlap_between_calls and 2-3-1-512-3000-float32'; do
echo "compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k '$k'"
timeout 240 compute-sanitizer --tool memcheck --print-limit 3 --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k "$k" 2>&1 | grep -E "========= (Invalid|Error|ERROR)|     at |passed|failed" | head -8
done
} > gpurun_out/r2ad_sanitizer.txt 2>&1
cat gpurun_out/r2ad_sanitizer.txt | cut -c1-200
