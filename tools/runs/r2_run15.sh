#!/bin/bash
# round 2: compute-sanitizer on the kernels added this round (fused hop with overlapping hops, packed multi-hop kernel, hop batches, per-pair restart)
mkdir -p gpurun_out
{
echo "# compute-sanitizer on a B200 (gpurun), round 2"
echo "compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -m gpu -x -q -k 'not bench_parity'"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -m gpu -x -q -k 'not bench_parity' 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | tail -6
echo "compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k 'overlap_between_calls and float32 and (512 or 2048 or 128) or matrix_device_calls or multi_hop_reuse and float32-None or hop_batches'"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k 'overlap_between_calls and float32 and (512 or 2048 or 128) or matrix_device_calls or multi_hop_reuse and float32-None or hop_batches' 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" | tail -6
echo "compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k 'overlap_between_calls and float32 or multi_hop_reuse and float32-None'"
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_conv.py -m gpu -x -q -k 'overlap_between_calls and float32 or multi_hop_reuse and float32-None' 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error" | tail -6
} > gpurun_out/r2ac_sanitizer.txt 2>&1
cat gpurun_out/r2ac_sanitizer.txt
