mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r1g_bench_c4.json 2> gpurun_out/r1g_bench_c4.err
for w in c1 c2 c3 c5; do timeout 200 python bench.py --workload $w > gpurun_out/r1g_bench_$w.json 2>/dev/null; done
timeout 300 python bench.py --impl reference > gpurun_out/r1g_bench_ref.json 2>&1
python tools/summarise_bench.py gpurun_out/r1g_bench_c4.json gpurun_out/r1g_bench_c1.json gpurun_out/r1g_bench_c2.json gpurun_out/r1g_bench_c3.json gpurun_out/r1g_bench_c5.json gpurun_out/r1g_bench_ref.json
