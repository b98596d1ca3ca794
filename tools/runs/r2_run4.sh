#!/bin/bash
# round 2: one tail stream against two alternating tail streams (hb_conv_set_tail_streams) on the launch-overhead-sensitive shapes
mkdir -p gpurun_out
for wl in c4r8 c5 c4; do
  for ts in 1 2; do
    python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-multi-hop --no-parity --tail-streams $ts > gpurun_out/r2i_${wl}_ts$ts.json 2> gpurun_out/r2i_${wl}_ts$ts.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_${wl}_ts$ts.json').read().strip().splitlines()[-1])
r=d['roofline']
print('$wl tail_streams=$ts value %.1f ms/block %.4f e2e %.1f kernel_ms %.4f achieved %.0f hop_frac %.3f clk %s' % (d['value'], d['timing']['ms_per_block'], d['e2e']['value'], r['kernel_ms'], r['achieved'], r['hop_frac'], d['clocks']['sm_mhz']))
PY
    tail -2 gpurun_out/r2i_${wl}_ts$ts.err
  done
done
python -m pytest tests/test_gpu_conv.py -x -q -k "schedules_agree or multi_hop or fft_paths" 2>&1 | tail -3
HB_TAIL_STREAMS=2 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fullsize.py -x -q -k "schedules_agree or multi_hop or fft_paths or config4 or config5 or golden" 2>&1 | tail -3
