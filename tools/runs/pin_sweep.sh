run() { # wl pin_p pin_s
  local tag="$1_p$2_s$3"
  local envs=""
  [ "$2" != "-" ] && envs="$envs HB_PIN_P=$2"
  [ "$3" != "-" ] && envs="$envs HB_PIN_S=$3"
  env $envs python bench.py --workload $1 --no-cpu --no-multi-hop --steps 40 > gpurun_out/s5_pin_$tag.json 2>/dev/null
  echo "$1 PIN_P=$2 PIN_S=$3 $(python -c "import json;d=json.loads(open('gpurun_out/s5_pin_$tag.json').read().strip().splitlines()[-1]);print('%.1f us/hop, tail %.1f us'%(d['ms_per_step']*1e3,d['roofline']['kernel_ms']*1e3))")"
}
for cfg in "- -" "0 0" "3 -" "5 -" "7 -" "5 0"; do run c4r8 $cfg; done
for cfg in "- -" "0 0" "17 16" "33 0" "25 24" "0 32"; do run c5 $cfg; done
