#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2p_pytest.log; tail -30 gpurun_out/r2p_pytest.log
for nh in 4 8; do python tools/mh_probe.py $nh 20; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cmac_mh2 -s 2 -c 1 -o gpurun_out/r2p_mh8 -f python tools/mh_probe.py 8 > gpurun_out/r2p_mh8.log 2>&1
ncu -i gpurun_out/r2p_mh8.ncu-rep --page details --csv > gpurun_out/r2p_mh8_details.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_mh8_details.csv')))
hdr=rows[0]; i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
print(' '.join('%s=%s'%(r[i_m],r[i_v]) for r in rows[1:] if r[i_m] in ('Duration','DRAM Throughput','Registers Per Thread','Issue Slots Busy','Dynamic Shared Memory Per Block')))
PY
