#!/usr/bin/env python
"""Condense an `ncu --page source --csv --print-source sass` export: instruction mix, stall reasons, hottest SASS lines."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(int(r[ix['Instructions Executed']]) for r in data)
tot_samp = sum(int(r[ix['# Samples']]) for r in data)
print(rows[0][1])
print('warp instructions', tot_inst, 'samples', tot_samp)
agg = defaultdict(lambda: [0, 0])
for r in data:
    toks = r[ix['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    agg[op][0] += int(r[ix['Instructions Executed']]); agg[op][1] += int(r[ix['# Samples']])
for op, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('%-10s inst %11d (%4.1f%%)  samples %7d (%4.1f%%)' % (op, n, 100 * n / tot_inst, s, 100 * s / tot_samp))
for h in [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]:
    v = sum(int(r[ix[h]]) for r in data)
    if v * 100 > tot_samp: print('%-24s %5.1f%%' % (h, 100 * v / tot_samp))
top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]
for r in top:
    print('%6d  %s' % (int(r[ix['# Samples']]), r[ix['Source']].strip()[:100]))
bc = sum(int(r[ix['L1 Wavefronts Shared Excessive']] or 0) for r in data); wf = sum(int(r[ix['L1 Wavefronts Shared']] or 0) for r in data)
print('shared wavefronts', wf, 'excessive', bc)
