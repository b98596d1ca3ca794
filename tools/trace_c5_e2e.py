"""Kernel timeline (hb_conv_set_trace) of config 5 fed from host rows: when do the forward FFTs, partition 0, the tail and the inverse
FFTs of consecutive hops start and end, relative to each other?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from hisstools_library_b200.convolve import _Engine

groups, B, P = 16, 8192, 128
taps = B * P
dev = torch.device("cuda", 0)
eng = _Engine(np.float64, groups, 1, 1, 2 * B, taps, 0, 0, 0)
eng.set_reset_offset(0)
ir = torch.randn(taps, device=dev, dtype=torch.float64) * 0.01
for g in range(groups):
    eng.set_ir_device(g, 0, 0, ir.data_ptr(), taps)
torch.cuda.synchronize()
host = len(sys.argv) < 2 or sys.argv[1] != "device"
x = [np.random.rand(groups, B) for _ in range(4)]
y = np.zeros((groups, B))
px = [eng.row_pointers([xi[r] for r in range(groups)]) for xi in x]
py = eng.row_pointers([y[r] for r in range(groups)])
xd = torch.rand(groups, B, device=dev, dtype=torch.float64)
yd = torch.zeros(groups, B, device=dev, dtype=torch.float64)
st = torch.cuda.Stream()


def call(k):
    if host:
        eng.process_pointers(px[k % 4], py, B)
    else:
        eng.process_device(xd.data_ptr(), B, yd.data_ptr(), B, B, False, st.cuda_stream)


for k in range(40):
    call(k)
torch.cuda.synchronize()
eng.set_trace(True)
for k in range(14):
    call(k)
tr, hop = eng.get_trace()
print("config 5, %s, tail streams %d, hops so far %d; microseconds from the first stamp shown" % ("host rows (hb_conv_process)" if host else "device-resident", eng.tail_streams, hop))
KIND = ["fwd", "head", "tail", "inv"]
t0 = None
for h in range(hop - 9, hop - 1):
    line = "hop %3d" % h
    for k in range(4):
        ent, ext = tr[h % 16, k, 0], tr[h % 16, k, 1]
        m = ent > 0
        if not m.any():
            line += "  %-4s      -      -" % KIND[k]
            continue
        if t0 is None:
            t0 = int(ent[m].min())
        line += "  %-4s %7.1f %7.1f" % (KIND[k], (int(ent[m].min()) - t0) / 1e3, (int(ext[m].max()) - t0) / 1e3)
    print(line)
