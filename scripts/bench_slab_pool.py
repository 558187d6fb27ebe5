"""Time the Pool kernels of the slab path at full size (forward U / D and their transposes), L2 flushed; ncu target.
   LEVELS=0,1,2,3  B=256  DT=bf16|fp32"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn, slab
from tests.golden.loader import Hierarchy

dev = "cuda:0"
B = int(os.environ.get("B", "256"))
planes = 1 if os.environ.get("DT", "bf16") == "bf16" else 2
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
h = Hierarchy("2222")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
chans = {0: 32, 1: 32, 2: 64, 3: 128}
for l in (int(v) for v in os.environ.get("LEVELS", "0,1,2,3").split(",")):
    pm = shb.PoolMatrix.from_scipy_padded(h.U_sp[l], dev)
    C = chans[l]
    xs = slab.from_rows(torch.randn(B, pm.rows_in, C, device=dev), None, planes)
    xs.t.requires_grad_(True); xs.act = 2; xs.masked = True
    gy = slab.from_rows(torch.randn(B, pm.rows_out, C, device=dev), None, planes).t
    fn.TIMER = None
    for i in range(7):
        if i == 2:
            fn.TIMER = fn.KernelTimer()
        flush.zero_()
        y = slab.pool(xs, pm)
        flush.zero_()
        y.t.backward(gy)
    for k, v in fn.TIMER.summary().items():
        ms = v["ms"] / v["launches"]; gbs = v["bytes"] / v["launches"] / ms / 1e6
        print(f"U{l} C={C:3d} {k:34s} {ms*1e3:8.1f} us {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of measured HBM peak")

# Reference points for the same traffic pattern, timed the same way (L2 flushed by a 256 MB zero-fill, i.e. 126 MB of dirty
# lines that the next kernel's traffic has to push out): what does a plain torch copy reach for U0's bytes (read 56 MB, write
# 113 MB), and for a straight 113 MB -> 113 MB copy?
def timed(f, n=5):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        flush.zero_(); a.record(); f(); b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in ev)
src = torch.empty(56 << 20, dtype=torch.uint8, device=dev); big = torch.empty(113 << 20, dtype=torch.uint8, device=dev)
big2 = torch.empty_like(big)
for name, f, nbytes in (("torch copy 56 MB -> 56 MB", lambda: big[:56 << 20].copy_(src), 2 * 56 * 2 ** 20),
                        ("torch copy 113 MB -> 113 MB", lambda: big2.copy_(big), 2 * 113 * 2 ** 20),
                        ("torch fill 113 MB", lambda: big.zero_(), 113 * 2 ** 20)):
    ms = timed(f)
    print(f"{name:34s} {ms*1e3:8.1f} us {nbytes/ms/1e6:8.1f} GB/s  {100*nbytes/ms/1e6/peak:5.1f}% of measured HBM peak")
# and without the flush (steady state of a step: the previous kernel's output is what sits dirty in L2)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
big2.copy_(big); ev0.record()
for _ in range(10):
    big2.copy_(big)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
print(f"{'torch copy 113 MB, back to back':34s} {ms*1e3:8.1f} us {2*113*2**20/ms/1e6:8.1f} GB/s")
