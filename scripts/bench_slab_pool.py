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
