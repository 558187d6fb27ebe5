"""Time one SpiralConv layer (fwd / dgrad / wgrad) at full size; target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn
from semantichuman_b200.assets import Hierarchy

lvl, cin, cout = (int(v) for v in os.environ.get("LAYER", "0,32,16").split(","))
B = int(os.environ.get("B", "256")); reps = int(os.environ.get("REPS", "5"))
dt = torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32
dev = "cuda:0"
h = Hierarchy("2222")
geom = shb.SpiralGeometry.from_spiral(h.spirals(dev)[lvl], dev)
x = torch.randn(B, geom.rows_in, cin, device=dev).to(dt).requires_grad_(True)
w = (torch.randn(cout, geom.S * cin, device=dev) / (geom.S * cin) ** 0.5).requires_grad_(True)
b = torch.zeros(cout, device=dev, requires_grad=True)
gy = torch.randn(B, geom.rows_out, cout, device=dev).to(dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(reps + 2):
    if i == 2:
        fn.TIMER = fn.KernelTimer()
    flush.zero_()
    y = shb.spiral_conv(x, w, b, geom, "elu")
    flush.zero_()
    y.backward(gy)
per = fn.TIMER.summary()
for k, v in per.items():
    ms = v["ms"] / v["launches"]
    print(f"{k:50s} {ms:8.4f} ms  {v['bytes']/v['launches']/ms/1e6:8.1f} GB/s  {v['flops']/v['launches']/ms/1e9:8.1f} TFLOP/s")
