"""Pool SpMM (fwd and transposed) at full size, B=256: achieved GB/s of algorithmic bytes vs the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn
from tests.golden.loader import Hierarchy
dev = "cuda:0"; B = 256
dt = torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6555.2
h = Hierarchy("2222")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
chans = {("D", 0): 16, ("D", 1): 32, ("D", 2): 64, ("D", 3): 128, ("U", 3): 128, ("U", 2): 64, ("U", 1): 32, ("U", 0): 32}
for (kind, l), C in chans.items():
    pm = shb.PoolMatrix.from_scipy_padded((h.D_sp if kind == "D" else h.U_sp)[l], dev)
    x = torch.randn(B, pm.rows_in, C, device=dev).to(dt).requires_grad_(True)
    fn.TIMER = None
    for i in range(7):
        if i == 2: fn.TIMER = fn.KernelTimer()
        flush.zero_()
        y = shb.pool(x, pm)
        flush.zero_()
        y.backward(torch.ones_like(y))
    for k, v in fn.TIMER.summary().items():
        ms = v["ms"] / v["launches"]; gbs = v["bytes"] / v["launches"] / ms / 1e6
        print(f"{kind}{l} C={C:3d} {k:34s} {ms*1e3:8.1f} us {gbs:8.1f} GB/s  {100*gbs/peak:5.1f}% of measured HBM peak")
