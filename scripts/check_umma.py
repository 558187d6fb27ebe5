"""bf16 tcgen05 SpiralConv (fwd + dgrad) vs the exact-fp32 CUDA-core kernels on the same bf16-rounded operands."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from tests.golden.loader import Hierarchy
from tests.helpers import relerr

B = int(os.environ.get("B", "5"))
h = Hierarchy("2222")
dev = "cuda:0"
layers = [(1, 16, 32, "elu"), (2, 32, 64, "elu"), (3, 64, 128, "elu"), (3, 128, 64, "elu"), (2, 64, 32, "elu"),
          (1, 32, 32, "tanh"), (0, 32, 16, "elu"), (0, 16, 3, "identity"), (0, 16, 16, "relu"), (0, 3, 16, "elu"), (2, 3, 8, "elu")]
gen = torch.Generator(device=dev).manual_seed(0)
worst = 0
for lvl, cin, cout, act in layers:
    geom = shb.SpiralGeometry.from_spiral(h.spirals(dev)[lvl], dev)
    x = torch.randn(B, geom.rows_in, cin, device=dev, generator=gen).bfloat16()
    w = (torch.randn(cout, geom.S * cin, device=dev, generator=gen) / (geom.S * cin) ** 0.5).bfloat16().float()
    b = torch.randn(cout, device=dev, generator=gen) * 0.1
    gy = torch.randn(B, geom.rows_out, cout, device=dev, generator=gen).bfloat16()
    xa = x.clone().requires_grad_(True); wa = w.clone().requires_grad_(True); ba = b.clone().requires_grad_(True)
    y = shb.spiral_conv(xa, wa, ba, geom, act)
    y.backward(gy)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    yr = shb.spiral_conv(xr, wr, br, geom, act)
    # differentiate the fp32 reference through the bf16-rounded output the bf16 path saw
    yr.backward(gy.float())
    torch.cuda.synchronize()
    e = (relerr(y.float(), yr), relerr(xa.grad.float(), xr.grad), relerr(wa.grad, wr.grad), relerr(ba.grad, br.grad))
    worst = max(worst, *e)
    print(f"L{lvl} {cin:3d}->{cout:3d} {act:8s} fwd {e[0]:.2e} gx {e[1]:.2e} gw {e[2]:.2e} gb {e[3]:.2e}", flush=True)
print("worst", worst, "OK" if worst < 2e-2 else "FAIL")
