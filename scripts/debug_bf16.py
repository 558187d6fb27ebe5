import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import semantichuman_b200 as shb
from tests.helpers import golden, ref_args, relerr
g = golden("golden_ops")
_, sizes, ssz, spirals, _, _ = ref_args("small")
for k in range(int(g["n_conv"])):
    pre = f"conv{k}_"
    lvl, cin, cout, S, B = g[pre + "meta"].tolist()
    act = str(g[pre + "act"])
    geom = shb.SpiralGeometry.from_spiral(spirals[lvl], "cuda:0")
    x = torch.from_numpy(g[pre + "x"]).cuda()
    w = torch.from_numpy(g[pre + "w"]).cuda()
    b = torch.from_numpy(g[pre + "b"]).cuda()
    y32 = shb.spiral_conv(x, w, b, geom, act)
    y16 = shb.spiral_conv(x.bfloat16(), w, b, geom, act)
    y16nb = shb.spiral_conv(x.bfloat16(), w, None, geom, act)
    y32nb = shb.spiral_conv(x, w, None, geom, act)
    torch.cuda.synchronize()
    print(k, lvl, cin, cout, act, "f32 err", relerr(y32, g[pre+"y"]), "bf16 err", relerr(y16.float(), g[pre+"y"]),
          "nobias bf16 vs f32", relerr(y16nb.float(), y32nb), "finite", bool(torch.isfinite(y16.float()).all()))
