"""Time one SpiralConv layer (fwd / dgrad / wgrad) at full size; target for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn
from tests.golden.loader import Hierarchy

lvl, cin, cout = (int(v) for v in os.environ.get("LAYER", "0,32,16").split(","))
B = int(os.environ.get("B", "256")); reps = int(os.environ.get("REPS", "5"))
dt = torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32
dev = "cuda:0"
h = Hierarchy("2222")
spiral = h.spirals(dev)[lvl]
if os.environ.get("PERM"):  # locality experiment: relabel vertices (rows and entries) by a geometric / graph ordering
    import numpy as np, scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    t = spiral[0].cpu().numpy(); V1, S = t.shape; V = V1 - 1
    t = np.where(t < 0, t + V1, t)
    if os.environ["PERM"] == "morton" and lvl == 0:
        v = np.asarray(h.verts0, dtype=np.float64)[:V]
        q = ((v - v.min(0)) / (np.ptp(v, 0) + 1e-9) * 1023).astype(np.int64)
        def part(x):
            x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3
            return (x | (x << 2)) & 0x09249249
        perm = np.argsort(part(q[:, 0]) | (part(q[:, 1]) << 1) | (part(q[:, 2]) << 2), kind="stable")
    else:
        rows = np.repeat(np.arange(V), S); cols = t[:V].ravel(); m = cols < V
        A = sp.coo_matrix((np.ones(m.sum()), (rows[m], cols[m])), shape=(V, V)).tocsr()
        perm = np.asarray(reverse_cuthill_mckee(((A + A.T) > 0).astype(np.int8).tocsr(), symmetric_mode=True))
    inv = np.empty(V1, np.int64); inv[perm] = np.arange(V); inv[V] = V
    spiral = torch.from_numpy(inv[t[np.concatenate([perm, [V]])]])[None].to(dev)
geom = shb.SpiralGeometry.from_spiral(spiral, dev)
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
