"""Fused orientation-adaptive pairwise-distance loss at full size: 6890 vertices in 17 parts (y-sorted slabs), B=256.
The reference materialises 2 x (B, n, n, 3) + 4 x (B, n, n) per part; this path keeps nothing of size n x n."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantichuman_b200 as shb
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import synthetic_meshes

dev = "cuda:0"; B = 256
h = Hierarchy("2222")
V = h.sizes[0]
order = np.argsort(np.asarray(h.verts0)[:V, 1], kind="stable")
parts = [np.sort(c) for c in np.array_split(order, 17)]
skl = [[15, 12], [15, 12], [12, 9], [6, 0], [0, 1, 2], [1, 4], [4, 7], [7, 10], [2, 5], [5, 8], [8, 11], [16, 18], [18, 20],
       [20, 22], [17, 19], [19, 21], [21, 23]]  # cfgs.py:18-20
lay = shb.PairLossLayout(parts, skl, dev, w_mode="linear", leaf_parts=(0, 7, 10, 13, 16))
tx = synthetic_meshes(h.verts0, B, seed=1)[:, :-1].contiguous().to(dev)
gen = torch.Generator(device=dev).manual_seed(0)
rec = (tx + 0.01 * torch.randn(tx.shape, device=dev, generator=gen)).requires_grad_(True)
kps = torch.randn(B, 24, 3, device=dev, generator=gen)
def step():
    rec.grad = None
    loss = shb.pair_loss(tx, rec, kps, lay)
    loss.backward()
    return loss
for _ in range(3):
    step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    loss = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
pairs = B * lay.pairs_per_sample
print(json.dumps({"ms_fwd_bwd": ms, "pairs": pairs, "gpairs_per_s": 2 * pairs / ms / 1e6, "loss": float(loss.detach()),
                  "reference_intermediates_bytes_per_part": int(B * 405 * 405 * (2 * 3 + 4) * 4)}))
