"""ncu target: two forward + backward passes of the bone-guided model's heads at full size (B = 256), e.g.
   ncu --metrics gpu__time_duration.sum -k regex:gl_ --csv python scripts/profile_heads.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantichuman_b200 as shb
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes
from tests.golden.constants import KPS_INDEX_LIST, PART_LIST

dev = "cuda:0"
FENC = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
FDEC = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
h = Hierarchy("2222")
Dsp, Usp = h.sparse_DU()
vc = h.level_verts(h.n_levels)
order = np.argsort(vc[:, 1], kind="stable")
parts = {n: np.sort(c) for n, c in zip(PART_LIST, np.array_split(order, len(PART_LIST)))}
model = shb.SpiralAutoencoder_multiz_partkps(KPS_INDEX_LIST, parts, FENC, FDEC, latent_size=8, part_kps_latent_size=8,
                                             sizes=h.sizes, spiral_sizes=h.spiral_sizes, spirals=h.spirals(dev), D=Dsp, U=Usp,
                                             device=dev)
fill_deterministic_(model, seed=2)
model = model.to(dev).set_compute_dtype(torch.bfloat16)
xs = synthetic_meshes(h.verts0, 256, seed=3).to(dev)
kps = torch.randn(256, 32, 3, device=dev)
for _ in range(2):
    model.zero_grad(set_to_none=True)
    xh, z, zk = model(xs, kps)
    shb.l1_loss(xs, xh).backward()
torch.cuda.synchronize()
