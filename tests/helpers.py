"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

from tests.golden.loader import Hierarchy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerances, per-tensor max|a-b| / max|b|  (SURVEY.md section 7 hard part 1)
TOL_F32 = 1e-4
TOL_BF16 = 2e-2


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def params_from_golden(g, dtype=torch.float32, prefix="p_"):
    return {k[len(prefix):]: torch.from_numpy(g[k]).to(dtype) for k in g.files if k.startswith(prefix)}


def grads_from_golden(g):
    return {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("g_")}


def filters_from_golden(g):
    fe = [g["filters_enc0"].tolist(), [int(v) if v else [] for v in g["filters_enc1"]]]
    fd = [g["filters_dec0"].tolist(), [int(v) if v else [] for v in g["filters_dec1"]]]
    return fe, fd


def ref_args(tag, cfg="A", dtype=torch.float32):
    """(hier, sizes, spiral_sizes, spirals, D_dense, U_dense) as main.py:183-205 builds them, on CPU."""
    h = Hierarchy(tag, cfg)
    D, U = h.dense_DU()
    return h, h.sizes, h.spiral_sizes, h.spirals(), [d.to(dtype) for d in D], [u.to(dtype) for u in U]


DEFAULT_FENC = [[3, 16, 32, 64, 128], [[], [], [], [], []]]  # configure/cfgs.py:11
DEFAULT_FDEC = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]  # configure/cfgs.py:12
