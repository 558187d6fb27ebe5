"""Loader for the hierarchy fixtures (tests/golden/hier_*.npz): template, D/U, spirals.

The fixtures are produced by the reference's own mesh_sampling.py / utils_spiral.py (tests/golden/make_golden.py);
this module (test infrastructure, like the fixtures it reads) only unpacks them into the constructor arguments of the models -- either exactly as main.py:183-205
builds them (dense padded D/U, (1, V+1, S) int64 spirals) or in the sparse form the product accepts directly.
"""
import os

import numpy as np
import scipy.sparse as sp
import torch

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))


class Hierarchy:
    def __init__(self, path_or_tag, spiral_cfg="A"):
        path = path_or_tag if os.path.exists(path_or_tag) else os.path.join(GOLDEN_DIR, f"hier_{path_or_tag}.npz")
        h = np.load(path)
        self.raw = h
        self.sizes = [int(v) for v in h["sizes"]]
        self.n_levels = len(self.sizes) - 1
        self.verts0 = h["verts0"]
        self.faces = [h["faces0"]] + [h[f"faces{l}"] for l in range(1, self.n_levels + 1)]
        self.refpts = [int(v) for v in h["refpts"]]
        self.spiral_sizes = [int(v) for v in h[f"sp{spiral_cfg}_sizes"]]
        self.spirals_np = [h[f"sp{spiral_cfg}{l}"].astype(np.int64) for l in range(self.n_levels + 1)]
        self.D_sp, self.U_sp = [], []
        for l in range(self.n_levels):
            n_out, n_in = self.sizes[l + 1], self.sizes[l]
            self.D_sp.append(sp.csr_matrix((np.ones(n_out), (np.arange(n_out), h[f"D{l}_col"])), shape=(n_out, n_in)))
            self.U_sp.append(sp.csr_matrix((h[f"U{l}_data"], h[f"U{l}_indices"], h[f"U{l}_indptr"]), shape=(n_in, n_out)))

    def level_verts(self, l):
        v = self.verts0
        for i in range(l):
            v = self.D_sp[i].dot(v)
        return v

    def spirals(self, device="cpu"):
        """[(1, V_l+1, S_l) int64]  -- main.py:203"""
        return [torch.from_numpy(s[None]).to(device) for s in self.spirals_np]

    @staticmethod
    def _pad_dense(m):
        d = np.zeros((1, m.shape[0] + 1, m.shape[1] + 1))
        d[0, :-1, :-1] = m.todense()
        d[0, -1, -1] = 1
        return torch.from_numpy(d).float()

    def dense_DU(self, device="cpu"):
        """Dense padded fp32 D/U lists exactly as main.py:183-205 builds them."""
        return ([self._pad_dense(m).to(device) for m in self.D_sp], [self._pad_dense(m).to(device) for m in self.U_sp])

    def sparse_DU(self):
        """Un-padded scipy matrices; the product pads them in CSR form (PoolMatrix.from_scipy_padded)."""
        return list(self.D_sp), list(self.U_sp)
