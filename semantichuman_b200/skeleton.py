"""Keypoint / skeleton glue around the bone-guided model: keypoint regression, keypoints -> bones,
bones -> keypoints.  Pure tensor code (any device), vectorised over the bones -- the reference loops over the 27 bones in
Python (utils_SH.py:26-84) and regresses keypoints with a dense 35 x 6890 matmul per step (train_funcs.py:131).

A bone k of ``skl_list`` is ``[head, tail]`` or ``[head, tail_a, tail_b]`` (tail = mean of two keypoints); its vector is
head - tail.  The full keypoint set has ``len(skl_list) + 4`` entries; data carries it without entries 3, 13, 14
(``kps_keep``, utils_SH.py:32-36).
"""
import numpy as np
import torch

from .models import DEFAULT_NEWSKL_LIST

DROPPED_KEYPOINTS = (3, 13, 14)
SKL_MODES = ("ori_m", "kps_ori_m", "vec_m", "vec", "m")


class Skeleton:
    def __init__(self, skl_list=None):
        self.skl_list = [list(b) for b in (DEFAULT_NEWSKL_LIST if skl_list is None else skl_list)]
        self.n_bones = len(self.skl_list)
        self.n_kps = self.n_bones + 4
        self.keep = [i for i in range(self.n_kps) if i not in DROPPED_KEYPOINTS]
        self._head = torch.tensor([b[0] for b in self.skl_list])
        self._tail_a = torch.tensor([b[1] for b in self.skl_list])
        self._tail_b = torch.tensor([b[2] if len(b) == 3 else b[1] for b in self.skl_list])
        # skl2kps walks the bones in order: kps[tail] = kps[head] - vec (utils_SH.py:71-79), keypoint 0 and every keypoint
        # that is never a tail stay at the origin.  Unrolled: kps = -path @ vec with path[j, k] = 1 when bone k lies on
        # the chain that produced keypoint j (later assignments overwrite earlier ones, as in the loop).
        path = np.zeros((self.n_kps, self.n_bones), np.float32)
        for k, b in enumerate(self.skl_list):
            path[b[1]] = path[b[0]]
            path[b[1], k] = 1.0
        self._path = torch.from_numpy(path)

    def _on(self, t, ref):
        return t.to(ref.device)

    def expand(self, kps):
        """(B, n_kps - 3, 3) -> (B, n_kps, 3) with zeros at the dropped keypoints; full sets pass through (copied)."""
        if kps.shape[1] == self.n_kps:
            return kps.clone()
        if kps.shape[1] != len(self.keep):
            raise ValueError(f"expected {len(self.keep)} or {self.n_kps} keypoints, got {kps.shape[1]}")
        full = torch.zeros((kps.shape[0], self.n_kps, 3), dtype=kps.dtype, device=kps.device)
        full[:, self.keep, :] = kps
        return full

    def kps2skl(self, kps, mode="ori_m"):
        """utils_SH.py:26-66.  'ori_m'/'kps_ori_m': (unit direction, length) (B, n_bones, 4); 'vec_m': (vector, length);
        'vec': vector (B, n_bones, 3); 'm': length (B, n_bones, 1)."""
        if mode not in SKL_MODES:
            raise NotImplementedError(mode)
        kps = self.expand(kps)
        head = kps[:, self._on(self._head, kps), :]
        tail = (kps[:, self._on(self._tail_a, kps), :] + kps[:, self._on(self._tail_b, kps), :]) / 2
        vec = head - tail
        length = torch.sqrt(torch.sum(vec ** 2, dim=2, keepdim=True))
        if mode in ("ori_m", "kps_ori_m"):
            return torch.cat([vec / length, length], dim=2)
        if mode == "vec_m":
            return torch.cat([vec, length], dim=2)
        return vec if mode == "vec" else length

    def skl2kps(self, skl, mode="ori_m"):
        """utils_SH.py:68-80: rebuild the keypoints from bone vectors, root (keypoint 0) at the origin; returns the kept
        keypoints (B, n_kps - 3, 3)."""
        if mode == "vec":
            vec = skl
        elif mode == "vec_m":
            vec = skl[:, :, :3]
        elif mode in ("ori_m", "kps_ori_m"):
            vec = skl[:, :, :3] * skl[:, :, 3:]
        else:
            raise NotImplementedError(mode)
        kps = -torch.matmul(self._on(self._path, skl).to(skl.dtype), vec)
        return kps[:, self.keep, :]


def regress_keypoints(j_regressor, verts):
    """kps = J_regressor @ verts (train_funcs.py:131).  `j_regressor` (n_kps, V) dense or torch sparse (SMPL's regressor has
    a few hundred non-zeros of 35 x 6890: keep it sparse); `verts` (B, V, 3) without the dummy row."""
    if j_regressor.is_sparse or j_regressor.layout != torch.strided:
        B, V, _ = verts.shape
        flat = verts.permute(1, 0, 2).reshape(V, B * 3)
        return torch.sparse.mm(j_regressor, flat).reshape(-1, B, 3).permute(1, 0, 2).float()
    return torch.matmul(j_regressor, verts).float()
