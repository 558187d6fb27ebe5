"""Batched forms of the reference's per-sample auxiliary losses: no Python loop over the batch, no
``.cpu().numpy()`` round trip per sample (train_funcs.py:137-143, 325-333), deterministic reductions (a one-hot matmul
instead of an atomic scatter).  Pure tensor code, any device; pinned against the reference's own functions
(tests/golden/golden_aux_losses.npz).
"""
import torch
import torch.nn.functional as F


def edge_ratio_loss(rec, gt, faces, eps=1e-5):
    """train_funcs.py:137-143 = mean over the batch of compute_score(rec[i], faces, get_target(gt[i])) (:22-40):
    mean over faces of sum over the three edges of | |e_rec| / (|e_gt| + eps) - 1 |.  rec, gt (B, V, 3); faces (F, 3)."""
    faces = torch.as_tensor(faces, dtype=torch.long, device=rec.device)
    a, b, c = faces[:, 0], faces[:, 1], faces[:, 2]

    def lengths(v):
        return (torch.sqrt(torch.sum((v[:, a] - v[:, b]) ** 2, dim=2)), torch.sqrt(torch.sum((v[:, b] - v[:, c]) ** 2, dim=2)),
                torch.sqrt(torch.sum((v[:, a] - v[:, c]) ** 2, dim=2)))

    with torch.no_grad():  # the target comes from numpy in the reference: no gradient
        target = [t.float() + eps for t in lengths(gt)]
    score = sum(torch.abs(l / t - 1) for l, t in zip(lengths(rec), target))
    return torch.mean(score)


def edge_length_loss(inp, rec, edge_verts_index):
    """Edge_loss (train_funcs.py:42-45): L1 between the lengths of the listed vertex pairs."""
    e = torch.as_tensor(edge_verts_index, dtype=torch.long, device=rec.device)

    def lengths(v):
        return torch.sqrt(torch.sum((v[:, e[:, 0], :] - v[:, e[:, 1], :]) ** 2, dim=2))

    return F.l1_loss(lengths(rec), lengths(inp))


class PartVolumes:
    """Signed part volumes for cal_volloss (train_funcs.py:56-72): a face belongs to a part when its three vertices do
    (train_funcs.py:84-89); volume of a part = sum over its faces of (a x b) . c."""

    def __init__(self, faces, vert_part_index_dict, device="cpu"):
        self.faces = torch.as_tensor(faces, dtype=torch.long, device=device)
        n_parts = len(vert_part_index_dict)
        n_verts = int(self.faces.max()) + 1
        vp = torch.full((max(n_verts, 1 + max(int(torch.as_tensor(v).max()) for v in vert_part_index_dict.values())),), -1,
                        dtype=torch.long)
        for k, v in enumerate(vert_part_index_dict.values()):
            vp[torch.as_tensor(v, dtype=torch.long)] = k
        f = self.faces.cpu()
        same = (vp[f[:, 0]] == vp[f[:, 1]]) & (vp[f[:, 0]] == vp[f[:, 2]]) & (vp[f[:, 0]] >= 0)
        part = torch.where(same, vp[f[:, 0]], torch.full_like(vp[f[:, 0]], -1))
        onehot = torch.zeros(len(f), n_parts)
        onehot[same, part[same]] = 1.0
        self.onehot = onehot.to(device)  # (F, P): the per-part sums are one deterministic matmul
        self.n_parts = n_parts

    def volumes(self, verts):
        a, b, c = (verts[:, self.faces[:, k], :] for k in range(3))
        per_face = torch.sum(torch.cross(a, b, dim=2) * c, dim=2)  # (B, F)
        return per_face @ self.onehot.to(per_face.dtype)           # (B, P)

    def loss(self, rec, gt, parts_used):
        """Mean over the batch of cal_volloss(rec[i], gt[i], ..., parts_used): (1/len(parts_used)) * sum_p | |V_rec/V_gt| - 1 |."""
        parts_used = torch.as_tensor(list(parts_used), dtype=torch.long, device=rec.device)
        vr = self.volumes(rec)[:, parts_used]
        vg = self.volumes(gt)[:, parts_used]
        return torch.mean(torch.sum(torch.abs(torch.abs(vr / vg) - torch.abs(vg / vg)), dim=1) / len(parts_used))
