"""Synthetic stand-ins for the assets the reference downloads (SMPL template, DFAUST meshes, weights).

SURVEY.md section 8(d): the real SMPL template is not available offline, so throughput and parity runs
use a closed genus-0 triangulation with SMPL's exact counts (V=6890, F=2V-4=13776, E=20664;
the reference hard-codes 6890/13776 at train_funcs.py:81,84).
"""
import numpy as np
import torch


def make_template(n_verts=6890, seed=0, scale=(0.3, 0.9, 0.2)):
    """Convex hull of ``n_verts`` random unit vectors, outward-oriented faces, anisotropically scaled.

    Returns (verts float64 (V,3), faces int64 (F,3)) with F = 2V-4.
    """
    from scipy.spatial import ConvexHull

    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n_verts, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    hull = ConvexHull(p)
    f = hull.simplices.astype(np.int64)
    a, b, c = p[f[:, 0]], p[f[:, 1]], p[f[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, c - a), a + b + c) < 0
    f[flip] = f[flip][:, [0, 2, 1]]
    assert len(np.unique(f)) == n_verts and len(f) == 2 * n_verts - 4
    return p * np.asarray(scale, dtype=np.float64), f


def make_open_template(n_side=12):
    """Small open (boundary) mesh: a regular triangulated grid patch with a gentle bump; exercises the
    boundary branches of the spiral construction (utils_spiral.py:193-194, 214-255, 373-403)."""
    xs, ys = np.meshgrid(np.arange(n_side, dtype=np.float64), np.arange(n_side, dtype=np.float64), indexing="ij")
    rng = np.random.default_rng(3)
    v = np.stack([xs.ravel(), ys.ravel(), 0.3 * np.sin(xs.ravel() * 0.7) * np.cos(ys.ravel() * 0.5)], axis=1)
    v[:, :2] += 0.15 * rng.standard_normal((len(v), 2))
    f = []
    for i in range(n_side - 1):
        for j in range(n_side - 1):
            a, b, c, d = i * n_side + j, (i + 1) * n_side + j, (i + 1) * n_side + j + 1, i * n_side + j + 1
            if (i + j) % 2 == 0:
                f += [(a, b, c), (a, c, d)]
            else:
                f += [(a, b, d), (b, c, d)]
    return v, np.asarray(f, dtype=np.int64)


def synthetic_meshes(template_verts, batch, seed=0, noise=0.01, dtype=torch.float32):
    """(B, V+1, 3) batch: template + noise, with the zero dummy vertex appended
    (the dataset appends it at autoencoder_dataset.py:26-57)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.as_tensor(np.asarray(template_verts), dtype=dtype)
    x = t[None] + noise * torch.randn((batch,) + tuple(t.shape), generator=g, dtype=dtype)
    return torch.cat([x, torch.zeros(batch, 1, 3, dtype=dtype)], dim=1).contiguous()


@torch.no_grad()
def fill_deterministic_(module, seed=2):
    """Overwrite every parameter (state_dict order) with U(-1/sqrt(fan_in), 1/sqrt(fan_in)) drawn from a
    seeded CPU generator.  Used so that the same weights can be rebuilt on a box that has neither the
    reference nor a checkpoint (the reference seeds with cfgs.py:46 and takes nn.Linear's default init)."""
    g = torch.Generator().manual_seed(seed)
    for name, p in module.state_dict().items():
        fan_in = p.shape[1] if p.dim() == 2 else p.shape[0]
        if p.dim() == 1:
            # biases: fan_in unknown from shape alone; a fixed small range keeps activations O(1)
            bound = 0.05
        else:
            bound = 1.0 / float(np.sqrt(fan_in))
        v = (torch.rand(p.shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        p.copy_(v.to(p.dtype))
    return module
