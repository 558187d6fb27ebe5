"""Skeleton glue: vectorised kps2skl / skl2kps and keypoint regression against numbers produced by the
reference's own utils_SH.kps2skl / skl2kps (golden_skeleton.npz)."""
import numpy as np
import pytest
import torch

from semantichuman_b200.skeleton import Skeleton, regress_keypoints
from tests.helpers import golden, relerr


@pytest.mark.parametrize("mode", ["ori_m", "vec_m", "vec", "m"])
def test_kps2skl_matches_reference(mode):
    g = golden("golden_skeleton")
    sk = Skeleton()
    for which in ("full", "keep"):
        got = sk.kps2skl(torch.from_numpy(g["kps_" + which]), mode)
        assert got.shape == g[f"skl_{which}_{mode}"].shape
        assert relerr(got, g[f"skl_{which}_{mode}"]) < 1e-6


@pytest.mark.parametrize("mode", ["ori_m", "vec_m", "vec"])
def test_skl2kps_matches_reference_and_inverts(mode):
    g = golden("golden_skeleton")
    sk = Skeleton()
    back = sk.skl2kps(torch.from_numpy(g["skl_keep_" + mode]), mode)
    assert relerr(back, g["back_" + mode]) < 1e-5
    # bones of the rebuilt keypoints are the bones we started from (two-keypoint bones; the root is moved to the origin)
    again = sk.kps2skl(back, "vec")
    two = [k for k, b in enumerate(sk.skl_list) if len(b) == 2]
    assert relerr(again[:, two], g["skl_keep_vec"][:, two]) < 1e-5


def test_unknown_mode_and_shapes_raise():
    sk = Skeleton()
    with pytest.raises(NotImplementedError):
        sk.kps2skl(torch.zeros(1, 31, 3), "nope")
    with pytest.raises(ValueError):
        sk.kps2skl(torch.zeros(1, 30, 3))


def test_regress_keypoints_dense_and_sparse_agree():
    gen = torch.Generator().manual_seed(0)
    j = torch.rand(35, 200, generator=gen) * (torch.rand(35, 200, generator=gen) < 0.05)
    j = j / j.sum(1, keepdim=True).clamp_min(1e-6)
    v = torch.randn(3, 200, 3, generator=gen)
    dense = regress_keypoints(j, v)
    sparse = regress_keypoints(j.to_sparse(), v)
    assert dense.shape == (3, 35, 3) and torch.allclose(dense, sparse, atol=1e-6)
