"""Hierarchy builder: semantichuman_b200.hierarchy must reproduce, exactly, what the reference's
mesh_sampling.generate_transform_matrices produced for every fixture -- kept vertices (D), faces of every level in
order (F), up-sampling matrices (U) -- and, chained with semantichuman_b200.spirals, the spiral tables: the whole setup
of main.py:93-205 from a raw template, without psbody / opendr."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from semantichuman_b200 import hierarchy as hy
from semantichuman_b200 import spirals as spr
from tests.golden.loader import Hierarchy


@pytest.mark.parametrize("tag", ["small", "open", "2222", "4444"])
def test_hierarchy_matches_reference_fixture(tag):
    h = Hierarchy(tag)
    factors = [int(f) for f in h.raw["factors"]]
    out = hy.build_hierarchy(h.verts0, h.faces[0], factors)
    assert [len(v) for v, _ in out["M_verts_faces"]] == h.sizes
    for l in range(len(factors)):
        d = out["D"][l].tocsr()
        assert d.shape == (h.sizes[l + 1], h.sizes[l]) and (d.data == 1.0).all()
        assert np.array_equal(d.indices, h.raw[f"D{l}_col"])
        assert np.array_equal(out["F"][l], h.faces[l + 1])
        assert np.array_equal(out["M_verts_faces"][l + 1][0], h.level_verts(l + 1))
        u = out["U"][l].tocsr()
        ref = h.U_sp[l]
        assert u.shape == ref.shape and u.nnz == ref.nnz == 3 * h.sizes[l]  # three stored entries per row
        assert abs(u - ref).max() == 0.0
        a = out["A"][l + 1]
        assert (a != a.T).nnz == 0 and a.diagonal().sum() == 0
    if tag == "small":  # the rest of the setup from the same raw template: reference points and spirals
        from sklearn.metrics.pairwise import euclidean_distances

        verts = [v for v, _ in out["M_verts_faces"]]
        faces = [h.faces[0]] + out["F"]
        refpts = [[h.refpts[0]]]
        for l in range(1, len(verts)):  # main.py:161-167: nearest coarse vertex to the level-0 reference vertex
            refpts.append(np.argmin(euclidean_distances(verts[l], verts[0][refpts[0]]), axis=0).tolist())
        assert [r[0] for r in refpts] == h.refpts
        tables, sizes, _ = spr.generate_spirals([2, 2, 1, 1, 1], verts, faces, refpts, dilation=[2, 2, 1, 1, 1])
        assert sizes == h.spiral_sizes
        for l, t in enumerate(tables):
            assert np.array_equal(t[0].astype(np.int64), h.spirals_np[l])


def test_cache_round_trip(tmp_path):
    h = Hierarchy("open")
    out = hy.build_hierarchy(h.verts0, h.faces[0], [2, 2])
    assert hy.cache_name([2, 2, 2, 2]) == "downsampling_matrices2222.pkl"  # main.py:93
    path = os.path.join(tmp_path, hy.cache_name([2, 2]))
    hy.save_cache(path, out)
    back = hy.load_cache(path)
    assert set(back) == {"M_verts_faces", "A", "D", "U", "F"}
    for l in range(2):
        assert abs(back["D"][l] - out["D"][l]).max() == 0 and abs(back["U"][l] - out["U"][l]).max() == 0
        assert np.array_equal(back["F"][l], out["F"][l])
    with pytest.raises(ValueError):
        import pickle
        bad = os.path.join(tmp_path, "bad.pkl")
        with open(bad, "wb") as fh:
            pickle.dump({"D": []}, fh)
        hy.load_cache(bad)


def test_decimation_invariants_and_closest_point():
    h = Hierarchy("small")
    new_f, kept = hy.decimate(h.verts0, h.faces[0], n_verts_desired=100)
    assert len(kept) == 100 and np.all(np.diff(kept) > 0)
    assert new_f.min() == 0 and new_f.max() == 99
    assert not np.any((new_f[:, 0] == new_f[:, 1]) | (new_f[:, 1] == new_f[:, 2]) | (new_f[:, 2] == new_f[:, 0]))
    q = hy.vertex_quadrics(h.verts0, h.faces[0])
    assert np.allclose(q, np.transpose(q, (0, 2, 1))) and q.shape == (h.sizes[0], 4, 4)
    # point-triangle distance: all seven regions
    a, b, c = np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([0.0, 1, 0])
    cases = {4: [-1, -1, 0.5], 5: [2, -0.5, 0], 6: [-0.5, 2, 0], 1: [0.5, -1, 0], 3: [-1, 0.5, 0], 2: [1, 1, 0], 0: [0.2, 0.3, 1]}
    for code, p in cases.items():
        q_pt, got = hy.closest_point_on_triangle(np.array(p, dtype=float), a, b, c)
        assert got == code
        assert abs(q_pt[2]) < 1e-15 and q_pt[0] >= -1e-15 and q_pt[1] >= -1e-15 and q_pt[0] + q_pt[1] <= 1 + 1e-15
    # rows of U reproduce the fine vertices that lie on the coarse surface (kept vertices map to themselves)
    u = hy.upsampling_matrix(h.level_verts(1), h.faces[1], h.verts0).tocsr()
    kept1 = h.raw["D0_col"]
    assert np.allclose(u[kept1].dot(h.level_verts(1)), h.verts0[kept1], atol=1e-12)
