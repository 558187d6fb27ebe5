"""Bit-exact spiral construction: semantichuman_b200.spirals, written from the behavioural spec, must
reproduce the tables the reference's utils_spiral.generate_spirals produced for every fixture hierarchy -- closed and
open (boundary) meshes, 1-ring and 2-ring dilated configurations, all levels -- including the spiral lengths."""
import numpy as np
import pytest

from semantichuman_b200 import spirals as sp
from tests.golden.loader import Hierarchy

CONFIGS = {"A": ([2, 2, 1, 1, 1], [2, 2, 1, 1, 1]), "B": ([1] * 5, [1] * 5)}


@pytest.mark.parametrize("tag", ["small", "open", "2222", "4444"])
@pytest.mark.parametrize("cfg", ["A", "B"])
def test_spirals_bit_exact(tag, cfg):
    h = Hierarchy(tag, cfg)
    steps, dil = CONFIGS[cfg]
    n = h.n_levels + 1
    verts = [h.level_verts(l) for l in range(n)]
    tables, sizes, _ = sp.generate_spirals(steps[:n], verts, h.faces, [[r] for r in h.refpts], dilation=dil[:n])
    assert sizes == h.spiral_sizes
    for l in range(n):
        assert tables[l].shape == (1, h.sizes[l] + 1, sizes[l]) and tables[l].dtype == np.float64
        assert np.array_equal(tables[l][0].astype(np.int64), h.spirals_np[l]), (tag, cfg, l)
        assert (tables[l][0, -1] == -1).all()


def test_adjacency_is_sorted_and_symmetric():
    h = Hierarchy("small")
    adj, trig = sp.adjacency_and_triangles(h.sizes[0], h.faces[0])
    for i, nb in enumerate(adj):
        assert nb == sorted(nb) and i not in nb
        assert all(i in adj[j] for j in nb)
        assert all(i in t for t in trig[i])
