"""Inert stand-ins for the third-party packages the reference imports but this image lacks.

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) to import the reference's ``models.py`` / ``utils_spiral.py`` /
``mesh_sampling.py`` *in place and unmodified* so that golden vectors can be generated from the
reference itself.  Nothing here runs on the hot-path arithmetic (that is ATen, called by the
reference's own code); the stand-ins only supply

* ``yacs.config.CfgNode``          -- attribute dict with ``merge_from_file``
* ``opendr.topology``              -- vertex adjacency / unique edge list (mesh_sampling.py:99,231)
* ``psbody.mesh.Mesh``             -- ``v``/``f`` holder + ``compute_aabb_tree().nearest`` returning psbody's
                                      (face, part-code, closest point) triple (mesh_sampling.py:53,72-85)
* ``torch_scatter``, ``trimesh``, ``tensorboardX`` -- import-only placeholders.
"""
import sys
import types

import numpy as np
import scipy.sparse as sp


class CfgNode(dict):
    def __init__(self, init_dict=None, new_allowed=False, **_):
        super().__init__()
        if init_dict:
            for k, v in init_dict.items():
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, path):
        import yaml

        with open(path) as fh:
            self._merge(yaml.safe_load(fh))

    def _merge(self, d):
        for k, v in d.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = v


def get_vert_connectivity(mesh_v, mesh_f):
    n = len(mesh_v)
    f = np.asarray(mesh_f, dtype=np.int64)
    rows = np.concatenate([f[:, 0], f[:, 1], f[:, 2], f[:, 1], f[:, 2], f[:, 0]])
    cols = np.concatenate([f[:, 1], f[:, 2], f[:, 0], f[:, 0], f[:, 1], f[:, 2]])
    m = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    m.data[:] = 1.0
    return m


def get_vertices_per_edge(mesh_v, mesh_f):
    vc = sp.coo_matrix(get_vert_connectivity(mesh_v, mesh_f))
    e = np.stack([vc.row, vc.col], axis=1)
    e = e[e[:, 0] < e[:, 1]]
    return e[np.lexsort((e[:, 1], e[:, 0]))]


def _closest_on_triangle(p, a, b, c):
    """Closest point on triangle abc to p, with psbody part codes:
    0 face, 1..3 edge (k-1 -> k%3), 4..6 vertex k-4."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return a, 4
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return b, 5
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        return a + ab * (d1 / (d1 - d3)), 1
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return c, 6
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        return a + ac * (d2 / (d2 - d6)), 3
    va = d3 * d6 - d5 * d4
    if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
        return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))), 2
    den = 1.0 / (va + vb + vc)
    return a + ab * (vb * den) + ac * (vc * den), 0


class _Tree:
    def __init__(self, mesh):
        self.m = mesh
        from scipy.spatial import cKDTree

        self.kd = cKDTree(mesh.v)
        inc = [[] for _ in range(len(mesh.v))]
        for fi, (u, v, w) in enumerate(mesh.f):
            inc[u].append(fi)
            inc[v].append(fi)
            inc[w].append(fi)
        self.inc = inc

    def nearest(self, pts, nearest_part=False):
        v, f = self.m.v, self.m.f
        k = min(6, len(v))
        _, nn = self.kd.query(pts, k=k)
        faces = np.zeros(len(pts), dtype=np.uint32)
        parts = np.zeros(len(pts), dtype=np.uint32)
        close = np.zeros((len(pts), 3))
        for i, p in enumerate(pts):
            best = (np.inf, 0, 0, None)
            cand = sorted({fi for u in np.atleast_1d(nn[i]) for fi in self.inc[u]})
            for fi in cand:
                q, code = _closest_on_triangle(p, v[f[fi, 0]], v[f[fi, 1]], v[f[fi, 2]])
                d = float(np.sum((p - q) ** 2))
                if d < best[0]:
                    best = (d, fi, code, q)
            faces[i], parts[i], close[i] = best[1], best[2], best[3]
        return faces.reshape(1, -1), parts.reshape(1, -1), close


class Mesh:
    def __init__(self, v=None, f=None, filename=None):
        self.v = np.asarray(v, dtype=np.float64)
        self.f = np.asarray(f, dtype=np.int64)

    def compute_aabb_tree(self):
        return _Tree(self)


def install():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    yacs = mod("yacs")
    yacs.config = mod("yacs.config", CfgNode=CfgNode)
    od = mod("opendr")
    od.topology = mod("opendr.topology", get_vert_connectivity=get_vert_connectivity,
                      get_vertices_per_edge=get_vertices_per_edge)
    ps = mod("psbody")
    ps.mesh = mod("psbody.mesh", Mesh=Mesh)
    mod("torch_scatter", scatter_add=None)
    tm = mod("trimesh")
    tm.base = mod("trimesh.base")
    tm.exchange = mod("trimesh.exchange")
    tm.exchange.export = mod("trimesh.exchange.export", export_mesh=None)
    mod("tensorboardX", SummaryWriter=object)
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
