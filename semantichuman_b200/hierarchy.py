"""Mesh hierarchy builder: QSlim decimation (D), closest-point barycentric up-sampling (U), adjacency (A), faces (F).

SURVEY 8 f-3.  What the reference gets from ``mesh_sampling.generate_transform_matrices`` (mesh_sampling.py:229-265, which
needs psbody + opendr) and caches as ``downsampling_matrices{f0}{f1}{f2}{f3}.pkl`` (main.py:93-116), built here from
numpy / scipy alone.  Setup-time host code: it produces the inputs of the hot path and runs once per template.

The decimation reproduces the reference's result EXACTLY (same kept vertices, same faces in the same order) because the
spirals, D and U of a trained checkpoint are only meaningful for that exact hierarchy.  That fixes the arithmetic
(per-face plane from the SVD null vector, quadrics accumulated in face order, costs as p^T Q p through the same numpy
products, mesh_sampling.py:20-45, 122-133) and the edge-queue semantics (a binary heap whose entries are relabelled in
place after every collapse, stale costs re-queued, mesh_sampling.py:136-190).  What differs is the data structure: the
reference rescans the whole queue and the whole face array after every collapse (O(E) and O(F) per collapse, 28 s for the
6890-vertex template); here every vertex keeps the queue entries and faces that mention it, so a collapse touches only its
neighbourhood (6-8 s for the same four levels, up-sampling matrices included).

The up-sampling matrix follows mesh_sampling.py:47-96: for every fine vertex the closest point on the coarse surface,
expressed in the vertices of the face it falls on -- inside the face: the exact 3x3 solve; on an edge: least squares of
the fine vertex on the two end points; at a vertex: 1.  The closest-point query (psbody's AABB tree in the reference) is
an exact point-triangle distance over the faces around the nearest coarse vertices.
"""
import heapq
import math
import pickle

import numpy as np
import scipy.sparse as sp


# ------------------------------------------------------------------------------------------------ topology helpers
def vertex_adjacency(n_verts, faces):
    """Symmetric 0/1 vertex adjacency (csc) from the three edges of every face (opendr.topology.get_vert_connectivity)."""
    f = np.asarray(faces, dtype=np.int64)
    rows = np.concatenate([f[:, 0], f[:, 1], f[:, 2], f[:, 1], f[:, 2], f[:, 0]])
    cols = np.concatenate([f[:, 1], f[:, 2], f[:, 0], f[:, 0], f[:, 1], f[:, 2]])
    a = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_verts, n_verts))
    a.data[:] = 1.0
    return a


# ------------------------------------------------------------------------------------------------ QSlim decimation
def vertex_quadrics(verts, faces):
    """(V, 4, 4) sum over incident faces of (plane)(plane)^T, plane = unit-normal plane equation of the face
    (mesh_sampling.py:20-45).  The plane is the null vector of [v0 1; v1 1; v2 1] from numpy's SVD, as in the reference."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    tri = np.concatenate([verts[faces], np.ones((len(faces), 3, 1))], axis=2)  # (F, 3, 4)
    planes = np.linalg.svd(tri)[2][:, -1, :]                                     # (F, 4)
    q = np.zeros((len(verts), 4, 4))
    outer = np.empty((len(faces), 4, 4))
    for i, eq in enumerate(planes):
        eq = eq.reshape(-1, 1)
        eq = eq / np.linalg.norm(eq[0:3])
        outer[i] = np.outer(eq, eq)
    # face-major, corner-minor accumulation order: every vertex adds its faces in face order
    np.add.at(q, faces.reshape(-1), np.repeat(outer, 3, axis=0))
    return q


def _quadric_costs(q, r, c, verts):
    qsum = q[r] + q[c]
    p1 = np.vstack((verts[r].reshape(-1, 1), np.array([1]).reshape(-1, 1)))
    p2 = np.vstack((verts[c].reshape(-1, 1), np.array([1]).reshape(-1, 1)))
    destroy_c = float(p1.T.dot(qsum).dot(p1)[0, 0])   # keep r: the merged quadric evaluated at r
    destroy_r = float(p2.T.dot(qsum).dot(p2)[0, 0])
    return destroy_c, destroy_r, qsum


def decimate(verts, faces, factor=None, n_verts_desired=None):
    """QSlim edge collapse without vertex re-positioning (mesh_sampling.py:99-209).  Returns (new_faces, kept):
    `kept` = ascending ids of the surviving vertices, `new_faces` = the surviving faces, in their original order, in the
    new numbering.  The down-sampling matrix is the row selection `kept` (selection_matrix)."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.array(faces, dtype=np.int64)
    n = len(verts)
    if factor is None and n_verts_desired is None:
        raise ValueError("need either factor or n_verts_desired")
    if n_verts_desired is None:
        n_verts_desired = math.ceil(n * factor)
    q = vertex_quadrics(verts, faces)

    # edge queue: entries are mutable [cost, r, c] lists (ordered like the reference's (cost, (r, c)) tuples); `mentions[v]`
    # lists the entries whose r or c is v, so that a collapse relabels them without scanning the queue
    adj = vertex_adjacency(n, faces).tocsc()
    adj.sort_indices()
    queue, mentions = [], [[] for _ in range(n)]

    def push(cost, r, c):
        e = [cost, r, c]
        heapq.heappush(queue, e)
        mentions[r].append(e)
        mentions[c].append(e)

    for c in range(n):  # column-major, rows ascending: the order the reference's coo iteration visits r <= c
        for r in adj.indices[adj.indptr[c]:adj.indptr[c + 1]]:
            if r > c:
                continue
            dc, dr, _ = _quadric_costs(q, r, c, verts)
            push(min(dc, dr), int(r), c)

    # faces: alive mask + per-vertex incident faces + per-vertex alive-face counts (== membership in unique(faces))
    alive = np.ones(len(faces), dtype=bool)
    incident = [[] for _ in range(n)]
    for fi, tri in enumerate(faces):
        for v in tri:
            incident[v].append(fi)
    count = np.array([len(x) for x in incident])
    n_total = n  # the reference starts from len(mesh.v) and recounts from the faces after every collapse

    while n_total > n_verts_desired:
        e = heapq.heappop(queue)
        cost_then, r, c = e
        e[1] = e[2] = -1  # no longer in the queue: later relabelling must not resurrect it
        if r == c:
            continue
        dc, dr, qsum = _quadric_costs(q, r, c, verts)
        cost_now = min(dc, dr)
        if cost_now > cost_then:  # stale entry: re-queue with the current cost
            push(cost_now, r, c)
            continue
        to_destroy, to_keep = (c, r) if dc < dr else (r, c)
        # relabel the queue entries that mention the destroyed vertex
        for m in mentions[to_destroy]:
            hit = False
            if m[1] == to_destroy:
                m[1], hit = to_keep, True
            if m[2] == to_destroy:
                m[2], hit = to_keep, True
            if hit:
                mentions[to_keep].append(m)
        mentions[to_destroy] = []
        q[r] = qsum
        q[c] = qsum
        # relabel its faces; drop the ones that became degenerate
        for fi in incident[to_destroy]:
            if not alive[fi]:
                continue
            tri = faces[fi]
            tri[tri == to_destroy] = to_keep
            count[to_destroy] -= 1
            if tri[0] == tri[1] or tri[1] == tri[2] or tri[2] == tri[0]:
                alive[fi] = False  # it held to_keep already: the two distinct vertices left each lose this face
                for v in set(tri.tolist()):
                    count[v] -= 1
            else:
                incident[to_keep].append(fi)
                count[to_keep] += 1
        incident[to_destroy] = []
        n_total = int(np.count_nonzero(count > 0))

    left = faces[alive]
    kept = np.unique(left.reshape(-1))
    remap = np.arange(0, left.max() + 1)
    remap[kept] = np.arange(len(kept))
    return remap[left.reshape(-1)].reshape(-1, 3), kept


def selection_matrix(kept, n_verts):
    """Down-sampling transform of a decimation: one 1.0 per row (mesh_sampling.py:212-227)."""
    kept = np.asarray(kept)
    return sp.csc_matrix((np.ones(len(kept)), (np.arange(len(kept)), kept)), shape=(len(kept), n_verts))


# ------------------------------------------------------------------------------------------------ closest point / U
def closest_point_on_triangle(p, a, b, c):
    """Closest point of triangle abc to p and where it lies: 0 inside the face, 1..3 on edge (k-1, k%3), 4..6 at
    vertex k-4 (the part codes of psbody's nearest(), as mesh_sampling.py:72-85 reads them)."""
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


def nearest_on_surface(src_verts, src_faces, points, k_nearest=6):
    """For every point: (face id, part code, closest point) on the source surface; candidates are the faces around the
    `k_nearest` nearest source vertices, in ascending face order, first minimum wins."""
    from scipy.spatial import cKDTree

    src_verts = np.asarray(src_verts, dtype=np.float64)
    src_faces = np.asarray(src_faces, dtype=np.int64)
    around = [[] for _ in range(len(src_verts))]
    for fi, tri in enumerate(src_faces):
        for v in tri:
            around[v].append(fi)
    _, nn = cKDTree(src_verts).query(points, k=min(k_nearest, len(src_verts)))
    face = np.zeros(len(points), dtype=np.int64)
    code = np.zeros(len(points), dtype=np.int64)
    close = np.zeros((len(points), 3))
    for i, p in enumerate(points):
        best = np.inf
        for fi in sorted({f for v in np.atleast_1d(nn[i]) for f in around[v]}):
            tri = src_faces[fi]
            qpt, part = closest_point_on_triangle(p, src_verts[tri[0]], src_verts[tri[1]], src_verts[tri[2]])
            d = float(np.sum((p - qpt) ** 2))
            if d < best:
                best, face[i], code[i], close[i] = d, fi, part, qpt
    return face, code, close


def upsampling_matrix(coarse_verts, coarse_faces, fine_verts):
    """(n_fine, n_coarse) csc: every fine vertex as a combination of the vertices of its closest coarse face
    (mesh_sampling.py:47-96); three stored entries per row (zeros included, as in the reference)."""
    coarse_verts = np.asarray(coarse_verts, dtype=np.float64)
    fine_verts = np.asarray(fine_verts, dtype=np.float64)
    n = len(fine_verts)
    face, code, close = nearest_on_surface(coarse_verts, coarse_faces, fine_verts)
    rows, cols, coef = np.zeros(3 * n), np.zeros(3 * n), np.zeros(3 * n)
    for i in range(n):
        tri = np.asarray(coarse_faces)[face[i]]
        rows[3 * i:3 * i + 3] = i
        cols[3 * i:3 * i + 3] = tri
        part = code[i]
        if part == 0:    # inside the face: point = c0 v0 + c1 v1 + c2 v2
            coef[3 * i:3 * i + 3] = np.linalg.lstsq(np.vstack(coarse_verts[tri]).T, close[i])[0]
        elif part <= 3:  # on an edge: the fine vertex itself, least squares on the two end points
            a = np.vstack((coarse_verts[tri[part - 1]], coarse_verts[tri[part % 3]])).T
            c2 = np.linalg.lstsq(a, fine_verts[i])[0]
            coef[3 * i + part - 1] = c2[0]
            coef[3 * i + part % 3] = c2[1]
        else:            # at a vertex
            coef[3 * i + part - 4] = 1.0
    return sp.csc_matrix((coef, (rows, cols)), shape=(n, len(coarse_verts)))


# ------------------------------------------------------------------------------------------------ the whole hierarchy
def build_hierarchy(verts, faces, factors):
    """generate_transform_matrices(mesh, factors) (mesh_sampling.py:229-265) -> dict with the keys of the reference's
    cache: 'M_verts_faces' [(verts, faces)] per level, 'A' adjacency, 'D' down-sampling, 'U' up-sampling, 'F' faces."""
    verts = np.asarray(verts, dtype=np.float64)
    faces = np.asarray(faces, dtype=np.int64)
    levels = [(verts, faces)]
    A, D, U, F = [vertex_adjacency(len(verts), faces)], [], [], []
    for factor in factors:
        v, f = levels[-1]
        new_f, kept = decimate(v, f, factor=1.0 / factor)
        d = selection_matrix(kept, len(v))
        new_v = d.dot(v)
        D.append(d)
        F.append(new_f)
        levels.append((new_v, new_f))
        A.append(vertex_adjacency(len(new_v), new_f))
        U.append(upsampling_matrix(new_v, new_f, v))
    return {"M_verts_faces": levels, "A": A, "D": D, "U": U, "F": F}


def cache_name(factors):
    """main.py:93: downsampling_matrices{f0}{f1}{f2}{f3}.pkl"""
    return "downsampling_matrices" + "".join(str(f) for f in factors) + ".pkl"


def save_cache(path, hierarchy):
    with open(path, "wb") as fh:
        pickle.dump({k: hierarchy[k] for k in ("M_verts_faces", "A", "D", "U", "F")}, fh)


def load_cache(path):
    with open(path, "rb") as fh:
        h = pickle.load(fh)
    missing = [k for k in ("M_verts_faces", "A", "D", "U", "F") if k not in h]
    if missing:
        raise ValueError(f"not a hierarchy cache (missing {missing})")
    return h
