"""BASELINE.json configs[4] stress case: the 27 554-vertex template = one midpoint subdivision of the 6890-vertex template
(V + E = 6890 + 20664), with its 4-level hierarchy built by semantichuman_b200/hierarchy.py (QSlim decimation + closest-point
up-sampling, the reference's mesh_sampling.py:229-265) and its spirals by semantichuman_b200/spirals.py (utils_spiral.py).
Writes tests/golden/hier_27554.npz in the format of the other hierarchy fixtures (tests/golden/loader.py).

    python scripts/build_stress_template.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from semantichuman_b200 import hierarchy as hy, spirals as spr
from tests.golden.loader import Hierarchy

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "hier_27554.npz")


def subdivide(v, f):
    """One midpoint subdivision: every edge gets a vertex, every triangle becomes four (orientation kept)."""
    f = np.asarray(f, np.int64)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    ue, inv = np.unique(e, axis=0, return_inverse=True)
    mid = len(v) + inv.reshape(3, -1).T            # (F, 3): midpoints of edges 01, 12, 20
    nv = np.concatenate([v, 0.5 * (v[ue[:, 0]] + v[ue[:, 1]])])
    a, b, c = f[:, 0], f[:, 1], f[:, 2]
    ab, bc, ca = mid[:, 0], mid[:, 1], mid[:, 2]
    nf = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)])
    return nv, nf


def main():
    h0 = Hierarchy("2222")
    v, f = subdivide(np.asarray(h0.verts0, np.float64), h0.faces[0])
    v = v / np.linalg.norm(v / np.array([0.3, 0.9, 0.2]), axis=1, keepdims=True)  # back onto the template's ellipsoid
    assert len(v) == 27554 and len(f) == 4 * 13776
    t0 = time.time()
    factors = [2, 2, 2, 2]
    H = hy.build_hierarchy(v, f.astype(np.int32), factors)
    verts = [np.asarray(vf[0]) for vf in H["M_verts_faces"]]
    faces = [np.asarray(vf[1]) for vf in H["M_verts_faces"]]
    print(f"hierarchy: {time.time() - t0:.1f} s; sizes {[len(m) for m in verts]}")
    refpts = [[int(h0.refpts[0])]]
    for l in range(1, len(verts)):
        refpts.append([int(np.argmin(np.linalg.norm(verts[l] - verts[0][refpts[0][0]], axis=1)))])
    out = {"verts0": verts[0], "faces0": faces[0].astype(np.int32), "factors": np.asarray(factors, np.int32),
           "refpts": np.asarray([r[0] for r in refpts], np.int32), "sizes": np.asarray([len(m) for m in verts], np.int32)}
    for l in range(len(factors)):
        out[f"faces{l + 1}"] = faces[l + 1].astype(np.int32)
        D = H["D"][l].tocsr()
        out[f"D{l}_col"] = D.indices.astype(np.int32)
        U = H["U"][l].tocsr()
        out[f"U{l}_indptr"], out[f"U{l}_indices"], out[f"U{l}_data"] = U.indptr.astype(np.int32), U.indices.astype(np.int32), U.data
    for tag, steps, dil in (("A", [2, 2, 1, 1, 1], [2, 2, 1, 1, 1]),):
        t0 = time.time()
        tables, sizes, _ = spr.generate_spirals(steps, verts, faces, refpts, dilation=dil)
        print(f"spirals {tag}: {time.time() - t0:.1f} s; lengths {sizes}")
        out[f"sp{tag}_sizes"] = np.asarray(sizes, np.int32)
        for l, t in enumerate(tables):
            out[f"sp{tag}{l}"] = t[0].astype(np.int32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
