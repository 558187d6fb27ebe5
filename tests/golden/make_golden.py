#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run once in the build container (``/root/reference`` mounted read-only):

    python tests/golden/make_golden.py [--only hier|ops|ae|pair|big]

It imports the reference's ``models.py``, ``utils_spiral.py`` and ``mesh_sampling.py`` in place and
unmodified (``_ref_stubs.install()`` supplies inert stand-ins for yacs / psbody / opendr / ...), runs them
on seeded synthetic inputs, and stores inputs + outputs as small ``.npz`` files.  The GPU box has no
``/root/reference``; tests there compare the CUDA path and the oracle against these files.

Fixtures
  hier_2222.npz / hier_4444.npz   6890-vertex template, reference QSlim hierarchy (mesh_sampling.py:229-265),
                                  reference spirals (utils_spiral.py:45-95) for the default config
                                  (step 2,2,1,1,1 / dilation 2,2,1,1,1) and the 1-hop undilated config.
  hier_small.npz / hier_open.npz  300-vertex closed and 144-vertex open meshes, same content.
  golden_ops.npz                  reference SpiralConv fwd/bwd for every activation; dense-pool fwd/bwd.
  golden_ae_small*.npz            reference SpiralAutoencoder fwd + l1 + bwd, all parameters stored.
  golden_multiz_small.npz         reference SpiralAutoencoder_multiz_partkps fwd + losses + bwd.
  golden_pair_loss.npz            reference angle_skl + calc_euclidean_dist_matrix + the pairwise-distance loss lines
                                  of train_funcs.py:243-284 (three weight/normalisation configurations) + d/d rec.
  golden_aux_losses.npz           reference edge regulariser / Edge_loss / cal_volloss (train_funcs.py:22-72), looped over
                                  the batch as the training loop does.
  golden_skeleton.npz             reference kps2skl / skl2kps (utils_SH.py:26-80), all modes.
  golden_ae_6890.npz              full-size SpiralAutoencoder (default filters, nz=256), B=2, deterministic
                                  weights (tests.golden.synthetic.fill_deterministic_), outputs and
                                  gradient samples.
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

import _ref_stubs  # noqa: E402

_ref_stubs.install()
import mesh_sampling  # noqa: E402  (reference)
import models as ref_models  # noqa: E402  (reference)
import utils_spiral  # noqa: E402  (reference)
from _ref_stubs import Mesh  # noqa: E402

from tests.golden.synthetic import (fill_deterministic_, make_open_template, make_template,  # noqa: E402
                                          synthetic_meshes)

DEFAULT_STEPS = [2, 2, 1, 1, 1]
DEFAULT_DIL = [2, 2, 1, 1, 1]


def build_hierarchy(verts, faces, factors, ref_vertex):
    """main.py:93-181 restated as a function call sequence over the reference's own builders."""
    from sklearn.metrics.pairwise import euclidean_distances

    M, A, D, U, F = mesh_sampling.generate_transform_matrices(Mesh(v=verts, f=faces), factors)
    refpts = [[ref_vertex]]
    for i in range(len(factors)):
        dist = euclidean_distances(M[i + 1].v, M[0].v[refpts[0]])
        refpts.append(np.argmin(dist, axis=0).tolist())
    Adj, Trigs = utils_spiral.get_adj_trigs(A, F, M[0], meshpackage="mpi-mesh")
    return M, A, D, U, F, refpts, Adj, Trigs


def ref_spirals(M, Adj, Trigs, refpts, steps, dil):
    n = len(M)
    sp_np, sizes, _ = utils_spiral.generate_spirals(steps[:n], M, Adj, Trigs, reference_points=refpts,
                                                    dilation=dil[:n], random=False, meshpackage="mpi-mesh",
                                                    counter_clockwise=True)
    return sp_np, sizes


def save_hier(tag, verts, faces, factors, ref_vertex):
    t0 = time.time()
    M, A, D, U, F, refpts, Adj, Trigs = build_hierarchy(verts, faces, factors, ref_vertex)
    out = {"verts0": verts.astype(np.float64), "faces0": faces.astype(np.int32),
           "factors": np.asarray(factors, np.int32), "refpts": np.asarray([r[0] for r in refpts], np.int32),
           "sizes": np.asarray([m.v.shape[0] for m in M], np.int32)}
    for l, f in enumerate(F):
        out[f"faces{l + 1}"] = np.asarray(f, np.int32)
    for l, d in enumerate(D):
        d = sp.csr_matrix(d)
        assert (np.diff(d.indptr) == 1).all() and (d.data == 1.0).all()
        out[f"D{l}_col"] = d.indices.astype(np.int32)
    for l, u in enumerate(U):
        u = sp.csr_matrix(u)  # keeps explicitly stored zeros, like .todense() sees them as 0
        out[f"U{l}_indptr"] = u.indptr.astype(np.int32)
        out[f"U{l}_indices"] = u.indices.astype(np.int32)
        out[f"U{l}_data"] = u.data.astype(np.float64)
    for cfg_tag, steps, dil in (("A", DEFAULT_STEPS, DEFAULT_DIL), ("B", [1] * 5, [1] * 5)):
        sp_np, sizes = ref_spirals(M, Adj, Trigs, refpts, steps, dil)
        out[f"sp{cfg_tag}_sizes"] = np.asarray(sizes, np.int32)
        for l, s in enumerate(sp_np):
            assert s.shape[0] == 1 and np.all(s == np.round(s)) and s.max() < 32767
            out[f"sp{cfg_tag}{l}"] = s[0].astype(np.int16)
    np.savez_compressed(os.path.join(HERE, f"hier_{tag}.npz"), **out)
    print(f"hier_{tag}: sizes {out['sizes'].tolist()} spA {out['spA_sizes'].tolist()} spB {out['spB_sizes'].tolist()}"
          f"  ({time.time() - t0:.1f}s)")


def load_hier_for_ref(tag, cfg_tag="A", device="cpu"):
    """Rebuild the reference constructor arguments (main.py:183-205) from a hier fixture."""
    h = np.load(os.path.join(HERE, f"hier_{tag}.npz"))
    sizes = h["sizes"].tolist()
    nl = len(sizes) - 1
    spirals = [torch.from_numpy(h[f"sp{cfg_tag}{l}"].astype(np.float64)[None]).long() for l in range(nl + 1)]
    bD, bU = [], []
    for l in range(nl):
        d = np.zeros((1, sizes[l + 1] + 1, sizes[l] + 1))
        d[0, np.arange(sizes[l + 1]), h[f"D{l}_col"]] = 1
        d[0, -1, -1] = 1
        u = np.zeros((1, sizes[l] + 1, sizes[l + 1] + 1))
        um = sp.csr_matrix((h[f"U{l}_data"], h[f"U{l}_indices"], h[f"U{l}_indptr"]), shape=(sizes[l], sizes[l + 1]))
        u[0, :-1, :-1] = um.todense()
        u[0, -1, -1] = 1
        bD.append(torch.from_numpy(d).float())
        bU.append(torch.from_numpy(u).float())
    return h, sizes, h[f"sp{cfg_tag}_sizes"].tolist(), spirals, bD, bU


def gen_ops():
    h, sizes, ssz, spirals, bD, bU = load_hier_for_ref("small")
    out = {}
    g = torch.Generator().manual_seed(11)
    acts = ["relu", "elu", "leaky_relu", "sigmoid", "tanh", "identity"]
    cases = [(0, 3, 16), (0, 16, 3), (1, 16, 32), (2, 32, 8), (3, 5, 7)]
    k = 0
    for (lvl, cin, cout) in cases:
        for act in (acts if (lvl, cin) == (0, 3) else ["elu", "identity"]):
            B = 3
            torch.manual_seed(100 + k)
            conv = ref_models.SpiralConv(cin, ssz[lvl], cout, activation=act, device="cpu")
            x = torch.randn(B, sizes[lvl] + 1, cin, generator=g)
            x[:, -1] = torch.randn(B, cin, generator=g) if k % 2 else 0.0  # live dummy row on odd cases
            x.requires_grad_(True)
            y = conv(x, spirals[lvl].repeat(B, 1, 1))
            gy = torch.randn(y.shape, generator=g)
            y.backward(gy)
            pre = f"conv{k}_"
            out[pre + "meta"] = np.asarray([lvl, cin, cout, ssz[lvl], B], np.int32)
            out[pre + "act"] = np.asarray(act)
            out[pre + "x"] = x.detach().numpy()
            out[pre + "w"] = conv.conv.weight.detach().numpy()
            out[pre + "b"] = conv.conv.bias.detach().numpy()
            out[pre + "y"] = y.detach().numpy()
            out[pre + "gy"] = gy.numpy()
            out[pre + "gx"] = x.grad.numpy()
            out[pre + "gw"] = conv.conv.weight.grad.numpy()
            out[pre + "gb"] = conv.conv.bias.grad.numpy()
            k += 1
    out["n_conv"] = np.asarray(k)
    # dense pools exactly as models.py:127,148 calls them
    k = 0
    for which, mats in (("D", bD), ("U", bU)):
        for l, P in enumerate(mats):
            C = [4, 16, 5, 32][l]
            x = torch.randn(2, P.shape[2], C, generator=g, requires_grad=True)
            y = torch.matmul(P, x)
            gy = torch.randn(y.shape, generator=g)
            y.backward(gy)
            pre = f"pool{k}_"
            out[pre + "which"] = np.asarray(f"{which}{l}")
            out[pre + "x"], out[pre + "y"] = x.detach().numpy(), y.detach().numpy()
            out[pre + "gy"], out[pre + "gx"] = gy.numpy(), x.grad.numpy()
            k += 1
    out["n_pool"] = np.asarray(k)
    np.savez_compressed(os.path.join(HERE, "golden_ops.npz"), **out)
    print("golden_ops:", int(out["n_conv"]), "conv cases,", k, "pool cases")


def _dump_model(out, model):
    for n, p in model.state_dict().items():
        out["p_" + n] = p.detach().numpy()
    for n, p in model.named_parameters():
        out["g_" + n] = p.grad.detach().numpy()


def gen_ae_small():
    for tag, fenc, fdec in (
        ("golden_ae_small", [[3, 8, 8, 16, 16], [[], [], [], [], []]], [[16, 16, 8, 8, 8], [[], [], [], [], 3]]),
        ("golden_ae_small2", [[3, 8, 8, 16, 16], [[], 8, [], 12, []]], [[16, 16, 8, 8, 3], [[], 8, [], [], []]]),
    ):
        h, sizes, ssz, spirals, bD, bU = load_hier_for_ref("small")
        torch.manual_seed(2)
        model = ref_models.SpiralAutoencoder(fenc, fdec, latent_size=12, sizes=sizes, spiral_sizes=ssz,
                                             spirals=spirals, D=bD, U=bU, device="cpu")
        x = synthetic_meshes(h["verts0"], 3, seed=5, noise=0.05)
        xh, z = model(x)
        loss = torch.nn.functional.l1_loss(x, xh)
        loss.backward()
        out = {"x": x.numpy(), "xhat": xh.detach().numpy(), "z": z.detach().numpy(), "loss": loss.detach().numpy(),
               "filters_enc0": np.asarray(fenc[0]), "filters_dec0": np.asarray(fdec[0]),
               "filters_enc1": np.asarray([v if v else 0 for v in fenc[1]]),
               "filters_dec1": np.asarray([v if v else 0 for v in fdec[1]]), "latent": np.asarray(12)}
        _dump_model(out, model)
        np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
        print(tag, "loss", float(loss))


PART_LIST = ['head', 'neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'left_feet', 'right_ham',
             'right_shank', 'right_feet', 'left_arm', 'left_forearm', 'left_hand', 'right_arm', 'right_forearm',
             'right_hand']  # configure/cfgs.py:36-38
NOLEAF = ['neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'right_ham', 'right_shank', 'left_arm',
          'left_forearm', 'right_arm', 'right_forearm']  # cfgs.py:40-41
MEASURE_PART_LIST = ['neck', 'chest', 'abdomen', 'hip', 'left_ham', 'left_shank', 'left_feet', 'right_ham',
                     'right_shank', 'right_feet', 'left_arm', 'left_forearm', 'left_hand', 'right_arm',
                     'right_forearm', 'right_hand']  # cfgs.py:42-44
KPS_INDEX_LIST = [[12, 25, 26, 27], [12, 11], [11, 8], [5, 0], [0, 1, 2], [1, 3], [3, 6], [6, 9, 28, 30], [2, 4],
                  [4, 7], [7, 10, 29, 31], [13, 15], [15, 17], [17, 19, 21, 23], [14, 16], [16, 18],
                  [18, 20, 22, 24]]  # configure/traincfg.yaml:56


def gen_multiz_small():
    h, sizes, ssz, spirals, bD, bU = load_hier_for_ref("small")
    fenc = [[3, 8, 8, 16, 16], [[], [], [], [], []]]
    fdec = [[16, 16, 8, 8, 8], [[], [], [], [], 3]]
    nc = sizes[-1]
    rng = np.random.default_rng(4)
    perm = rng.permutation(nc)
    cuts = np.sort(rng.choice(np.arange(1, nc), size=16, replace=False)) if nc > 17 else np.arange(1, 17)
    parts = np.split(perm, cuts)
    vdict = {name: np.sort(p).astype(np.int64) for name, p in zip(PART_LIST, parts)}
    torch.manual_seed(2)
    model = ref_models.SpiralAutoencoder_multiz_partkps(KPS_INDEX_LIST, vdict, fenc, fdec, latent_size=8,
                                                        part_kps_latent_size=8, sizes=sizes, spiral_sizes=ssz,
                                                        spirals=spirals, D=bD, U=bU, device="cpu")
    B = 3
    x = synthetic_meshes(h["verts0"], B, seed=6, noise=0.05)
    g = torch.Generator().manual_seed(8)
    kps = torch.randn(B, 32, 3, generator=g) * 0.3
    measure = torch.rand(B, 32, generator=g) * 0.8 + 0.2
    xh, z, zk = model(x, kps)
    rec = torch.nn.functional.l1_loss(x, xh)
    # train_funcs.py:145-152 (relat_flag=True), weight zpartreg_w = 1e-2 (traincfg.yaml:41)
    P = [PART_LIST.index(n) for n in NOLEAF]
    Q = [MEASURE_PART_LIST.index(n) for n in NOLEAF]
    m = torch.sqrt(torch.sum(z ** 2, dim=2))
    zreg_rel = torch.nn.functional.l1_loss(m[:, P] / measure[:, Q], torch.ones_like(measure[:, Q]))
    zreg_abs = torch.nn.functional.l1_loss(m[:, P], measure[:, Q])
    loss = rec + 1e-2 * zreg_rel
    loss.backward()
    out = {"x": x.numpy(), "kps": kps.numpy(), "measure": measure.numpy(), "xhat": xh.detach().numpy(),
           "z": z.detach().numpy(), "zkps": zk.detach().numpy(), "rec": rec.detach().numpy(),
           "zreg_rel": zreg_rel.detach().numpy(), "zreg_abs": zreg_abs.detach().numpy(),
           "loss": loss.detach().numpy(), "P": np.asarray(P), "Q": np.asarray(Q)}
    for name, p in vdict.items():
        out["part_" + name] = p
    _dump_model(out, model)
    np.savez_compressed(os.path.join(HERE, "golden_multiz_small.npz"), **out)
    print("golden_multiz_small loss", float(loss), "parts", [len(p) for p in parts])


def ref_pair_loss(tx, rec, kps, vdict, names, skl_list, w_mode, w_threshold, leaf_list, relat_flag, scale):
    """train_funcs.py:243-284 over the reference's own angle_skl (utils_SH.py:442) and calc_euclidean_dist_matrix
    (utils_distance.py:366); the loop body is inline in the reference's training loop, so it is restated here line by
    line (w_part_mode '1/K', cfgs.py:95)."""
    import utils_SH  # noqa: E402  (reference)
    from utils_distance import calc_euclidean_dist_matrix  # noqa: E402  (reference)

    F = torch.nn.functional
    angle_w = utils_SH.angle_skl(tx, kps, names, vdict, skl_list)                                        # :243
    loss = 0.0
    for i in range(len(names)):
        idx = vdict[names[i]]
        De = calc_euclidean_dist_matrix(tx[:, idx, :])                                                   # :246
        De_r = calc_euclidean_dist_matrix(rec[:, idx, :])                                                # :247
        if scale is not None:
            De = De * scale[:, i][:, None, None]                                                         # :248-249
        w_part = 1 / len(names)                                                                          # :252-253
        if w_mode == 'all_one' or i in leaf_list:                                                        # :259
            w = torch.ones_like(angle_w[i].squeeze(-1))
        elif w_mode == 'linear':
            w = angle_w[i].squeeze(-1).float() / 90
        elif w_mode == 'sin':
            w = torch.sin(angle_w[i].squeeze(-1).float() / 180 * torch.pi)
        elif w_mode == 'threshold':
            w = angle_w[i].squeeze(-1).float() / 90
            w = torch.where(w < w_threshold, torch.full_like(w, 0), w)
        for b in range(w.shape[0]):                                                                      # :268-269
            w[b, ...] = w[b, ...] - torch.diag_embed(torch.diag(w[b, ...]))
        nz = torch.where((w * De) != 0)                                                                  # :272
        if not relat_flag:
            loss = loss + w_part * F.l1_loss(w[nz] * De_r[nz].float(), w[nz] * De[nz])                   # :274
        else:
            loss = loss + w_part * F.l1_loss(w[nz] * De_r[nz].float() / De[nz], w[nz] * torch.ones_like(w[nz]))  # :276
    return loss


def gen_pair_loss():
    """golden_pair_loss.npz: the orientation-adaptive pairwise-distance loss on the 300-vertex
    template, 5 parts, three configurations."""
    h = np.load(os.path.join(HERE, "hier_small.npz"))
    verts = h["verts0"]
    V = verts.shape[0]
    rng = np.random.default_rng(11)
    order = np.argsort(verts[:, 1], kind="stable")
    sizes = [37, 80, 61, 90, V - 268]
    parts = np.split(order, np.cumsum(sizes)[:-1])
    names = ["p%d" % i for i in range(5)]
    vdict = {n: torch.from_numpy(np.sort(p).astype(np.int64)) for n, p in zip(names, parts)}
    skl_list = [[3, 1], [0, 2], [4, 1, 5], [2, 6], [7, 3, 0]]
    B = 3
    tx = synthetic_meshes(verts, B, seed=21, noise=0.02)[:, :-1, :].contiguous()
    g = torch.Generator().manual_seed(9)
    kps = torch.randn(B, 8, 3, generator=g) * 0.4
    scale = torch.rand(B, 5, generator=g) * 0.4 + 0.8
    out = {"tx": tx.numpy(), "kps": kps.numpy(), "scale": scale.numpy(), "skl": np.asarray([b + [-1] * (3 - len(b)) for b in skl_list]),
           "part_sizes": np.asarray([len(p) for p in parts]), "part_idx": np.concatenate([vdict[n].numpy() for n in names])}
    configs = {"lin_rel_leaf": ("linear", 0.8, [0, 4], True, None),
               "thr_abs": ("threshold", 0.8, [], False, None),
               "sin_rel_scale": ("sin", 0.8, [], True, scale)}
    for tag, (mode, thr, leaf, rel, sc) in configs.items():
        rec = (tx + 0.03 * torch.randn(tx.shape, generator=g)).requires_grad_(True)
        loss = ref_pair_loss(tx, rec, kps, vdict, names, skl_list, mode, thr, leaf, rel, sc)
        loss.backward()
        out[tag + "_rec"] = rec.detach().numpy()
        out[tag + "_loss"] = loss.detach().numpy()
        out[tag + "_grec"] = rec.grad.numpy()
        print("golden_pair_loss", tag, float(loss), float(rec.grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "golden_pair_loss.npz"), **out)


def gen_skeleton():
    """golden_skeleton.npz: the reference's kps2skl / skl2kps (utils_SH.py:26-80) on random keypoints, every mode, for the
    full (31) and the kept (28) keypoint sets."""
    import utils_SH  # noqa: E402  (reference)

    g = torch.Generator().manual_seed(5)
    kps_full = torch.randn(4, 31, 3, generator=g)
    kps_keep = torch.randn(4, 28, 3, generator=g)
    out = {"kps_full": kps_full.numpy(), "kps_keep": kps_keep.numpy()}
    for mode in ["ori_m", "vec_m", "vec", "m"]:
        out["skl_full_" + mode] = utils_SH.kps2skl(kps_full, mode).numpy()
        out["skl_keep_" + mode] = utils_SH.kps2skl(kps_keep, mode).numpy()
    for mode in ["ori_m", "vec_m", "vec"]:
        out["back_" + mode] = utils_SH.skl2kps(torch.from_numpy(out["skl_keep_" + mode]), mode).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_skeleton.npz"), **out)
    print("golden_skeleton", {k: v.shape for k, v in out.items()})


def gen_aux_losses():
    """golden_aux_losses.npz: the reference's per-sample edge regulariser (get_target + compute_score, train_funcs.py:22-40,
    looped over the batch as in :137-143), Edge_loss (:42-45) and part-volume loss (cal_volloss :56-72, looped as in
    :325-333 with the face->part table of :84-89) on the 300-vertex template; values and d/d rec."""
    import train_funcs  # noqa: E402  (reference)

    h = np.load(os.path.join(HERE, "hier_small.npz"))
    verts, faces = h["verts0"], h["faces0"].astype(np.int32)
    V = verts.shape[0]
    B = 3
    g = torch.Generator().manual_seed(13)
    tx = synthetic_meshes(verts, B, seed=31, noise=0.02)[:, :-1, :].contiguous()
    order = np.argsort(verts[:, 1], kind="stable")
    parts = np.array_split(order, 4)
    vdict = {"p%d" % i: np.sort(p).astype(np.int64) for i, p in enumerate(parts)}
    out = {"tx": tx.numpy(), "faces": faces, "part_sizes": np.asarray([len(p) for p in parts]),
           "part_idx": np.concatenate([vdict["p%d" % i] for i in range(4)]), "parts_used": np.asarray([0, 2, 3])}
    # edge regulariser, train_funcs.py:137-143
    rec = (tx + 0.02 * torch.randn(tx.shape, generator=g)).requires_grad_(True)
    loss = 0
    for i in range(B):
        loss = loss + train_funcs.compute_score(rec[i].unsqueeze(0), faces, train_funcs.get_target(tx[i].numpy(), faces, 1, "cpu"))
    loss = loss / B
    loss.backward()
    out.update(edge_rec=rec.detach().numpy(), edge_loss=loss.detach().numpy(), edge_grec=rec.grad.numpy())
    # Edge_loss on the unique edges
    e = np.unique(np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [0, 2]]]), axis=1), axis=0)
    rec = (tx + 0.02 * torch.randn(tx.shape, generator=g)).requires_grad_(True)
    loss = train_funcs.Edge_loss(tx, rec, torch.from_numpy(e).long())
    loss.backward()
    out.update(edges=e, elen_rec=rec.detach().numpy(), elen_loss=loss.detach().numpy(), elen_grec=rec.grad.numpy())
    # part volumes, train_funcs.py:82-89 tables + :325-333 loop
    ft = torch.from_numpy(faces)
    vpi = torch.ones(V)
    for k, v in enumerate(vdict.values()):
        vpi[v] = k
    fpi = torch.ones(len(faces))
    for k, f in enumerate(ft):
        fpi[k] = vpi[f[0]] if (vpi[f[0]] == vpi[f[1]] and vpi[f[0]] == vpi[f[2]]) else 100
    rec = (tx + 0.02 * torch.randn(tx.shape, generator=g)).requires_grad_(True)
    used = [0, 2, 3]
    loss = 0
    for i in range(B):
        loss = loss + train_funcs.cal_volloss(rec[i], tx[i], ft, vpi, fpi, vdict, used)
    loss = loss / B
    loss.backward()
    out.update(vol_rec=rec.detach().numpy(), vol_loss=loss.detach().numpy(), vol_grec=rec.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "golden_aux_losses.npz"), **out)
    print("golden_aux_losses", float(out["edge_loss"]), float(out["elen_loss"]), float(out["vol_loss"]))


def gen_big():
    h, sizes, ssz, spirals, bD, bU = load_hier_for_ref("2222")
    fenc = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
    fdec = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
    model = ref_models.SpiralAutoencoder(fenc, fdec, latent_size=256, sizes=sizes, spiral_sizes=ssz,
                                         spirals=spirals, D=bD, U=bU, device="cpu")
    fill_deterministic_(model, seed=2)
    x = synthetic_meshes(h["verts0"], 2, seed=0, noise=0.01)
    xh, z = model(x)
    loss = torch.nn.functional.l1_loss(x, xh)
    loss.backward()
    out = {"xhat": xh.detach().numpy(), "z": z.detach().numpy(), "loss": loss.detach().numpy(),
           "n_params": np.asarray(sum(p.numel() for p in model.parameters()))}
    for n, p in model.named_parameters():
        gflat = p.grad.detach().reshape(-1)
        out["gsum_" + n] = gflat.double().sum().numpy()
        out["gabs_" + n] = gflat.double().abs().sum().numpy()
        stride = max(1, gflat.numel() // 4096)
        out["gsmp_" + n] = gflat[::stride][:4096].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_ae_6890.npz"), **out)
    print("golden_ae_6890 loss", float(loss), "params", int(out["n_params"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    a = ap.parse_args()
    torch.set_num_threads(8)
    if a.only in ("all", "hier"):
        v, f = make_template(300, seed=1, scale=(0.3, 0.9, 0.2))
        save_hier("small", v, f, [2, 2, 2, 2], ref_vertex=14)
        v, f = make_open_template(12)
        save_hier("open", v, f, [2, 2], ref_vertex=40)
        v, f = make_template(6890, seed=0)
        save_hier("2222", v, f, [2, 2, 2, 2], ref_vertex=414)
        save_hier("4444", v, f, [4, 4, 4, 4], ref_vertex=414)
    if a.only in ("all", "ops"):
        gen_ops()
    if a.only in ("all", "ae"):
        gen_ae_small()
        gen_multiz_small()
    if a.only in ("all", "pair"):
        gen_pair_loss()
        gen_skeleton()
        gen_aux_losses()
    if a.only in ("all", "big"):
        gen_big()


if __name__ == "__main__":
    main()
