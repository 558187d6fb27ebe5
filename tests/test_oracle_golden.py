"""Pin the CPU oracle against golden vectors produced by the reference itself (tests/golden/make_golden.py).

The oracle is the checker of every GPU parity test, so it is checked first, against everything the reference
generated: op-level SpiralConv / dense-pool fwd+bwd, two small SpiralAutoencoder configs, the bone-guided model with
its losses, and the full-size 6890-vertex autoencoder."""
import numpy as np
import pytest
import torch

from tests.helpers import (DEFAULT_FDEC, DEFAULT_FENC, filters_from_golden, golden, grads_from_golden, params_from_golden,
                     ref_args, relerr)
from oracle import spiral_oracle as so
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes

TIGHT = 5e-6  # the oracle and the golden run the same ATen ops; thread-order noise of the CPU reductions reaches ~2e-6


def test_spiral_conv_cases():
    g = golden("golden_ops")
    _, sizes, ssz, spirals, _, _ = ref_args("small")
    for k in range(int(g["n_conv"])):
        pre = f"conv{k}_"
        lvl, cin, cout, S, B = g[pre + "meta"].tolist()
        x = torch.from_numpy(g[pre + "x"]).requires_grad_(True)
        w = torch.from_numpy(g[pre + "w"]).requires_grad_(True)
        b = torch.from_numpy(g[pre + "b"]).requires_grad_(True)
        y = so.spiral_conv(x, spirals[lvl], w, b, str(g[pre + "act"]))
        y.backward(torch.from_numpy(g[pre + "gy"]))
        assert relerr(y, g[pre + "y"]) < TIGHT, k
        assert relerr(x.grad, g[pre + "gx"]) < TIGHT, k
        assert relerr(w.grad, g[pre + "gw"]) < 5e-6, k
        assert relerr(b.grad, g[pre + "gb"]) < 5e-6, k


def test_unknown_activation_raises():
    with pytest.raises(NotImplementedError):
        so.spiral_conv(torch.zeros(1, 3, 2), torch.zeros(3, 2, dtype=torch.long), torch.zeros(1, 4), None, "gelu")


def test_dense_pool_cases():
    g = golden("golden_ops")
    _, _, _, _, D, U = ref_args("small")
    mats = {f"D{l}": m for l, m in enumerate(D)} | {f"U{l}": m for l, m in enumerate(U)}
    for k in range(int(g["n_pool"])):
        pre = f"pool{k}_"
        x = torch.from_numpy(g[pre + "x"]).requires_grad_(True)
        y = so.pool(mats[str(g[pre + "which"])], x)
        y.backward(torch.from_numpy(g[pre + "gy"]))
        assert relerr(y, g[pre + "y"]) < TIGHT and relerr(x.grad, g[pre + "gx"]) < TIGHT


@pytest.mark.parametrize("name", ["golden_ae_small", "golden_ae_small2"])
def test_autoencoder_small(name):
    g = golden(name)
    _, sizes, ssz, spirals, D, U = ref_args("small")
    fe, fd = filters_from_golden(g)
    params = {k: v.requires_grad_(True) for k, v in params_from_golden(g).items()}
    x = torch.from_numpy(g["x"])
    xh, z = so.autoencoder_forward(params, x, fe, fd, sizes, spirals, D, U)
    loss = so.l1_loss(x, xh)
    loss.backward()
    assert relerr(xh, g["xhat"]) < 1e-5 and relerr(z, g["z"]) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for n, gr in grads_from_golden(g).items():
        assert relerr(params[n].grad, gr) < 2e-5, n


def test_multiz_small():
    g = golden("golden_multiz_small")
    from tests.golden.constants import KPS_INDEX_LIST, PART_LIST  # noqa: F401

    _, sizes, ssz, spirals, D, U = ref_args("small")
    fe = [[3, 8, 8, 16, 16], [[], [], [], [], []]]
    fd = [[16, 16, 8, 8, 8], [[], [], [], [], 3]]
    parts = [g["part_" + n] for n in PART_LIST]
    params = {k: v.requires_grad_(True) for k, v in params_from_golden(g).items()}
    x, kps, measure = (torch.from_numpy(g[k]) for k in ("x", "kps", "measure"))
    xh, z, zk = so.multiz_forward(params, x, kps, KPS_INDEX_LIST, parts, fe, fd, sizes, spirals, D, U)
    rec = so.l1_loss(x, xh)
    zr = so.zpart_reg(z, measure, g["P"].tolist(), g["Q"].tolist(), relative=True)
    za = so.zpart_reg(z, measure, g["P"].tolist(), g["Q"].tolist(), relative=False)
    (rec + 1e-2 * zr).backward()
    assert relerr(xh, g["xhat"]) < 1e-5 and relerr(z, g["z"]) < 1e-5 and relerr(zk, g["zkps"]) < 1e-5
    assert abs(float(zr) - float(g["zreg_rel"])) < 1e-6 and abs(float(za) - float(g["zreg_abs"])) < 1e-6
    for n, gr in grads_from_golden(g).items():
        assert relerr(params[n].grad, gr) < 2e-5, n


def test_autoencoder_6890_full_size():
    """Full default config (28.56 M parameters) on the 6890-vertex template, B=2."""
    g = golden("golden_ae_6890")
    h, sizes, ssz, spirals, D, U = ref_args("2222")

    class _P(torch.nn.Module):  # parameter container with the reference's state_dict keys and order
        def __init__(self):
            super().__init__()
            enc, dec = so.conv_plan(DEFAULT_FENC, DEFAULT_FDEC, 4)
            self.conv = torch.nn.ModuleList(_L(ssz[l] * ci, co) for (l, ci, co, _) in enc)
            self.fc_latent_enc = torch.nn.Linear((sizes[-1] + 1) * 128, 256)
            self.fc_latent_dec = torch.nn.Linear(256, (sizes[-1] + 1) * 128)
            self.dconv = torch.nn.ModuleList(_L(ssz[l] * ci, co) for (l, ci, co, _) in dec)

    class _L(torch.nn.Module):
        def __init__(self, k, n):
            super().__init__()
            self.conv = torch.nn.Linear(k, n)

    m = fill_deterministic_(_P(), seed=2)
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"]) == 28559811
    params = dict(m.named_parameters())
    x = synthetic_meshes(h.verts0, 2, seed=0, noise=0.01)
    xh, z = so.autoencoder_forward(params, x, DEFAULT_FENC, DEFAULT_FDEC, sizes, spirals, D, U)
    loss = so.l1_loss(x, xh)
    loss.backward()
    assert relerr(xh, g["xhat"]) < 1e-5 and relerr(z, g["z"]) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for n, p in params.items():
        flat = p.grad.reshape(-1)
        stride = max(1, flat.numel() // 4096)
        assert relerr(flat[::stride][:4096], g["gsmp_" + n]) < 5e-5, n
        assert abs(float(flat.double().abs().sum()) - float(g["gabs_" + n])) <= 1e-4 * float(g["gabs_" + n]) + 1e-12, n


PAIR_CONFIGS = {"lin_rel_leaf": ("linear", 0.8, (0, 4), True, False), "thr_abs": ("threshold", 0.8, (), False, False),
                "sin_rel_scale": ("sin", 0.8, (), True, True)}


def pair_inputs(g):
    sizes = [int(v) for v in g["part_sizes"]]
    parts = [torch.from_numpy(p.astype("int64")) for p in np.split(g["part_idx"], np.cumsum(sizes)[:-1])]
    skl = [[int(v) for v in row if v >= 0] for row in g["skl"]]
    return parts, skl


@pytest.mark.parametrize("tag", list(PAIR_CONFIGS))
def test_pair_distance_loss_oracle_matches_reference(tag):
    """Orientation-adaptive pairwise-distance loss (utils_SH.py:442-478 + train_funcs.py:243-284): the oracle's
    restatement against numbers produced by the reference's own angle_skl / calc_euclidean_dist_matrix."""
    g = golden("golden_pair_loss")
    parts, skl = pair_inputs(g)
    mode, thr, leaf, rel, use_scale = PAIR_CONFIGS[tag]
    tx, kps = torch.from_numpy(g["tx"]), torch.from_numpy(g["kps"])
    rec = torch.from_numpy(g[tag + "_rec"]).requires_grad_(True)
    loss = so.pair_distance_loss(tx, rec, kps, parts, skl, w_mode=mode, w_threshold=thr, leaf_parts=leaf, relative=rel,
                                 scale=torch.from_numpy(g["scale"]) if use_scale else None)
    loss.backward()
    assert abs(loss.item() - float(g[tag + "_loss"])) < 2e-6
    assert relerr(rec.grad, g[tag + "_grec"]) < 2e-5
