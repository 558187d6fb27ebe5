"""GPU parity of the slab-layout operators (semantichuman_b200/slab.py -> shb_slab_* through the C ABI) against the CPU
oracle and the reference-generated goldens: layout round trips, SpiralConv forward + all three gradients in both
precision modes, Pool, and the persistent multi-tile path at the benchmark's batch size.

Tolerances are north_star's: per-tensor max|a-b|/max|b| <= 1e-4 in fp32 mode (planes = 2), <= 2e-2 in bf16 mode."""
import numpy as np
import pytest
import torch

from oracle import spiral_oracle as so
from tests.helpers import TOL_BF16, TOL_F32, golden, ref_args, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def sl():
    from semantichuman_b200 import slab

    return slab


@pytest.fixture(scope="module")
def shb():
    import semantichuman_b200 as m

    return m


def _bf16(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("B", [1, 3, 130, 256])
@pytest.mark.parametrize("C", [3, 8, 16, 24, 32, 64, 128, 256])
@pytest.mark.parametrize("planes", [1, 2])
def test_rows_slab_round_trip(sl, B, C, planes):
    R = 37 if C < 256 else 5
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, R, C, generator=g)
    perm = torch.randperm(R, generator=g).to(torch.int32)
    s = sl.from_rows(x.to(DEV), perm.to(DEV), planes)
    assert s.Cp % 8 == 0 and s.Cp >= C and s.t.shape == sl.Slab.shape_for(R, B, s.Cp, planes)
    # internal row i holds caller row perm[i]; padded channels and tail samples are zero
    raw = s.t.float().cpu()  # (R, NB, P, Cp/8, 128, 8)
    val = raw.sum(2).permute(0, 1, 3, 2, 4).reshape(R, -1, s.Cp)  # (R, NB*128, Cp)
    want = x[:, perm.long(), :].permute(1, 0, 2)
    tol = 2.0 ** -8 if planes == 1 else 2.0 ** -15
    assert relerr(val[:, :B, :C], want) <= tol
    assert (val[:, B:, :] == 0).all() and (val[:, :, C:] == 0).all()
    back = sl.to_rows(s, perm.to(DEV), torch.float32)
    assert relerr(back, x) <= tol
    if planes == 1:
        assert torch.equal(back.cpu(), _bf16(x))


def _slab_values(s):
    """Slab tensor -> (rows, NB*128, Cp) fp32 on the CPU (planes summed)."""
    raw = s.t.detach().float().cpu()  # (R, NB, P, Cp/8, 128, 8)
    return raw.sum(2).permute(0, 1, 3, 2, 4).reshape(s.rows, -1, s.Cp)


@pytest.mark.parametrize("C", [3, 32, 128])
@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("use_perm", [False, True])
def test_rows_slab_conversions_bf16_rows_and_gradient_path(sl, C, planes, use_perm):
    """The tiled conversion kernels with bf16 row tensors, without a permutation, over several row tiles and a ragged batch
    chunk, and the gradient path of to_rows: incoming gradient times act'(y) of the slab's producer, dummy row masked."""
    R, B = 70, 131
    g = torch.Generator().manual_seed(C * 7 + planes)
    x = torch.randn(B, R, C, generator=g)
    perm = torch.randperm(R, generator=g).to(torch.int32).to(DEV) if use_perm else None
    pl = perm.long().cpu() if use_perm else torch.arange(R)
    xb = x.bfloat16()
    s = sl.from_rows(xb.to(DEV), perm, planes)
    val = _slab_values(s)
    assert torch.equal(val[:, :B, :C], xb.float()[:, pl, :].permute(1, 0, 2))
    assert (val[:, B:, :] == 0).all() and (val[:, :, C:] == 0).all()
    back = sl.to_rows(s, perm, torch.bfloat16)
    assert back.dtype == torch.bfloat16 and torch.equal(back.cpu(), xb)
    # gradient path: y = the slab above taken as an ELU layer's masked output
    s.t.requires_grad_(True)
    s.act, s.masked = 2, True
    out = sl.to_rows(s, perm, torch.float32)
    go = torch.randn(B, R, C, generator=g)
    out.backward(go.to(DEV))
    gs = sl.Slab(s.t.grad, s.rows, s.B, s.C, s.Cp, s.planes, 0, False)
    got = _slab_values(gs)
    y = val[:, :B, :C]
    want = go[:, pl, :].permute(1, 0, 2) * (torch.clamp(y, max=0.0) + 1.0)
    want[-1] = 0
    tol = 2.0 ** -8 if planes == 1 else 2.0 ** -15
    assert relerr(got[:, :B, :C], want) <= tol
    assert (got[:, B:, :] == 0).all() and (got[:, :, C:] == 0).all() and (got[-1] == 0).all()


def _conv_case(g, k):
    pre = f"conv{k}_"
    lvl, cin, cout, S, B = g[pre + "meta"].tolist()
    return pre, lvl, cin, cout, S, B, str(g[pre + "act"])


def _run_slab_conv(sl, x, w, b, table, act, gy, planes, perm=None, **flags):
    """x (B, R, Cin) fp32 leaf on DEV -> y rows, and grads (through from_rows / to_rows)."""
    geom = sl.SlabGeometry(table, table.shape[0], DEV, **flags)
    s = sl.from_rows(x, perm, planes)
    y = sl.to_rows(sl.spiral_conv(s, w, b, geom, act), perm, torch.float32)
    y.backward(gy)
    return y


@pytest.mark.parametrize("planes", [2, 1])
def test_slab_spiralconv_matches_reference_goldens(shb, sl, planes):
    """Every activation, live and dead dummy rows, odd channel counts, vs the reference's own outputs (fp32 mode) and vs
    the oracle on bf16-rounded operands (bf16 mode)."""
    from semantichuman_b200.indexing import normalise_spiral

    g = golden("golden_ops")
    _, sizes, ssz, spirals, _, _ = ref_args("small")
    for k in range(int(g["n_conv"])):
        pre, lvl, cin, cout, S, B, act = _conv_case(g, k)
        table = normalise_spiral(spirals[lvl])
        xr, wr, br, gyr = (torch.from_numpy(g[pre + n]) for n in ("x", "w", "b", "gy"))
        if planes == 1:
            xr, wr, gyr = _bf16(xr), _bf16(wr), _bf16(gyr)
        x = xr.to(DEV).requires_grad_(True)
        w = wr.to(DEV).requires_grad_(True)
        b = br.to(DEV).requires_grad_(True)
        y = _run_slab_conv(sl, x, w, b, table, act, gyr.to(DEV), planes)
        assert (y[:, -1] == 0).all()
        if planes == 2:
            ref = {n: g[pre + n] for n in ("y", "gx", "gw", "gb")}
            tol = TOL_F32
        else:
            xo, wo, bo = xr.clone().requires_grad_(True), wr.clone().requires_grad_(True), br.clone().requires_grad_(True)
            yo = so.spiral_conv(xo, spirals[lvl], wo, bo, act)
            yo.backward(gyr)
            ref = {"y": yo, "gx": xo.grad, "gw": wo.grad, "gb": bo.grad}
            tol = TOL_BF16
        assert relerr(y, ref["y"]) < tol, (k, act, "y")
        assert relerr(x.grad, ref["gx"]) < tol, (k, act, "gx")
        assert relerr(w.grad, ref["gw"]) < tol, (k, act, "gw")
        assert relerr(b.grad, ref["gb"]) < tol, (k, act, "gb")


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("shape", [(3, 16, 14), (16, 3, 14), (32, 16, 14), (32, 32, 13), (64, 32, 8), (128, 64, 8),
                                   (64, 128, 8), (256, 48, 9)])
def test_slab_spiralconv_many_tiles_vs_oracle(sl, planes, shape):
    """The persistent multi-tile path the benchmark runs: B = 256 (two batch chunks), several hundred tiles per CTA, ring
    and TMEM double-buffer wrap-around, a permuted internal order, restricted output rows.  Samples are independent, so the
    oracle evaluates a few of them (first, chunk boundary, last); weight/bias gradients are checked with gz confined to
    those samples."""
    cin, cout, S = shape
    B, R = 256, 700
    if planes == 2 and cin > 128:  # a two-plane 256-channel slab (128 KB) leaves no room for a ring: loud, not silent
        geom = sl.SlabGeometry(np.zeros((4, S), np.int32), 4, DEV)
        with pytest.raises(NotImplementedError):
            sl.spiral_conv(sl.from_rows(torch.zeros(2, 4, cin, device=DEV), None, 2), torch.zeros(cout, S * cin, device=DEV),
                           None, geom, "elu")
        return
    g = torch.Generator().manual_seed(cin * 7 + cout)
    table = torch.randint(0, R, (R, S), generator=g)
    table[:, 0] = torch.arange(R)
    table[torch.rand(R, S, generator=g) < 0.05] = R - 1          # padded spirals -> dummy row
    table[R - 1] = R - 1
    keep = torch.cat([torch.arange(0, R - 1, 2), torch.tensor([R - 1])])  # a fused selection pool: every second row
    tab = table[keep].numpy().astype(np.int32)
    x = torch.randn(B, R, cin, generator=g)
    x[:, -1] = 0                                                       # masked producer: dummy row is zero
    w = torch.randn(cout, S * cin, generator=g) / np.sqrt(S * cin)
    b = torch.randn(cout, generator=g) * 0.1
    pick = [0, 1, 127, 128, 255]
    gy = torch.zeros(B, len(keep), cout)
    gy[pick] = torch.randn(len(pick), len(keep), cout, generator=g)
    if planes == 1:
        x, w, gy = _bf16(x), _bf16(w), _bf16(gy)
    perm = torch.randperm(R - 1, generator=g)
    perm = torch.cat([perm, torch.tensor([R - 1])]).to(torch.int32)  # internal order; dummy stays last
    pos = torch.empty(R, dtype=torch.long)
    pos[perm.long()] = torch.arange(R)
    tab_int = pos[torch.from_numpy(tab).long()].numpy().astype(np.int32)  # internal row ids
    xd = x.to(DEV).requires_grad_(True)
    wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    geom = sl.SlabGeometry(tab_int, R, DEV, src_dummy_zero=True, dummy_row_grad=False)
    s = sl.from_rows(xd, perm.to(DEV), planes)
    ys = sl.spiral_conv(s, wd, bd, geom, "elu")
    y = sl.to_rows(ys, None, torch.float32)
    y.backward(gy.to(DEV))
    xo = x[pick].clone().requires_grad_(True)
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yo = so.spiral_conv(xo, table, wo, bo, "elu")[:, keep]
    yo = yo * torch.cat([torch.ones(len(keep) - 1), torch.zeros(1)]).view(1, -1, 1)  # restricted conv masks ITS last row
    yo.backward(gy[pick])
    tol = TOL_F32 if planes == 2 else TOL_BF16
    assert relerr(y[pick], yo) < tol
    gxo = xo.grad.clone()
    gxo[:, -1] = 0  # dummy_row_grad=False: the producer masks that row
    assert relerr(xd.grad[pick], gxo) < tol
    assert float(xd.grad[[2, 126, 129, 254]].abs().max()) == 0.0
    assert relerr(wd.grad, wo.grad) < tol
    assert relerr(bd.grad, bo.grad) < tol


@pytest.mark.parametrize("planes", [2, 1])
def test_slab_pool_matches_reference_goldens(shb, sl, planes):
    g = golden("golden_ops")
    _, _, _, _, D, U = ref_args("small")
    mats = {f"D{l}": m for l, m in enumerate(D)} | {f"U{l}": m for l, m in enumerate(U)}
    for k in range(int(g["n_pool"])):
        pre = f"pool{k}_"
        pm = shb.PoolMatrix.from_dense(mats[str(g[pre + "which"])].to(DEV))
        x = torch.from_numpy(g[pre + "x"]).to(DEV).requires_grad_(True)
        y = sl.to_rows(sl.pool(sl.from_rows(x, None, planes), pm), None, torch.float32)
        y.backward(torch.from_numpy(g[pre + "gy"]).to(DEV))
        tol = 3e-5 if planes == 2 else TOL_BF16
        assert relerr(y, g[pre + "y"]) < tol and relerr(x.grad, g[pre + "gx"]) < tol


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("C,B", [(32, 130), (8, 256), (128, 5)])
def test_slab_pool_every_row_length(shb, sl, planes, C, B):
    """Pool on a random matrix whose rows have 0, 1..4 (the specialised loops), 5..32 (four entries at a time) and more than
    32 entries (entries re-fetched per block), and whose transpose has its own mix: forward and backward against the dense
    product, the backward with the act'(y) factor and the dummy-row mask of the producer."""
    g = torch.Generator().manual_seed(11 * C + planes)
    rows_out, rows_in = 60, 75
    lens = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 13, 16, 17, 31, 32, 33, 40, 64, 75] + [1, 3] * 20
    lens = (lens + [2] * rows_out)[:rows_out]
    dense = torch.zeros(rows_out, rows_in)
    for r, n in enumerate(lens):
        cols = torch.randperm(rows_in, generator=g)[:n]
        dense[r, cols] = torch.randn(n, generator=g)
    pm = shb.PoolMatrix.from_dense(dense.to(DEV))
    x = torch.randn(B, rows_in, C, generator=g)
    s = sl.from_rows(x.to(DEV), None, planes)
    s.t.requires_grad_(True)
    s.act, s.masked = 2, True   # as if x were a masked ELU layer's output
    y = sl.pool(s, pm)
    yr = sl.to_rows(y, None, torch.float32)
    xq = _slab_values(s)[:, :B, :C].permute(1, 0, 2)   # what the kernels read (bf16 or hi + lo)
    want = torch.einsum("rk,bkc->brc", dense, xq)
    tol = 3e-5 if planes == 2 else TOL_BF16
    assert relerr(yr, want) < tol
    gy = torch.randn(B, rows_out, C, generator=g)
    y.t.backward(sl.from_rows(gy.to(DEV), None, planes).t)
    gq = _slab_values(sl.from_rows(gy.to(DEV), None, planes))[:, :B, :C].permute(1, 0, 2)
    gwant = torch.einsum("rk,brc->bkc", dense, gq) * (torch.clamp(xq, max=0.0) + 1.0)
    gwant[:, -1] = 0
    got = _slab_values(sl.Slab(s.t.grad, s.rows, s.B, s.C, s.Cp, s.planes, 0, False))[:, :B, :C].permute(1, 0, 2)
    assert relerr(got, gwant) < tol
    assert (got[:, -1] == 0).all()


def test_slab_backward_is_bit_reproducible(sl):
    g = torch.Generator().manual_seed(5)
    R, S, cin, cout, B = 300, 9, 32, 16, 200
    table = torch.randint(0, R, (R, S), generator=g).numpy().astype(np.int32)
    geom = sl.SlabGeometry(table, R, DEV)
    x = torch.randn(B, R, cin, generator=g).to(DEV)
    w = (torch.randn(cout, S * cin, generator=g) / 17).to(DEV)
    b = torch.zeros(cout, device=DEV)
    gy = torch.randn(B, R, cout, generator=g).to(DEV)
    outs = []
    for _ in range(3):
        xd, wd, bd = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = sl.to_rows(sl.spiral_conv(sl.from_rows(xd, None, 1), wd, bd, geom, "tanh"), None, torch.float32)
        y.backward(gy)
        outs.append((y.detach().clone(), xd.grad.clone(), wd.grad.clone(), bd.grad.clone()))
    for o in outs[1:]:
        for a, c in zip(outs[0], o):
            assert torch.equal(a, c)


@pytest.mark.parametrize("planes", [1, 2])
@pytest.mark.parametrize("B,R,C", [(1, 5, 3), (3, 37, 3), (130, 70, 3), (256, 101, 8), (200, 64, 1)])
@pytest.mark.parametrize("act,masked", [("identity", True), ("elu", False), ("tanh", True)])
@pytest.mark.parametrize("use_perm", [True, False])
def test_slab_l1_loss_matches_rows_path(sl, planes, B, R, C, act, masked, use_perm):
    """slab.l1_loss (shb_slab_l1_fwd / _bwd) against the path it replaces -- to_rows, F.l1_loss on the rows, and back through
    from_rows with act' and the dummy-row mask -- and against torch on the CPU: same loss (summation order aside) and the
    IDENTICAL gradient slab (both round gscale / n * sign to bf16 hi (+ lo) the same way).  Ragged batches, row counts that
    are not multiples of the 32-row tile, 1 / 3 / 8 channels, a non-unit upstream gradient."""
    from semantichuman_b200._capi import ACT_ENUM
    from semantichuman_b200 import functions as fn

    g = torch.Generator().manual_seed(B * 131 + R * 7 + C)
    rec = torch.randn(B, R, C, generator=g)
    target = (rec + 0.3 * torch.randn(B, R, C, generator=g)).to(DEV)
    perm = torch.randperm(R, generator=g).to(torch.int32).to(DEV) if use_perm else None
    base = sl.from_rows(rec.to(DEV), perm, planes)
    # exact ties (sign(0) = 0): the target takes the value the slab holds for a few elements
    held = sl.to_rows(base, perm, torch.float32)
    target[:, 0, :] = held[:, 0, :]
    res = {}
    for fused in (True, False):
        t = base.t.detach().clone().requires_grad_(True)
        s = sl.Slab(t, R, B, C, base.Cp, planes, ACT_ENUM[act], masked)
        loss = sl.l1_loss(s, target, perm) if fused else fn.l1_loss(sl.to_rows(s, perm, torch.float32), target)
        (loss * 1.7).backward()
        res[fused] = (loss.item(), t.grad.float().cpu())
    assert abs(res[True][0] - res[False][0]) <= 2e-6 * abs(res[False][0]) + 1e-8
    assert torch.equal(res[True][1], res[False][1])
    # and against plain torch on what the slab actually holds
    rows = sl.to_rows(base, perm, torch.float32).cpu()
    want = (rows - target.cpu()).abs().mean().item()
    assert abs(res[True][0] - want) <= 2e-6 * want + 1e-8
    assert res[True][1].abs().max() > 0
