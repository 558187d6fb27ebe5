"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/shb200.h declares, and
its host entry points (index construction) agree with the numpy oracle bit for bit.  No kernel is launched."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import spiral_oracle as so
from semantichuman_b200 import _capi, indexing
from tests.golden.loader import Hierarchy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "shb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 16
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/shb200.h but not exported by libshb200.so"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature in _capi.py"
    assert sorted(_capi.SIGNATURES) == names
    assert _capi.lib.shb_abi_version() == 2
    assert b"invalid argument" in _capi.lib.shb_error_string(-1)


def test_argument_errors_are_reported_not_crashed():
    # null pointers are rejected before any launch (no GPU needed)
    assert _capi.lib.shb_slab_conv(None, None, None, None, None, None, None, 1, 1, 9, 16, 16, 16, 0, 0, 0, 1, None) == -1
    assert _capi.lib.shb_slab_pool(None, None, None, None, None, None, 1, 1, 16, 0, 0, 1, None) == -1
    assert _capi.lib.shb_slab_wgrad(None, None, None, None, None, None, 0, 1, 1, 9, 16, 16, 16, 16, 0, -1, 1, None) == -1
    assert _capi.lib.shb_adam_step(0, None, None, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0.0, None) == -1
    # shapes the kernels cannot take are refused loudly (no silent fallback): spiral length, channel counts, planes
    assert _capi.lib.shb_slab_conv_supported(33, 16, 16, 1) == 0 and _capi.lib.shb_slab_conv_supported(9, 24, 16, 1) == 0
    assert _capi.lib.shb_slab_conv_supported(14, 32, 16, 1) == 1 and _capi.lib.shb_slab_conv_supported(8, 128, 64, 2) == 1
    assert _capi.lib.shb_slab_wgrad_supported(8, 128, 64, 2) == 1 and _capi.lib.shb_slab_wgrad_supported(8, 48, 64, 1) == 0
    assert _capi.lib.shb_slab_tensor_bytes(6891, 256, 32, 1) == 6891 * 2 * 32 * 256
    with pytest.raises(RuntimeError):
        _capi.check(-3, "x")


@pytest.mark.parametrize("tag,cfg", [("small", "A"), ("open", "A"), ("2222", "A"), ("4444", "B")])
def test_inverse_spiral_tables_bit_exact(tag, cfg):
    h = Hierarchy(tag, cfg)
    for l, s in enumerate(h.spirals()):
        t = indexing.normalise_spiral(s.repeat(2, 1, 1))  # batch-replicated as models.py:122 passes it
        assert (t == so.normalise_spiral(s.numpy())).all()
        n = t.shape[0]
        assert t.min() >= 0 and t.max() == n - 1 and (t[-1] == n - 1).all()
        rp, sl = indexing.build_inverse_spiral_csr(t, n)
        rp2, sl2 = so.inverse_spiral_csr(t, n)
        assert rp.dtype == np.int32 and (rp == rp2).all() and (sl == sl2).all()
        # every slot appears exactly once, and the relation inverts the table
        assert sorted(sl.tolist()) == list(range(t.size))
        u = np.repeat(np.arange(n), np.diff(rp))
        assert (t.reshape(-1)[sl] == u).all()


def test_entry_lists_of_the_kernels():
    """SpiralGeometry (what the slab kernels consume): forward lists = the table itself, backward lists = the inverse-spiral
    CSR in ascending (j, s) order with masked rows dropped; restricted output rows; dummy-row handling."""
    h = Hierarchy("small")
    t = indexing.normalise_spiral(h.spirals()[0])
    n, S = t.shape
    g = indexing.SpiralGeometry(t, n, "cpu", zero_last_row=False, dummy_row_grad=True)
    assert g.dummy_split is None or g.dummy_split[0] >= 1
    pf, ef = g.ptr_f.numpy(), g.ent_f.numpy()
    assert (pf == np.arange(n + 1) * S).all() and ((ef >> 5) == t.reshape(-1)).all() and ((ef & 31) == np.tile(np.arange(S), n)).all()
    # backward: together with the split-off dummy list, exactly the canonical inverse CSR (stable counting sort)
    rp, sl = so.inverse_spiral_csr(t, n)
    pb, eb = g.ptr_b.numpy(), g.ent_b.numpy()[: g.n_bwd_entries]
    pos = (eb >> 5).astype(np.int64) * S + (eb & 31)
    if g.dummy_split is None:
        assert (pb == rp).all() and (pos == sl).all()
    else:
        T, bounds, ents = g.dummy_split[0], g.dummy_split[1].numpy(), g.dummy_split[2].numpy()
        assert (pb[:-1] == rp[:-1]).all() and (pos == sl[: rp[n - 1]]).all()
        dpos = (ents >> 5).astype(np.int64) * S + (ents & 31)
        assert (dpos == sl[rp[n - 1]:]).all() and bounds[0] == 0 and bounds[-1] == len(ents) and len(bounds) == T + 1
    # masked output row and dead dummy source row are dropped; forward entries on a known-zero dummy row too
    g2 = indexing.SpiralGeometry(t, n, "cpu", zero_last_row=True, dummy_row_grad=False, src_dummy_zero=True)
    assert g2.n_fwd_entries == int((t != n - 1).sum())
    e2 = g2.ent_b.numpy()[: g2.n_bwd_entries]
    assert ((e2 >> 5) != n - 1).all() and g2.ptr_b.numpy()[-1] == g2.ptr_b.numpy()[-2]
    # a conv fused with a selection pool evaluates only the kept rows
    keep = np.concatenate([h.raw["D0_col"], [n - 1]])
    g3 = g2.restricted(keep)
    assert g3.rows_out == len(keep) and g3.rows_in == n and (g3.table_host == t[keep]).all()
    with pytest.raises(ValueError):
        indexing.SpiralGeometry(np.zeros((4, 33), np.int32), 4, "cpu")


def test_batch_varying_spiral_rejected():
    import torch

    h = Hierarchy("small")
    s = h.spirals()[0].repeat(2, 1, 1)
    s[1, 0, 1] = 5
    with pytest.raises(ValueError):
        indexing.normalise_spiral(s)
    with pytest.raises(ValueError):
        indexing.normalise_spiral(torch.full((1, 4, 2), 9))


@pytest.mark.parametrize("tag", ["small", "2222"])
def test_dense_to_csr_and_transpose(tag):
    h = Hierarchy(tag)
    levels = range(h.n_levels) if tag == "small" else [2, 3]  # dense level 0 of 6890 is 95 MB; keep CPU suite light
    D, U = h.dense_DU()
    for l in levels:
        for m in (D[l], U[l]):
            d = m[0].numpy()
            rp, ci, v = indexing.dense_to_csr(d)
            rp2, ci2, v2 = so.dense_to_csr(d)
            assert (rp == rp2).all() and (ci == ci2).all() and (v == v2).all()
            trp, tci, tv = indexing.csr_transpose(rp, ci, v, d.shape[0], d.shape[1])
            rp3, ci3, v3 = so.dense_to_csr(d.T.copy())
            assert (trp == rp3).all() and (tci == ci3).all() and (tv == v3).all()
    # structural facts the fused conv + down-pool relies on
    import scipy.sparse as sp

    for l in range(h.n_levels):
        rp = h.D_sp[l].indptr
        assert (np.diff(rp) == 1).all() and (h.D_sp[l].data == 1.0).all()
        assert (np.diff(sp.csr_matrix(h.U_sp[l]).indptr) == 3).all()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "semantichuman_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("numpy oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in txt, f


def test_parameter_order_matches_reference_state_dict():
    """Optimizer state in the reference's checkpoints is indexed by parameter order (main.py:288): the drop-in must
    register conv, fc_latent_*, dconv in the reference's order (models.py:81-86,113).  Checked against the key order
    of the golden state_dicts, which were dumped from the reference modules."""
    import inspect

    from semantichuman_b200 import models
    from tests.helpers import golden

    for name, cls in (("golden_ae_small", "SpiralAutoencoder"), ("golden_multiz_small", "SpiralAutoencoder_multiz_partkps")):
        keys = [k[2:] for k in golden(name).files if k.startswith("p_")]
        groups = []
        for k in keys:
            g = k.split(".")[0]
            if not groups or groups[-1] != g:
                groups.append(g)
        src = inspect.getsource(getattr(models, cls)) + inspect.getsource(models._SpiralTrunk._init_trunk)
        pos = [src.index("self." + g + " =") if g not in ("conv", "dconv") else None for g in groups]
        assert groups[0] == "conv" and groups[-1] == "dconv", groups
        inner = [p for p in pos if p is not None]
        assert inner == sorted(inner), groups


def test_locality_order_and_permuted_pool_matrices():
    """Host logic of the internal vertex re-ordering: the order is a permutation that shrinks the per-tile gather
    footprint, and PoolMatrix.permuted / from_permutation are exact relabellings (checked against dense algebra)."""
    import scipy.sparse as sp
    from tests.golden.loader import Hierarchy
    from semantichuman_b200.indexing import PoolMatrix, locality_order, normalise_spiral

    h = Hierarchy("2222")
    table = normalise_spiral(h.spirals()[0])
    V = table.shape[0] - 1
    perm = locality_order(table)
    assert sorted(perm.tolist()) == list(range(V))
    full = np.concatenate([perm, [V]])
    pos = np.empty(V + 1, np.int64)
    pos[full] = np.arange(V + 1)
    relabelled = pos[table[full]]

    def footprint(t):
        return np.mean([len(np.unique(t[r:r + 128])) for r in range(0, V - 128, 128)])

    assert footprint(relabelled) < 0.5 * footprint(table)  # 1274 -> ~440 distinct rows per 128-row tile

    hs = Hierarchy("small")
    pm = PoolMatrix.from_scipy_padded(hs.U_sp[0], "cpu")
    rng = np.random.default_rng(0)
    pr = np.concatenate([rng.permutation(pm.rows_out - 1), [pm.rows_out - 1]])
    pc = np.concatenate([rng.permutation(pm.rows_in - 1), [pm.rows_in - 1]])

    def dense(m):
        rp, ci, va = m._host_csr
        return sp.csr_matrix((va, ci, rp), shape=(m.rows_out, m.rows_in)).toarray()

    q = pm.permuted(pr, pc)
    assert np.array_equal(dense(q), dense(pm)[pr][:, pc])
    assert np.all(np.diff(q._host_csr[0]) >= 0) and q.nnz == pm.nnz
    p = PoolMatrix.from_permutation(pr, "cpu")
    assert p.is_selection and np.array_equal(dense(p), np.eye(len(pr))[pr])
    # the transposed CSR of a permutation is the inverse gather
    inv = np.empty(len(pr), np.int64)
    inv[pr] = np.arange(len(pr))
    assert np.array_equal(p.t_colidx.numpy(), inv)


def test_group_layout_host_logic():
    """Packed-parameter offsets and the overlap check of the grouped heads (functions.GroupLayout)."""
    import torch
    from semantichuman_b200.functions import GroupLayout

    parts = [np.array([4, 0, 2]), np.array([1, 3]), np.array([5])]
    enc = GroupLayout(parts, rows=7, channels=4, latent=8, gather=True, device="cpu")
    assert enc.G == 3 and enc.max_rows == 3 and enc.disjoint and enc.supported()
    assert enc.gptr.tolist() == [0, 3, 5, 6] and enc.idx.tolist() == [4, 0, 2, 1, 3, 5]
    assert enc.woff.tolist() == [0, 3 * 4 * 8, 5 * 4 * 8] and enc.boff.tolist() == [0, 8, 16]
    dec = GroupLayout(parts, rows=7, channels=4, latent=16, gather=False, device="cpu")
    assert dec.boff.tolist() == [0, 12, 20] and dec.b_numel == [12, 8, 4]
    layers = [torch.nn.Linear(len(p) * 4, 8) for p in parts]
    w, b = enc.pack(layers)
    assert w.numel() == sum(enc.w_numel) and b.numel() == 24
    assert torch.equal(w[enc.woff[1]:enc.woff[1] + enc.w_numel[1]].view(8, 8), layers[1].weight)
    assert not GroupLayout([np.array([0, 1]), np.array([1, 2])], 4, 3, 8, True, "cpu").disjoint
    assert not GroupLayout(parts, 7, 4, 64, True, "cpu").supported()
    with pytest.raises(ValueError):
        GroupLayout([np.array([7])], 7, 4, 8, True, "cpu")
    with pytest.raises(ValueError):
        enc.pack([torch.nn.Linear(5, 8)] * 3)


@pytest.mark.parametrize("R,SPS", [(16, 4), (8, 8), (4, 2), (1, 1)])
def test_conv_group_program_covers_every_entry_once(R, SPS):
    """shb_build_conv_groups: the group program is a re-arrangement of the entry lists -- every (destination, slot, source)
    triple exactly once, every destination in exactly one group, record limits and flags as the kernel expects them."""
    from semantichuman_b200.indexing import build_conv_groups

    rng = np.random.default_rng(R * 10 + SPS)
    rows_dst, rows_src, S = 301, 280, 9
    table = (np.arange(rows_dst)[:, None] * 7 // 8 + rng.integers(-6, 7, size=(rows_dst, S))).clip(0, rows_src - 1)
    table[5] = rows_src - 1                       # one destination listing a single source S times
    keep = rng.random((rows_dst, S)) < 0.85
    keep[17] = False                              # a destination without entries
    counts = keep.sum(1)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    slots = np.tile(np.arange(S), rows_dst).reshape(rows_dst, S)
    ent = ((table.astype(np.int64) << 5) | slots)[keep].astype(np.int32)
    gptr, recs, gdst, gmask = build_conv_groups(ptr, ent, rows_dst, rows_src, R, SPS)
    rec = recs.reshape(-1, 48)
    assert gptr[0] == 0 and gptr[-1] == len(rec) and (np.diff(gptr) >= 1).all()
    got = []
    for g in range(len(gmask)):
        seen = set()
        last_src = -1
        for r in range(gptr[g], gptr[g + 1]):
            w = rec[r]
            ns, npair = w[0] & 15, (w[0] >> 4) & 63
            assert ns <= SPS and npair <= 32
            assert ((w[0] >> 10) & 1) == (r == gptr[g]) and ((w[0] >> 11) & 1) == (r == gptr[g + 1] - 1)
            assert (np.diff(w[1:1 + ns]) >= 0).all() and (ns == 0 or w[1] >= last_src)   # ascending sources: fixed order
            last_src = w[ns] if ns else last_src
            used = set()
            for pw in w[16:16 + npair]:
                k, dl, first, sl = pw & 7, (pw >> 3) & 31, (pw >> 8) & 1, (pw >> 9) & 31
                assert k < ns and dl < R and gdst[g * R + dl] >= 0
                assert first == (dl not in seen)
                seen.add(dl)
                used.add(k)
                got.append((int(gdst[g * R + dl]), int(sl), int(w[1 + k])))
            assert used == set(range(ns))         # no slab is loaded without a pair that reads it
        for i in range(R):
            d = gdst[g * R + i]
            if d >= 0:
                assert bool((gmask[g] >> i) & 1) == (counts[d] == 0)
    want = [(j, int(e & 31), int(e >> 5)) for j in range(rows_dst) for e in ent[ptr[j]:ptr[j + 1]]]
    assert sorted(got) == sorted(want)
    assert sorted(int(d) for d in gdst if d >= 0) == list(range(rows_dst))


def test_inverse_permutation_helper_is_cached_on_the_tensor():
    """slab.inverse_perm (the perm_inv argument of shb_slab_from_rows / shb_slab_to_rows): perm_inv[perm[i]] = i, built once."""
    import torch

    from semantichuman_b200 import slab

    g = torch.Generator().manual_seed(3)
    perm = torch.randperm(101, generator=g).to(torch.int32)
    inv = slab.inverse_perm(perm)
    assert inv.dtype == torch.int32 and torch.equal(inv[perm.long()].long(), torch.arange(101))
    assert slab.inverse_perm(perm) is inv and slab.inverse_perm(None) is None
