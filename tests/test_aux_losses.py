"""Batched auxiliary losses against the reference's per-sample functions looped over the batch the way its
training loop does (golden_aux_losses.npz): values and gradients w.r.t. the reconstruction."""
import numpy as np
import torch

from semantichuman_b200.aux_losses import PartVolumes, edge_length_loss, edge_ratio_loss
from tests.helpers import golden, relerr


def test_edge_ratio_loss():
    g = golden("golden_aux_losses")
    tx = torch.from_numpy(g["tx"])
    rec = torch.from_numpy(g["edge_rec"]).requires_grad_(True)
    loss = edge_ratio_loss(rec, tx, g["faces"])
    loss.backward()
    assert abs(loss.item() - float(g["edge_loss"])) < 1e-6 and relerr(rec.grad, g["edge_grec"]) < 1e-5


def test_edge_length_loss():
    g = golden("golden_aux_losses")
    tx = torch.from_numpy(g["tx"])
    rec = torch.from_numpy(g["elen_rec"]).requires_grad_(True)
    loss = edge_length_loss(tx, rec, g["edges"])
    loss.backward()
    assert abs(loss.item() - float(g["elen_loss"])) < 1e-7 and relerr(rec.grad, g["elen_grec"]) < 1e-5


def test_part_volume_loss():
    g = golden("golden_aux_losses")
    sizes = [int(v) for v in g["part_sizes"]]
    vdict = {"p%d" % i: p for i, p in enumerate(np.split(g["part_idx"], np.cumsum(sizes)[:-1]))}
    pv = PartVolumes(g["faces"], vdict)
    assert pv.onehot.shape == (len(g["faces"]), 4) and float(pv.onehot.sum(1).max()) == 1.0
    tx = torch.from_numpy(g["tx"])
    rec = torch.from_numpy(g["vol_rec"]).requires_grad_(True)
    loss = pv.loss(rec, tx, [int(v) for v in g["parts_used"]])
    loss.backward()
    assert abs(loss.item() - float(g["vol_loss"])) < 1e-6 and relerr(rec.grad, g["vol_grec"]) < 1e-4
