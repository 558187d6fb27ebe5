"""Drop-in test through the REFERENCE'S OWN TRAINING LOOP (train_funcs.train_autoencoder_dataloader, train_funcs.py:474-582):
the loop is imported unmodified from the staged copy under oracle/_ref (oracle/build_ref.py) and driven twice with the
same data, weights, optimizer and scheduler -- once with the reference's models.SpiralAutoencoder on the CPU, once with
semantichuman_b200.SpiralAutoencoder on the GPU.  Covers what the class-level tests cannot: optimizer owned by the caller,
the per-epoch ``model.cpu()`` -> ``torch.save(state_dict)`` -> ``model.to(device)`` round trip (train_funcs.py:442-457),
``model.eval()`` / ``no_grad`` validation passes, checkpoints interchangeable between the two implementations."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import filters_from_golden, golden, params_from_golden, ref_args, relerr

build_ref = pytest.importorskip("oracle.build_ref")


class _Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step):
        self.scalars.setdefault(tag, []).append(float(value))


class _Data(torch.utils.data.Dataset):
    def __init__(self, x):
        self.x = x

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return {"verts": self.x[i], "idx": i}


class _Shape:
    class reference_mesh:  # noqa: N801
        f = np.zeros((4, 3), np.int64)

    @staticmethod
    def save_meshes(*a, **k):
        pass


def _run_loop(tf, model, device, tmpdir, tag, x_train, x_val, epochs=2):
    cfg = tf.cfg
    cfg.TRAIN.edgereg_epoch, cfg.TRAIN.edgereg_w, cfg.TRAIN.ck_frequency = 10 ** 9, 0.0, 1
    optim = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-5)  # main.py:262: the caller owns the optimizer
    sched = torch.optim.lr_scheduler.StepLR(optim, 1, 0.99)                    # main.py:265
    w = _Writer()
    dl_tr = torch.utils.data.DataLoader(_Data(x_train), batch_size=4, shuffle=False)
    dl_va = torch.utils.data.DataLoader(_Data(x_val), batch_size=4, shuffle=False)
    tf.train_autoencoder_dataloader(dl_tr, dl_va, device, model, optim, torch.nn.functional.l1_loss, 1, epochs, 1, None, sched, w,
                                    _Shape, str(tmpdir), str(tmpdir), tag, np.zeros((24, x_train.shape[1] - 1)), {}, [], False)
    return w.scalars, [os.path.join(str(tmpdir), f"{tag}{e}.pth.tar") for e in range(1, epochs + 1)]


def _data(h, n, seed):
    from tests.golden.synthetic import synthetic_meshes

    return synthetic_meshes(h.verts0, n, seed=seed, noise=0.05)


def test_reference_loop_runs_with_the_reference_model(tmp_path):
    """CPU sanity of the harness itself (no GPU needed): the staged loop + staged model run and write checkpoints."""
    if not build_ref.available(loop=True):
        pytest.skip("oracle/_ref is not staged")
    tf = build_ref.import_reference_module("train_funcs")
    ref = build_ref.import_reference_models()
    g = golden("golden_ae_small")
    h, sizes, ssz, spirals, D, U = ref_args("small")
    fe, fd = filters_from_golden(g)
    model = ref.SpiralAutoencoder(fe, fd, int(g["latent"]), sizes, ssz, spirals, D, U, device=torch.device("cpu"))
    model.load_state_dict(params_from_golden(g))
    scalars, cks = _run_loop(tf, model, torch.device("cpu"), tmp_path, "ref", _data(h, 8, 1), _data(h, 4, 2), epochs=1)
    assert len(scalars["avg_epoch_train_loss"]) == 1 and os.path.exists(cks[0])


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [torch.float32, torch.bfloat16])
def test_reference_loop_drives_the_cuda_model(tmp_path, mode):
    if not build_ref.available(loop=True):
        pytest.skip("oracle/_ref is not staged")
    import semantichuman_b200 as shb

    tf = build_ref.import_reference_module("train_funcs")
    ref = build_ref.import_reference_models()
    g = golden("golden_ae_small")
    h, sizes, ssz, spirals, D, U = ref_args("small")
    fe, fd = filters_from_golden(g)
    xtr, xva = _data(h, 8, 1), _data(h, 4, 2)
    rmodel = ref.SpiralAutoencoder(fe, fd, int(g["latent"]), sizes, ssz, spirals, D, U, device=torch.device("cpu"))
    rmodel.load_state_dict(params_from_golden(g))
    rs, rck = _run_loop(tf, rmodel, torch.device("cpu"), tmp_path, "ref", xtr, xva)
    dev = torch.device("cuda:0")
    model = shb.SpiralAutoencoder(fe, fd, latent_size=int(g["latent"]), sizes=sizes, spiral_sizes=ssz,
                                  spirals=[s.to(dev) for s in spirals], D=[d.to(dev) for d in D], U=[u.to(dev) for u in U],
                                  device=dev).to(dev).set_compute_dtype(mode)
    model.load_state_dict(params_from_golden(g), strict=True)
    ms, mck = _run_loop(tf, model, dev, tmp_path, "shb", xtr, xva)
    tol = 1e-4 if mode == torch.float32 else 2e-2
    for tag in ("avg_epoch_train_loss", "avg_epoch_valid_loss", "loss/loss/rec_loss"):
        assert len(ms[tag]) == len(rs[tag]) > 0
        assert relerr(torch.tensor(ms[tag]), torch.tensor(rs[tag])) < tol, tag
    # after the loop the model is back on the device (train_funcs.py:457) and still works
    assert next(model.parameters()).is_cuda
    model(xtr[:2].to(dev))
    # checkpoints are interchangeable: same keys, and (fp32 mode) the same trained weights
    a, b = torch.load(mck[-1], weights_only=False), torch.load(rck[-1], weights_only=False)
    assert list(a["autoencoder_state_dict"].keys()) == list(b["autoencoder_state_dict"].keys())
    rmodel.load_state_dict(a["autoencoder_state_dict"], strict=True)
    model.load_state_dict(b["autoencoder_state_dict"], strict=True)
    if mode == torch.float32:
        for k in a["autoencoder_state_dict"]:
            assert relerr(a["autoencoder_state_dict"][k], b["autoencoder_state_dict"][k]) < 1e-3, k
