"""Reference checkpoint format (train_funcs.py:451-455, main.py:277-290): keys, resume and fine-tune semantics, and a
file written the way the reference writes it (plain torch.save of the four-key dict) loading through the helper."""
import os

import pytest
import torch

from semantichuman_b200.checkpoint import KEYS, load_checkpoint, save_checkpoint


def _setup(seed):
    torch.manual_seed(seed)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    optim = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-5)  # main.py:262
    sched = torch.optim.lr_scheduler.StepLR(optim, 2, gamma=0.5)              # main.py:264
    return model, optim, sched


def _train(model, optim, sched, steps):
    g = torch.Generator().manual_seed(3)
    for _ in range(steps):
        x = torch.randn(8, 6, generator=g)
        optim.zero_grad()
        model(x).abs().mean().backward()
        optim.step()
        sched.step()


def test_round_trip_resume_and_finetune(tmp_path):
    model, optim, sched = _setup(0)
    _train(model, optim, sched, 3)
    path = os.path.join(tmp_path, "checkpoint7.pth.tar")
    save_checkpoint(path, model, optim, sched, epoch=7)
    raw = torch.load(path, weights_only=False)
    assert tuple(raw.keys()) == KEYS and raw["epoch"] == 7

    m2, o2, s2 = _setup(1)
    assert load_checkpoint(path, m2, o2, s2) == 8
    for a, b in zip(model.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    assert o2.state_dict()["state"][0]["exp_avg"].equal(optim.state_dict()["state"][0]["exp_avg"])
    assert s2.state_dict()["last_epoch"] == sched.state_dict()["last_epoch"]
    _train(model, optim, sched, 2)
    _train(m2, o2, s2, 2)  # identical continuation
    for a, b in zip(model.parameters(), m2.parameters()):
        assert torch.equal(a, b)

    m3, o3, s3 = _setup(2)
    assert load_checkpoint(path, m3, o3, s3, finetune=True) == 1
    assert len(o3.state_dict()["state"]) == 0  # optimizer untouched when fine-tuning


def test_reference_written_file_and_bad_file(tmp_path):
    model, optim, sched = _setup(0)
    path = os.path.join(tmp_path, "ref.pth.tar")
    torch.save({"epoch": 3, "autoencoder_state_dict": model.state_dict(), "optimizer_state_dict": optim.state_dict(),
                "scheduler_state_dict": sched.state_dict()}, path)  # exactly train_funcs.py:451-455
    m2, o2, s2 = _setup(5)
    assert load_checkpoint(path, m2, o2, s2) == 4
    bad = os.path.join(tmp_path, "bad.pth.tar")
    torch.save({"weights": {}}, bad)
    with pytest.raises(KeyError):
        load_checkpoint(bad, m2)
