"""World-size-2 gloo test (CPU) of the data-parallel host logic in semantichuman_b200/dp.py: equal shards +
averaged flat-bucket all-reduce reproduce the single-process full-batch gradients; buckets fire in backward order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Lin(torch.nn.Module):
    def __init__(self, k, n):
        super().__init__()
        self.conv = torch.nn.Linear(k, n)


class _Toy(torch.nn.Module):
    """Same parameter naming as the autoencoder (conv.*, fc_latent_*, dconv.*) so the bucket rules apply."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.conv = torch.nn.ModuleList([_Lin(6, 8), _Lin(8, 8)])
        self.fc_latent_enc = torch.nn.Linear(8, 4)
        self.fc_latent_dec = torch.nn.Linear(4, 8)
        self.dconv = torch.nn.ModuleList([_Lin(8, 6)])

    def forward(self, x):
        for c in self.conv:
            x = torch.nn.functional.elu(c.conv(x))
        x = self.fc_latent_dec(self.fc_latent_enc(x))
        return self.dconv[0].conv(x)


class _ToySink(_Toy):
    """The two FC layers run through functions.LinearShadowFn (plain torch GEMMs, so it runs on CPU) and offer their weights
    as direct gradient sinks, as the autoencoder does in bf16 mode."""

    def direct_grad_params(self):
        return [self.fc_latent_enc.weight, self.fc_latent_dec.weight]

    def forward(self, x):
        from semantichuman_b200.functions import LinearShadowFn

        for c in self.conv:
            x = torch.nn.functional.elu(c.conv(x))
        for layer in (self.fc_latent_enc, self.fc_latent_dec):
            x = LinearShadowFn.apply(x, layer.weight, layer.bias, layer.weight.detach(), layer.bias.detach())
        return self.dconv[0].conv(x)


def _sink_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semantichuman_b200.dp import GradSync, shard_batch

    model = _ToySink()
    sync = GradSync(model, sink_dtype=torch.float32)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    lo, hi = shard_batch(8, rank, world)
    for _ in range(2):  # a sink is overwritten every step, never zeroed
        sync.reset()
        xs = x[lo:hi]
        (model(xs) - xs).abs().mean().backward()
        sync.finish()
    gm = sync.grad_map()
    grads = [(gm[p] if p in gm else p.grad).detach().clone().numpy() for p in model.parameters()]
    no_pgrad = all(p.grad is None for p in gm)
    q.put((rank, grads, len(sync.buckets), no_pgrad))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_sinks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sink_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    model = _Toy()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    (model(x) - x).abs().mean().backward()
    full = [p.grad for p in model.parameters()]
    for rank, grads, nb, no_pgrad in got:
        assert nb == 6 and no_pgrad   # 4 flat buckets + one sink bucket per FC weight; a sink never becomes p.grad
        for a, b in zip(grads, full):
            assert torch.allclose(torch.from_numpy(a), b, atol=1e-6, rtol=1e-5)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semantichuman_b200.dp import GradSync, shard_batch

    model = _Toy()
    sync = GradSync(model)
    order = []
    for i, b in enumerate(sync.buckets):
        b["params"][-1].register_post_accumulate_grad_hook(lambda _p, i=i: order.append(i))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    lo, hi = shard_batch(8, rank, world)
    for _ in range(2):  # two steps: reset() must clear the buckets
        sync.reset()
        xs = x[lo:hi]
        (model(xs) - xs).abs().mean().backward()
        sync.finish()
    # numpy copies, not tensors: a tensor travels through the queue as a file descriptor owned by this process, and the
    # parent may unpickle it after this process has exited
    q.put((rank, [p.grad.detach().clone().numpy() for p in model.parameters()], order, sync.grad_bytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_average_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    model = _Toy()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 6, generator=g)
    (model(x) - x).abs().mean().backward()
    full = [p.grad for p in model.parameters()]
    for rank, grads, order, nbytes in got:
        for a, b in zip(grads, full):
            assert torch.allclose(torch.from_numpy(a), b, atol=1e-6, rtol=1e-5)
        assert order[:4] == [0, 1, 2, 3]  # decoder convs -> fc_latent_dec -> fc_latent_enc -> encoder convs
        # every member starts on a 16-byte boundary inside its flat bucket (the optimizer kernel moves float4s)
        assert nbytes == 4 * sum((p.numel() + 3) // 4 * 4 for p in model.parameters())
    for a, b in zip(got[0][1], got[1][1]):
        assert (a == b).all()


def test_shard_batch():
    from semantichuman_b200.dp import shard_batch

    assert [shard_batch(2048, r, 8) for r in (0, 7)] == [(0, 256), (1792, 2048)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


def test_default_comm_sms_and_host_binding_without_gpu():
    import os

    from semantichuman_b200.dp import bind_host_to_device, default_comm_sms

    assert [default_comm_sms(w) for w in (1, 2, 4, 8, 16)] == [16, 16, 16, 32, 32]
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert bind_host_to_device(0) is None   # no device, no topology: nothing is changed, nothing raises
        assert os.sched_getaffinity(0) == before
