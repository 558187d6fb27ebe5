"""Two-rank NCCL test of the data-parallel step; skipped on a box with fewer than two GPUs.

  * gradients of the batch-sharded step (GradSync buckets, NCCL AVG over NVLink) == gradients of the single-GPU step on the
    whole batch, to 1e-6 (fp32 mode);
  * the step replayed as ONE CUDA graph per rank -- the bucket all-reduces captured inside -- follows the eager
    data-parallel trajectory, and the replicas end bit-identical across ranks."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    import semantichuman_b200 as shb
    from semantichuman_b200.train import TrainStep
    from tests.golden.loader import Hierarchy
    from tests.golden.synthetic import fill_deterministic_, synthetic_meshes

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    h = Hierarchy("small")
    fe = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
    fd = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
    Dsp, Usp = h.sparse_DU()
    full = [synthetic_meshes(h.verts0, 8, seed=s, noise=0.05).to(dev) for s in range(3)]

    def make(graph, dtype):
        model = shb.SpiralAutoencoder(fe, fd, latent_size=32, sizes=h.sizes, spiral_sizes=h.spiral_sizes,
                                      spirals=h.spirals(dev), D=Dsp, U=Usp, device=dev)
        fill_deterministic_(model, seed=2)
        return TrainStep(model.to(dev).set_compute_dtype(dtype), graph=graph)

    # single-GPU reference gradients on the WHOLE batch, before the process group exists (GradSync then has world 1)
    single = make(False, torch.float32)
    single.optim = None
    single(full[0])
    ref = {n: p.grad.clone() for n, p in single.model.named_parameters()}

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    per = 8 // world
    shard = [x[rank * per:(rank + 1) * per].contiguous() for x in full]
    ok = True
    dp = make(False, torch.float32)
    dp.optim = None
    dp(shard[0])
    worst = max(float((p.grad - ref[n]).abs().max() / ref[n].abs().max()) for n, p in dp.model.named_parameters())
    ok = ok and worst < 1e-5  # equal shards, AVG: the global-batch mean gradient (summation order differs: 1e-6 level)
    msgs = [f"rank {rank}: DP grads vs single-GPU full batch: worst rel err {worst:.2e}"]
    for dtype in (torch.float32, torch.bfloat16):
        eager = make(False, dtype)
        for _ in range(3):
            eager(shard[0])
        le = [eager(x).item() for x in shard * 2]
        graph = make(True, dtype).capture(shard[0])
        lg = [graph(x).item() for x in shard * 2]
        good = all(abs(a - b) <= 1e-5 * abs(b) + 1e-7 for a, b in zip(lg, le))
        flat = torch.cat([p.detach().flatten().float() for p in graph.model.parameters()])
        other = flat.clone()
        dist.broadcast(other, src=0)
        same = bool((flat == other).all())
        msgs.append(f"rank {rank} {dtype}: graph == eager {good}, replicas identical {same}")
        ok = ok and good and same
        graph.release()
    dist.barrier()
    dist.destroy_process_group()
    out[rank] = (ok, msgs)


def test_two_rank_nccl_step_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(2):
        ok, msgs = out[r]
        print("\n".join(msgs))
        assert ok, msgs
