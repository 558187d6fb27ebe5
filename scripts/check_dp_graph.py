"""2-rank check (torchrun, one rank per GPU): TrainStep replayed as a CUDA graph -- bucket all-reduces captured inside --
follows the eager data-parallel trajectory, and every rank ends with identical parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_dp_graph.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import semantichuman_b200 as shb
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes
from semantichuman_b200.train import TrainStep

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
h = Hierarchy("small")
fe = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
fd = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
Dsp, Usp = h.sparse_DU()
xs = [synthetic_meshes(h.verts0, 8, seed=10 * rank + s, noise=0.05).to(dev) for s in range(3)]


def make(graph, dtype):
    model = shb.SpiralAutoencoder(fe, fd, latent_size=32, sizes=h.sizes, spiral_sizes=h.spiral_sizes,
                                  spirals=h.spirals(dev), D=Dsp, U=Usp, device=dev)
    fill_deterministic_(model, seed=2)
    return TrainStep(model.to(dev).set_compute_dtype(dtype), graph=graph)


ok = True
for dtype in (torch.float32, torch.bfloat16):
    eager = make(False, dtype)
    for _ in range(3):
        eager(xs[0])
    le = [eager(x).item() for x in xs * 2]
    graph = make(True, dtype).capture(xs[0])
    lg = [graph(x).item() for x in xs * 2]
    good = all(abs(a - b) <= 1e-5 * abs(b) + 1e-7 for a, b in zip(lg, le))
    # replicas stay in lockstep: parameters identical across ranks after the graph-replayed steps
    flat = torch.cat([p.detach().flatten().float() for p in graph.model.parameters()])
    other = flat.clone()
    dist.broadcast(other, src=0)
    same = bool((flat == other).all())
    pe = torch.cat([p.detach().flatten().float() for p in eager.model.parameters()])
    drift = float((flat - pe).abs().max() / pe.abs().max())
    print(f"rank {rank} {dtype}: losses match={good} replicas identical={same} param drift vs eager={drift:.2e}", flush=True)
    ok = ok and good and same and drift < 1e-4
    graph.release()
dist.barrier()
dist.destroy_process_group()
print("DP-GRAPH", "OK" if ok else "FAIL", flush=True)
sys.exit(0 if ok else 1)
