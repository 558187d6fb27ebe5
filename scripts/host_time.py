"""Host enqueue time vs GPU time per training step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes
from semantichuman_b200.train import TrainStep
import bench
dev = torch.device("cuda", 0)
h = Hierarchy("2222"); Dsp, Usp = h.sparse_DU()
model = shb.SpiralAutoencoder(bench.FENC, bench.FDEC, latent_size=256, sizes=h.sizes, spiral_sizes=h.spiral_sizes,
                              spirals=h.spirals(dev), D=Dsp, U=Usp, device=dev)
fill_deterministic_(model, seed=2); model = model.to(dev).set_compute_dtype(torch.bfloat16)
step = TrainStep(model)
x = synthetic_meshes(h.verts0, 256, seed=1).to(dev)
for _ in range(5): step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(30): step(x)
t_host = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"host enqueue {1e3*t_host/30:.2f} ms/step, GPU {e0.elapsed_time(e1)/30:.2f} ms/step")
