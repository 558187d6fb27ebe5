"""BASELINE.json configs[4], second half: the 27 554-vertex subdivided template (tests/golden/hier_27554.npz).
Latent decode throughput (no_grad, 4096 meshes in slices of 256) and the full training step (B = 256, CUDA graph) of the
plain SpiralAutoencoder with default filters, nz = 256; bf16 and fp32 modes.   python scripts/bench_stress.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semantichuman_b200 as shb
from semantichuman_b200.train import TrainStep
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes

dev = "cuda:0"
FENC = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
FDEC = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
out = {}
for tag in (sys.argv[2:] or ["27554", "2222"]):
    h = Hierarchy(tag)
    Dsp, Usp = h.sparse_DU()
    model = shb.SpiralAutoencoder(FENC, FDEC, latent_size=256, sizes=h.sizes, spiral_sizes=h.spiral_sizes, spirals=h.spirals(dev),
                                  D=Dsp, U=Usp, device=dev)
    fill_deterministic_(model, seed=2)
    model = model.to(dev)
    res = {"vertices": h.sizes[0], "params": sum(p.numel() for p in model.parameters())}
    xs = [synthetic_meshes(h.verts0, 256, seed=s).to(dev) for s in range(2)]
    for dtype, name in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        model.set_compute_dtype(dtype)
        with torch.no_grad():
            z = model.encode(xs[0], False)
            for _ in range(2):
                model.decode(z)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(16):
                y = model.decode(z)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        res[f"decode_{name}"] = {"ms_per_4096": ms, "meshes_per_s": 4096 / ms * 1e3}
        step = TrainStep(model, graph=True).capture(xs[0])
        for i in range(3):
            step(xs[i % 2])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(10):
            loss = step(xs[i % 2])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[f"train_step_{name}"] = {"ms_per_step": ms, "meshes_per_s": 256 / ms * 1e3, "loss": float(loss)}
        step.release()
        del step
    out[tag] = res
    print(tag, json.dumps(res), flush=True)
    del model
    torch.cuda.empty_cache()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
