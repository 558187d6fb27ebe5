"""Bone-guided model (SpiralAutoencoder_multiz_partkps) at full size -- SURVEY 8(d) config 5 and the f-1 row:
encode -> scale part codes -> decode and decode-only (no_grad, B=4096 in slices), plus one training step (B=256),
with the per-part heads as grouped kernels and as the reference's per-part nn.Linear loop.

    python scripts/bench_multiz.py            # one B200
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn
from tests.golden.loader import Hierarchy
from tests.golden.synthetic import fill_deterministic_, synthetic_meshes
from tests.golden.constants import KPS_INDEX_LIST, PART_LIST

dev = "cuda:0"
FENC = [[3, 16, 32, 64, 128], [[], [], [], [], []]]
FDEC = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]
h = Hierarchy("2222")
Dsp, Usp = h.sparse_DU()
# 17 parts = y-sorted equal slabs of the coarsest level
vc = h.level_verts(h.n_levels)
order = np.argsort(vc[:, 1], kind="stable")
parts = {n: np.sort(c) for n, c in zip(PART_LIST, np.array_split(order, len(PART_LIST)))}
out = {}
for grouped in (True, False):
    model = shb.SpiralAutoencoder_multiz_partkps(KPS_INDEX_LIST, parts, FENC, FDEC, latent_size=8, part_kps_latent_size=8,
                                                 sizes=h.sizes, spiral_sizes=h.spiral_sizes, spirals=h.spirals(dev), D=Dsp,
                                                 U=Usp, device=dev, grouped_heads=grouped)
    fill_deterministic_(model, seed=2)
    model = model.to(dev).set_compute_dtype(torch.bfloat16)
    tag = "grouped" if grouped else "per_part_linear"
    res = {"params": sum(p.numel() for p in model.parameters())}
    gen = torch.Generator(device=dev).manual_seed(0)
    # ---- inference: B = 4096 as 16 slices of 256
    xs = synthetic_meshes(h.verts0, 256, seed=3).to(dev)
    kps = torch.randn(256, 32, 3, device=dev, generator=gen)
    with torch.no_grad():
        for mode in ("encode_scale_decode", "decode_only"):
            z, zk, dummy = model.encode(xs, kps)
            def run():
                if mode == "encode_scale_decode":
                    z2, zk2, d2 = model.encode(xs, kps)
                    return model.decode(z2 * 1.1, zk2, d2)
                return model.decode(z, zk, dummy)
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(16):
                y = run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            res[mode] = {"ms_per_4096": ms, "meshes_per_s": 4096 / ms * 1e3}
    if grouped:
        # ---- the step the paper trains (train_funcs.py:128-392): three passes + six loss terms + one backward + Adam, B = 256
        from semantichuman_b200.train import BoneGuidedStep
        v0 = np.asarray(h.verts0)
        fine = [np.sort(c) for c in np.array_split(np.argsort(v0[:, 1], kind="stable"), len(PART_LIST))]
        g = torch.Generator().manual_seed(1)
        J = torch.rand(35, v0.shape[0], generator=g) ** 16
        J = J / J.sum(1, keepdim=True)
        skl = [[15, 12], [15, 12], [12, 9], [6, 0], [0, 1, 2], [1, 4], [4, 7], [7, 10], [2, 5], [5, 8], [8, 11], [16, 18],
               [18, 20], [20, 22], [17, 19], [19, 21], [21, 23]]
        keep = [i for i in range(35) if i not in (3, 13, 14)]
        data = [synthetic_meshes(h.verts0, 256, seed=40 + i).to(dev) for i in range(3)]
        meas = torch.rand(256, 16, device=dev, generator=gen) * 0.8 + 0.2
        for use_graph in (True, False):  # capture first: autograd's grad accumulators must first be created on the capture stream
            bg = BoneGuidedStep(model, J, keep, fine, skl, list(range(1, 13)), list(range(0, 12)), graph=use_graph)
            if use_graph:
                bg.capture(*data, meas)
            for _ in range(3):
                bg(*data, meas)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10):
                loss = bg(*data, meas)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["bone_guided_step_graph" if use_graph else "bone_guided_step_eager"] = {
                "ms": ms, "meshes_per_s": 3 * 256 / ms * 1e3, "loss": float(loss),
                "note": "3 x 256 meshes per step (reconstruction, interpolation and exchange batches)"}
    if grouped:
        # per-entry-point CUDA-event times of one eager bone-guided step (families; the remainder is torch glue)
        import collections, re
        fn.TIMER = fn.KernelTimer()
        bg(*data, meas)
        fam = collections.defaultdict(float)
        for k, v in fn.TIMER.summary().items():
            fam[re.sub(r"\[.*", "", k)] += v["ms"]
        fn.TIMER = None
        res["bone_guided_step_families_ms"] = {k: round(v, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])}
    # ---- one training step (recon + part-norm loss), B = 256, eager
    from semantichuman_b200.optim import Adam
    opt = Adam(model.parameters(), lr=1e-3)
    measure = torch.rand(256, 32, device=dev, generator=gen) * 0.8 + 0.2
    P = torch.arange(1, 13, dtype=torch.int32, device=dev)
    Q = torch.arange(0, 12, dtype=torch.int32, device=dev)
    def step():
        opt.zero_grad()
        xh, z, zk = model(xs, kps)
        loss = shb.l1_loss(xs, xh) + 1e-2 * shb.partnorm_loss(z, measure, P, Q, relative=True)
        loss.backward(); opt.step()
        return loss
    for _ in range(5):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    res["train_step"] = {"ms": ms, "meshes_per_s": 256 / ms * 1e3, "loss": float(loss)}
    out[tag] = res
    out[tag + "_xhat_checksum"] = float(y.double().abs().sum())
print(json.dumps(out))
