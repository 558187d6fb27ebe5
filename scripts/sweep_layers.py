"""Layer sweep (BASELINE.json configs[1] / SURVEY 8d config 2): every SpiralConv shape of the default net plus wide
(128/256-channel) variants on both hierarchies (ds 2222 and 4444, 2-ring dilated and 1-ring spirals) and every Pool, forward /
weight gradient / input gradient separately, B = 256, L2 flushed between launches, through the slab operators.  Prints a
markdown table: ms, GB/s of algorithmic bytes and its fraction of the measured HBM peak, TFLOP/s and its fraction of the
measured bf16 tensor peak.

    python scripts/sweep_layers.py [--dtype bf16|fp32] [--no-reorder] > profiles/r02_sweep_bf16.md      # one B200
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn, slab
from semantichuman_b200.indexing import locality_order, normalise_spiral
from tests.golden.loader import Hierarchy

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--no-reorder", action="store_true", help="keep the mesh's own vertex numbering (the models re-order)")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--hier", default="2222:A,2222:B,4444:A,4444:B")
a = ap.parse_args()
dev, B = "cuda:0", a.batch
planes = 1 if a.dtype == "bf16" else 2
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peaks = json.load(open(pk)) if os.path.exists(pk) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
hbm, tf = peaks["hbm_gbs"], peaks["bf16_tflops"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CONVS = [(0, 3, 16), (1, 16, 32), (2, 32, 64), (3, 64, 128), (3, 128, 64), (2, 64, 32), (1, 32, 32), (0, 32, 16), (0, 16, 3),
         (3, 128, 128), (3, 256, 128), (2, 128, 256)]
POOL_C = {("D", 0): 16, ("D", 1): 32, ("D", 2): 64, ("D", 3): 128, ("U", 3): 128, ("U", 2): 64, ("U", 1): 32, ("U", 0): 32}


def timed(run):
    fn.TIMER = None
    for i in range(a.reps + 2):
        if i == 2:
            fn.TIMER = fn.KernelTimer()
        run()
    out = fn.TIMER.summary()
    fn.TIMER = None
    return out


def row(tag, op, shape, v):
    ms = v["ms"] / v["launches"]
    gbs = v["bytes"] / v["launches"] / ms / 1e6
    tfs = v["flops"] / v["launches"] / ms / 1e9
    print(f"| {tag} | {op} | {shape} | {ms:.4f} | {gbs:.0f} | {100 * gbs / hbm:.1f} % | {tfs:.1f} | {100 * tfs / tf:.1f} % |", flush=True)


print(f"B = {B}, {a.dtype} mode ({'bf16 operands' if planes == 1 else 'bf16 hi + lo operands, 3 products'}), vertex order: "
      f"{'mesh numbering' if a.no_reorder else 'locality order (as in the models)'}; peaks: HBM {hbm:.0f} GB/s, bf16 {tf:.0f} TFLOP/s (measured)\n")
print("| hierarchy / spirals | op | shape (rows_in > rows_out x S x Cin > Cout) | ms | GB/s (algorithmic) | of HBM peak | TFLOP/s | of bf16 peak |")
print("|---|---|---|---|---|---|---|---|")
for item in a.hier.split(","):
    tag, cfg = item.split(":")
    h = Hierarchy(tag, spiral_cfg=cfg)
    for lvl, cin, cout in CONVS:
        if planes == 2 and max(cin, cout) > 128:
            continue  # a two-plane 256-channel slab leaves no room for a ring (the operator raises)
        t = normalise_spiral(h.spirals()[lvl])
        if not a.no_reorder:
            perm = np.concatenate([locality_order(t), [t.shape[0] - 1]])
            pos = np.empty(len(perm), np.int64); pos[perm] = np.arange(len(perm))
            t = pos[t[perm]].astype(np.int32)
        geom = shb.SpiralGeometry(t, t.shape[0], dev, src_dummy_zero=True, dummy_row_grad=False)
        x = torch.randn(B, geom.rows_in, cin, device=dev)
        x[:, -1] = 0
        xs = slab.from_rows(x, None, planes)
        xs.t.requires_grad_(True); xs.act = 2; xs.masked = True
        w = (torch.randn(cout, geom.S * cin, device=dev) / (geom.S * cin) ** 0.5).requires_grad_(True)
        b = torch.zeros(cout, device=dev, requires_grad=True)
        gy = slab.from_rows(torch.randn(B, geom.rows_out, cout, device=dev), None, planes).t

        def run():
            flush.zero_()
            y = slab.spiral_conv(xs, w, b, geom, "elu")
            flush.zero_()
            y.t.backward(gy)
        for k, v in timed(run).items():
            if "weight_images" not in k:
                row(f"{tag} / {cfg}", k.split("[")[0].replace("slabconv_", ""), k.split("[")[1].rstrip("]"), v)
        del x, xs, gy
    for (kind, l), C in POOL_C.items():
        pm = shb.PoolMatrix.from_scipy_padded((h.D_sp if kind == "D" else h.U_sp)[l], dev)
        xs = slab.from_rows(torch.randn(B, pm.rows_in, C, device=dev), None, planes)
        xs.t.requires_grad_(True); xs.act = 2; xs.masked = True
        gy = slab.from_rows(torch.randn(B, pm.rows_out, C, device=dev), None, planes).t

        def run():
            flush.zero_()
            y = slab.pool(xs, pm)
            flush.zero_()
            y.t.backward(gy)
        for k, v in timed(run).items():
            row(f"{tag} / {cfg}", ("pool " + kind + str(l)) + (" bwd" if "bwd" in k else ""), k.split("[")[1].rstrip("]"), v)
