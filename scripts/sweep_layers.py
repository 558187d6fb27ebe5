"""Layer sweep: every SpiralConv shape of the default net on both hierarchies (ds 2222 and 4444, the
2-ring and the 1-ring spirals) plus every pool, forward / weight-gradient / input-gradient separately, B = 256, L2 flushed
between launches.  Prints a markdown table: ms, GB/s of algorithmic bytes, fraction of the measured HBM peak, TFLOP/s.

    python scripts/sweep_layers.py [--dtype bf16|fp32] [--reorder]      # one B200
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import semantichuman_b200 as shb
from semantichuman_b200 import functions as fn
from tests.golden.loader import Hierarchy
from semantichuman_b200.indexing import locality_order, normalise_spiral

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--reorder", action="store_true", help="relabel the vertices with the model's locality order first")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev, B = "cuda:0", a.batch
dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
peaks = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6555.2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CONVS = [(0, 3, 16), (1, 16, 32), (2, 32, 64), (3, 64, 128), (3, 128, 64), (2, 64, 32), (1, 32, 32), (0, 32, 16), (0, 16, 3)]
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


print("| hierarchy / spirals | op | shape | ms | GB/s (algorithmic) | of HBM peak | TFLOP/s |")
print("|---|---|---|---|---|---|---|")
for tag, cfg in (("2222", "A"), ("2222", "B"), ("4444", "A"), ("4444", "B")):
    h = Hierarchy(tag, spiral_cfg=cfg)
    for lvl, cin, cout in CONVS:
        table = normalise_spiral(h.spirals()[lvl])
        if a.reorder:
            V = table.shape[0] - 1
            full = np.concatenate([locality_order(table), [V]])
            pos = np.empty(V + 1, np.int64)
            pos[full] = np.arange(V + 1)
            table = pos[table[full]]
        geom = shb.SpiralGeometry(table.astype(np.int32), table.shape[0], dev)
        x = torch.randn(B, geom.rows_in, cin, device=dev).to(dt).requires_grad_(True)
        w = (torch.randn(cout, geom.S * cin, device=dev) / (geom.S * cin) ** 0.5).requires_grad_(True)
        b = torch.zeros(cout, device=dev, requires_grad=True)
        gy = torch.randn(B, geom.rows_out, cout, device=dev).to(dt)

        def run():
            flush.zero_()
            y = shb.spiral_conv(x, w, b, geom, "elu")
            flush.zero_()
            y.backward(gy)

        for k, v in timed(run).items():
            if "bwd_act" in k or "pad" in k:
                continue
            ms = v["ms"] / v["launches"]
            gbs = v["bytes"] / v["launches"] / ms / 1e6
            print(f"| {tag}/{cfg} | {k.split('[')[0]} | {geom.rows_in}x{geom.S}x{cin}>{cout} | {ms:.4f} | {gbs:.0f} | "
                  f"{100 * gbs / hbm:.1f} % | {v['flops'] / v['launches'] / ms / 1e9:.1f} |")
    if cfg == "A":
        for (kind, l), C in POOL_C.items():
            pm = shb.PoolMatrix.from_scipy_padded((h.D_sp if kind == "D" else h.U_sp)[l], dev)
            x = torch.randn(B, pm.rows_in, C, device=dev).to(dt).requires_grad_(True)

            def run():
                flush.zero_()
                y = shb.pool(x, pm)
                flush.zero_()
                y.backward(torch.ones_like(y))

            for k, v in timed(run).items():
                ms = v["ms"] / v["launches"]
                gbs = v["bytes"] / v["launches"] / ms / 1e6
                print(f"| {tag} | {k.split('[')[0]} {kind}{l} | {pm.rows_in}>{pm.rows_out}x{C} | {ms:.4f} | {gbs:.0f} | "
                      f"{100 * gbs / hbm:.1f} % | - |")
