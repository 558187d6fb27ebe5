"""Time one SpiralConv layer of the slab path (fwd / wgrad / dgrad) at full size, L2 flushed between launches; also the
target command for ncu captures.   LAYER=lvl,cin,cout  B=256  DT=bf16|fp32  REORDER=1  GCONV=0 (row-per-tile kernel only)
TRACE=fwd|all (library built with SHB_NVCC_FLAGS=-DSHB_GCONV_TRACE: role timeline of the last grouped-conv launch)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from semantichuman_b200 import functions as fn, slab
from semantichuman_b200.indexing import locality_order, normalise_spiral
from tests.golden.loader import Hierarchy

dev = "cuda:0"
slab.GCONV_ENABLED = os.environ.get("GCONV", "1") != "0"
B = int(os.environ.get("B", "256")); reps = int(os.environ.get("REPS", "5"))
planes = 1 if os.environ.get("DT", "bf16") == "bf16" else 2
h = Hierarchy(os.environ.get("HIER", "2222"))
layers = os.environ.get("LAYERS", os.environ.get("LAYER", "0,32,16"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for spec in layers.split(";"):
    lvl, cin, cout = (int(v) for v in spec.split(","))
    t = normalise_spiral(h.spirals()[lvl])
    if os.environ.get("REORDER", "1") == "1":
        perm = np.concatenate([locality_order(t), [t.shape[0] - 1]])
        pos = np.empty(len(perm), np.int64); pos[perm] = np.arange(len(perm))
        t = pos[t[perm]].astype(np.int32)
    geom = slab.SlabGeometry(t, t.shape[0], dev, src_dummy_zero=True, dummy_row_grad=False)
    x = torch.randn(B, geom.rows_in, cin, device=dev)
    x[:, -1] = 0
    xs = slab.from_rows(x, None, planes)
    xs.t.requires_grad_(True); xs.act = 2; xs.masked = True   # as if produced by a masked ELU layer
    w = (torch.randn(cout, geom.S * cin, device=dev) / (geom.S * cin) ** 0.5).requires_grad_(True)
    b = torch.zeros(cout, device=dev, requires_grad=True)
    gy = slab.from_rows(torch.randn(B, geom.rows_out, cout, device=dev), None, planes).t
    fn.TIMER = None
    for i in range(reps + 2):
        if i == 2:
            fn.TIMER = fn.KernelTimer()
        flush.zero_()
        y = slab.spiral_conv(xs, w, b, geom, "elu")
        flush.zero_()
        if os.environ.get("TRACE") != "fwd":
            y.t.backward(gy)
    per = fn.TIMER.summary()
    for k, v in per.items():
        if "weight_images" in k:
            continue
        ms = v["ms"] / v["launches"]
        print(f"{k:46s} {ms:8.4f} ms  {v['bytes']/v['launches']/ms/1e6:8.1f} GB/s alg  {v['flops']/v['launches']/ms/1e9:8.1f} TFLOP/s"
              f"  entries f/b {geom.n_fwd_entries}/{geom.n_bwd_entries}", flush=True)
    if os.environ.get("TRACE"):
        import ctypes
        from semantichuman_b200._capi import lib as _lib
        raw = ctypes.CDLL(_lib._name)
        buf = (ctypes.c_longlong * (296 * 16))()
        torch.cuda.synchronize()
        raw.shb_gconv_trace_read(buf)
        a = np.array(buf[:]).reshape(296, 16).astype(np.float64)
        a = a[a[:, 7] > 0]
        names = ["prod0 wait-empty", "prod0 total", "mma wait-full", "mma wait-acc-free", "mma total", "epi wait-acc", "epi total",
                 "tiles", "mma issue", "prod0 records"]
        print(f"  trace of the LAST slab_gconv launch ({len(a)} CTAs), mean cycles per CTA:")
        for i, n in enumerate(names):
            print(f"    {n:18s} {a[:, i].mean():10.0f}   (min {a[:, i].min():.0f}, max {a[:, i].max():.0f})")
