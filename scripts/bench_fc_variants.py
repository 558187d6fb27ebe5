"""Which cuBLAS call shape serves the two skinny FC GEMMs best (B = 256, K = 55296, N = 256, bf16)?  L2 flushed between
launches (256 MB fill), CUDA events, median of 20.  Prints microseconds per variant."""
import torch

dev = torch.device("cuda")
B, K, N = 256, 55296, 256
h = torch.randn(B, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.01).to(torch.bfloat16)        # fc_latent_enc.weight shadow
b = torch.randn(N, device=dev).to(torch.bfloat16)
g = torch.randn(B, N, device=dev).to(torch.bfloat16)                 # gradient w.r.t. z
wd = (torch.randn(K, N, device=dev) * 0.01).to(torch.bfloat16)       # fc_latent_dec.weight shadow
gd = torch.randn(B, K, device=dev).to(torch.bfloat16)                # gradient w.r.t. the decoder FC output
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
sink = torch.empty(N, K, dtype=torch.bfloat16, device=dev)
sinkd = torch.empty(K, N, dtype=torch.bfloat16, device=dev)
SPLITS = (8, 16, 32)


def timeit(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def splitk_fwd(s):
    hp = h.view(B, s, K // s).transpose(0, 1)                # (s, B, K/s)
    wp = w.view(N, s, K // s).permute(1, 2, 0)               # (s, K/s, N)
    return lambda: torch.bmm(hp, wp).sum(0, dtype=torch.float32)


variants = {
    "enc fwd: addmm(b, h, w.t())": lambda: torch.addmm(b, h, w.t()),
    "enc fwd: mm(h, w.t())": lambda: torch.mm(h, w.t()),
    "enc fwd: mm(w, h.t()).t()": lambda: torch.mm(w, h.t()),
    "enc fwd: linear(h, w, b)": lambda: torch.nn.functional.linear(h, w, b),
    "enc gW: mm(g.t(), h, out=sink)": lambda: torch.mm(g.t(), h, out=sink),
    "enc gW: mm(g.t().contiguous(), h, out=sink)": lambda: torch.mm(g.t().contiguous(), h, out=sink),
    "enc gW: mm(h.t(), g) [transposed result]": lambda: torch.mm(h.t(), g),
    "enc gx: mm(g, w)": lambda: torch.mm(g, w),
    "dec fwd: addmm(bd, z, wd.t())": lambda: torch.addmm(gd[0], g, wd.t()),
    "dec gW: mm(gd.t(), z, out=sinkd)": lambda: torch.mm(gd.t(), g, out=sinkd),
    "dec gz: mm(gd, wd)": lambda: torch.mm(gd, wd),
    "colsum 256x55296: sum(gd, 0, f32)": lambda: torch.sum(gd, 0, dtype=torch.float32),
    "colsum 256x256: sum(g, 0, f32)": lambda: torch.sum(g, 0, dtype=torch.float32),
    "colsum 256x55296 as ones @ gd": lambda: torch.mm(torch.ones(1, B, device=dev, dtype=torch.bfloat16), gd),
    "copy 28 MB bf16 (reference point)": lambda: sink.copy_(w),
}
for s in SPLITS:
    variants[f"enc fwd: bmm split-K {s} + sum"] = splitk_fwd(s)
for name, fn in variants.items():
    for _ in range(3):
        fn()
    print(f"{timeit(fn):8.1f} us  {name}", flush=True)
