"""Dump CTA 0's per-stage timeline of the tcgen05 gather-GEMM (producer / MMA / epilogue roles)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import semantichuman_b200 as shb
from semantichuman_b200 import _capi
from tests.golden.loader import Hierarchy
lvl, cin, cout = (int(v) for v in os.environ.get("LAYER", "0,32,16").split(","))
dev = "cuda:0"; B = 256
h = Hierarchy("2222")
geom = shb.SpiralGeometry.from_spiral(h.spirals(dev)[lvl], dev)
x = torch.randn(B, geom.rows_in, cin, device=dev).bfloat16()
MODE = os.environ.get("MODE", "fwd")
if MODE == "dgrad":
    x.requires_grad_(True)
w = torch.randn(cout, geom.S * cin, device=dev) / (geom.S * cin) ** 0.5
b = torch.zeros(cout, device=dev)
lib = ctypes.CDLL(_capi.LIB_PATH)
lib.shbdbg_set_trace.argtypes = [ctypes.c_void_p]
for _ in range(3): y = shb.spiral_conv(x, w, b, geom, "elu")
tr = torch.zeros(3 * 4 * 512, dtype=torch.int64, device=dev)
lib.shbdbg_set_trace(tr.data_ptr())
y = shb.spiral_conv(x, w, b, geom, "elu")
if MODE == "dgrad":
    y.backward(torch.randn_like(y))   # the input-gradient kernel runs last and owns the trace buffer
torch.cuda.synchronize()
lib.shbdbg_set_trace(None)
t = tr.cpu().numpy().reshape(3, 4, 512)
t0 = t[t > 0].min()
P, M, E = t[0], t[1], t[2]
print("stage: prod[wait_begin wait_end issued] mma[wait_begin wait_end committed]  (cycles since start)")
for i in list(range(0, 16)):
    if P[0, i] == 0: break
    print(f"{i:4d}  P {P[0,i]-t0:8d} {P[1,i]-t0:8d} {P[2,i]-t0:8d}   M {M[0,i]-t0:8d} {M[1,i]-t0:8d} {M[2,i]-t0:8d}")
n = int((P[0] > 0).sum())
print("stages traced", n, "avg producer period", (P[0, n-1]-P[0, 8])/(n-9), "avg mma period", (M[1, n-1]-M[1, 8])/(n-9))
print("producer: avg wait", np.mean(P[1,8:n]-P[0,8:n]), "avg issue", np.mean(P[2,8:n]-P[1,8:n]), "avg rest-of-loop", np.mean(P[0,9:n]-P[2,8:n-1]))
print("mma: avg wait", np.mean(M[1,8:n]-M[0,8:n]), "avg issue+commit", np.mean(M[2,8:n]-M[1,8:n]), "avg rest", np.mean(M[0,9:n]-M[2,8:n-1]))
print("mma lag behind producer issue (M wait_end - P issued):", np.mean(M[1,8:n]-P[2,8:n]))
ne = int((E[0] > 0).sum())
print("epilogue tiles", ne, "avg period", (E[1, ne-1]-E[1,2])/(ne-3), "avg wait", np.mean(E[1,2:ne]-E[0,2:ne]))
