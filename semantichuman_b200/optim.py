"""Adam for the training step (main.py:262: ``torch.optim.Adam(model.parameters(), lr, weight_decay)``) on libshb200.

One multi-tensor kernel per step (plus a one-thread step-counter tick): fp32 master weights and moments, L2 weight decay
folded into the gradient exactly as torch does, bias corrections computed on the device from a device-resident step count
(so the step replays inside a CUDA graph), and -- in the same pass -- the bf16 shadow of the weights the bf16 mode feeds
to its FC GEMMs.  ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout, so the optimizer entry of the
reference's checkpoints (main.py:288) loads here and vice versa.
"""
import ctypes

import torch

from ._capi import check, lib
from .functions import _count, _stream


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, shadows=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("semantichuman_b200.optim.Adam needs contiguous float32 CUDA parameters (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.shadows = shadows if shadows is not None else {}   # {param: [bf16 tensor, version]} shared with the model
        n = len(self.params)
        self._arr = {k: (ctypes.c_void_p * n)() for k in "pgmvs"}
        self._numel = (ctypes.c_int64 * n)()
        self._gbf16 = (ctypes.c_uint8 * n)()

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def tick(self):
        """Advance the device-side step count on the current stream (for callers that split one optimizer step over several
        ``step(..., tick=False)`` launches, e.g. an early launch for the FC weights beside the rest of the backward)."""
        check(lib.shb_adam_tick(self.step_count.data_ptr(), _stream()), "shb_adam_tick")
        _count()

    @torch.no_grad()
    def step(self, grads=None, only=None, skip=None, tick=True):
        """`grads`: optional {parameter: gradient tensor} that takes precedence over ``p.grad`` -- the data-parallel
        gradient sinks of dp.GradSync (fp32 or bf16 buckets written by the producing GEMM and reduced in place).
        `only` / `skip`: collections of parameters to restrict this launch to / leave out; `tick=False`: the step count was
        advanced by tick() already."""
        k = 0
        live = []
        for i, p in enumerate(self.params):
            if (only is not None and not any(p is q for q in only)) or (skip is not None and any(p is q for q in skip)):
                continue
            g = grads.get(p) if grads else None
            if g is None:
                g = p.grad
            if g is None:
                continue
            if g.shape != p.shape:
                raise ValueError("gradient shape does not match its parameter")
            if g.dtype not in (torch.float32, torch.bfloat16) or not g.is_contiguous():
                g = g.float().contiguous()
            self._gbf16[k] = 1 if g.dtype == torch.bfloat16 else 0
            live.append((p, g))  # keeps a converted gradient alive until the launch is enqueued
            sh = self.shadows.get(p)
            self._arr["p"][k], self._arr["g"][k] = p.data_ptr(), g.data_ptr()
            self._arr["m"][k], self._arr["v"][k] = self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr()
            self._arr["s"][k] = sh[0].data_ptr() if sh is not None else None
            self._numel[k] = p.numel()
            k += 1
        if k == 0:
            return
        st = _stream()
        if tick:
            check(lib.shb_adam_tick(self.step_count.data_ptr(), st), "shb_adam_tick")
        check(lib.shb_adam_step_mixed(k, self._arr["p"], self._arr["g"], self._gbf16, self._arr["m"], self._arr["v"],
                                      self._arr["s"], self._numel, self.step_count.data_ptr(), self.lr, self.betas[0],
                                      self.betas[1], self.eps, self.weight_decay, st), "shb_adam_step_mixed")
        _count((1 if tick else 0) + (k + 31) // 32)
        for p, _ in live:  # the kernel wrote through raw pointers: tell autograd / the shadow cache
            torch.autograd.graph.increment_version(p)
            sh = self.shadows.get(p)
            if sh is not None:
                sh[1] = p._version

    # ---- torch.optim.Adam-compatible state
    def state_dict(self):
        t = self.step_count.detach().clone().reshape(())
        state = {i: {"step": t.clone(), "exp_avg": self.exp_avg[i], "exp_avg_sq": self.exp_avg_sq[i]}
                 for i in range(len(self.params))}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.params):
            raise ValueError("optimizer state does not match the parameter list")
        self.lr, self.eps, self.weight_decay = float(group["lr"]), float(group["eps"]), float(group["weight_decay"])
        self.betas = (float(group["betas"][0]), float(group["betas"][1]))
        steps = set()
        for i, pid in enumerate(group["params"]):
            st = sd["state"].get(pid)
            if st is None:
                continue
            self.exp_avg[i].copy_(st["exp_avg"])
            self.exp_avg_sq[i].copy_(st["exp_avg_sq"])
            steps.add(float(st["step"]))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; this optimizer keeps one")
        self.step_count.fill_(steps.pop() if steps else 0.0)
