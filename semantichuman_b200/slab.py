"""Slab-layout trunk operators: the autograd layer over libshb200's shb_slab_* entry points.

Inside the model trunks activations live batch-innermost in 128-sample chunks ("slab layout", csrc/shb_slab.cuh):
what SpiralConv gathers for one output vertex (``x[:, spiral_idx]``, models.py:42) is then S contiguous slabs, each
moved by one TMA bulk copy straight into tensor-core operand form; Pool (models.py:127,148) is a weighted sum of whole
slabs.  A :class:`Slab` wraps one such tensor together with what its consumers need to know about its producer.

Gradient convention inside a trunk (private to this module): the gradient that flows back INTO a slab tensor is the
gradient w.r.t. the producer's *pre-activation*.  Every consumer's backward kernel multiplies by act'(y) of the producer
(read from the saved output y) and applies the producer's dummy-row mask in its epilogue, so no stand-alone
activation-derivative pass over the gradients exists.  ``from_rows`` / ``to_rows`` are the only ways in and out, and they
translate between this convention and ordinary autograd gradients.

planes = 1: bf16 activations and operands (bf16 mode, 2e-2 parity).  planes = 2: every activation and weight is the
sum of two bf16 numbers (hi + lo, ~16 mantissa bits) and every product runs as four tensor-core MMAs
(hi.hi + lo.hi + hi.lo + lo.lo, fp32 accumulation) -- the fp32 mode, 1e-4 parity (hard part 1).
"""
import numpy as np
import torch

from . import functions as fn
from ._capi import ACT_ENUM, BF16, F32, lib
from .functions import _call, _count, _cuda, _p, _stream
from .indexing import SpiralGeometry

SlabGeometry = SpiralGeometry  # the kernels' geometry class (kept under its first name for the tests and scripts)

CHUNK = 128


def pad_channels(c):
    """Channel count of the slab tensor that carries c channels: 8, 16, 32, 64, 128, 256 (zeros above c)."""
    for p in (8, 16, 32, 64, 128, 256):
        if c <= p:
            return p
    raise ValueError(f"slab kernels support at most 256 channels per layer, got {c}")


def _pad16(c):
    return (c + 15) // 16 * 16


class Slab:
    """A slab tensor plus what the consumers of its gradient need: the activation that produced it (its derivative is taken
    through the stored output) and whether the producer zeroed the dummy row (models.py:48-51)."""

    __slots__ = ("t", "rows", "B", "C", "Cp", "planes", "act", "masked")

    def __init__(self, t, rows, B, C, Cp, planes, act=0, masked=False):
        self.t, self.rows, self.B, self.C, self.Cp, self.planes, self.act, self.masked = t, rows, B, C, Cp, planes, act, masked

    @staticmethod
    def shape_for(rows, B, Cp, planes):
        return (rows, (B + CHUNK - 1) // CHUNK, planes, Cp // 8, CHUNK, 8)

    @staticmethod
    def empty(rows, B, C, Cp, planes, device, act=0, masked=False):
        t = torch.empty(Slab.shape_for(rows, B, Cp, planes), dtype=torch.bfloat16, device=device)
        return Slab(t, rows, B, C, Cp, planes, act, masked)

    def like(self, t):
        return Slab(t, self.rows, self.B, self.C, self.Cp, self.planes, self.act, self.masked)

    @property
    def esize_bytes(self):
        return 2 * self.planes


def _dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"semantichuman_b200 supports float32 and bfloat16 tensors, got {t.dtype}")


def inverse_perm(perm):
    """perm_inv[perm[i]] = i, made once per permutation tensor and kept on it (the 3-channel conversions walk the caller's
    tensor in the caller's order and need the inverse map; built eagerly by the models, so never inside a graph capture)."""
    if perm is None:
        return None
    inv = getattr(perm, "_shb_inv", None)
    if inv is None or inv.device != perm.device:
        inv = torch.empty_like(perm)
        inv[perm.long()] = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
        perm._shb_inv = inv
    return inv


def _from_rows_raw(x, perm, Cp, planes, ymul=None, act_mul=0, zero_last=False):
    B, R, Cs = x.shape
    out = torch.empty(Slab.shape_for(R, B, Cp, planes), dtype=torch.bfloat16, device=x.device)
    _call(f"slab_from_rows[{R}x{Cs}>{Cp}]", {"bytes": float(B) * R * (Cs * x.element_size() + Cp * 2 * planes)},
          lib.shb_slab_from_rows, _p(x), _dt(x), _p(perm), _p(inverse_perm(perm)), _p(out), _p(ymul), B, R, Cs, Cp, int(act_mul),
          int(bool(zero_last)),
          planes, _stream())
    _count()
    return out


def _to_rows_raw(t, perm, B, R, Cp, Cd, planes, dtype):
    out = torch.empty((B, R, Cd), dtype=dtype, device=t.device)
    _call(f"slab_to_rows[{R}x{Cp}>{Cd}]", {"bytes": float(B) * R * (Cd * out.element_size() + Cp * 2 * planes)},
          lib.shb_slab_to_rows, _p(t), _p(perm), _p(inverse_perm(perm)), _p(out), _dt(out), B, R, Cp, Cd, planes, _stream())
    _count()
    return out


class FromRowsFn(torch.autograd.Function):
    """Caller layout (B, rows, C) -> slab layout (vertex permutation, channel padding, dtype conversion in one pass)."""

    @staticmethod
    def forward(ctx, x, perm, Cp, planes):
        _cuda(x)
        x = x.contiguous()
        ctx.perm, ctx.meta, ctx.xdtype = perm, (x.shape[0], x.shape[1], x.shape[2], Cp, planes), x.dtype
        return _from_rows_raw(x, perm, Cp, planes)

    @staticmethod
    def backward(ctx, g):
        B, R, Cs, Cp, planes = ctx.meta
        return _to_rows_raw(g.contiguous(), ctx.perm, B, R, Cp, Cs, planes, ctx.xdtype), None, None, None


class ToRowsFn(torch.autograd.Function):
    """Slab layout -> caller layout (B, rows, C).  Backward re-enters the trunk's gradient convention: the incoming
    gradient is multiplied by act'(y) of the slab's producer and masked."""

    @staticmethod
    def forward(ctx, t, perm, meta, dtype):
        rows, B, C, Cp, planes, act, masked = meta
        ctx.perm, ctx.meta = perm, meta
        ctx.save_for_backward(t)
        return _to_rows_raw(t, perm, B, rows, Cp, C, planes, dtype)

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        rows, B, C, Cp, planes, act, masked = ctx.meta
        g = g.contiguous()
        ymul = t if act != 0 else None
        return _from_rows_raw(g, ctx.perm, Cp, planes, ymul, act, masked), None, None, None


def from_rows(x, perm=None, planes=1, Cp=None):
    """(B, rows, C) float32/bfloat16 CUDA tensor -> Slab (internal row i = caller's row perm[i])."""
    B, R, C = x.shape
    Cp = pad_channels(C) if Cp is None else Cp
    return Slab(FromRowsFn.apply(x, perm, Cp, planes), R, B, C, Cp, planes, 0, False)


def to_rows(s, perm=None, dtype=torch.float32):
    return ToRowsFn.apply(s.t, perm, (s.rows, s.B, s.C, s.Cp, s.planes, s.act, s.masked), dtype)


class SlabL1LossFn(torch.autograd.Function):
    """mean |rec - target| with rec still a slab tensor (the last decoder SpiralConv's output) and target the caller's row-major
    (B, rows, C) tensor: stands in for to_rows + F.l1_loss (train_funcs.py:501) and, on the way back, for the L1 backward +
    from_rows -- the row-major reconstruction and its gradient are never written.  The gradient it returns is already in the
    trunk's convention (times act' of the producer, producer's dummy-row mask applied)."""

    @staticmethod
    def forward(ctx, t, target, perm, meta):
        rows, B, C, Cp, planes, act, masked = meta
        _cuda(target)
        if Cp != 8:
            raise NotImplementedError("slab l1_loss: the reconstruction must be an 8-channel (<= 8 real channels) slab tensor")
        if tuple(target.shape) != (B, rows, C):
            raise ValueError(f"slab l1_loss: target must be ({B}, {rows}, {C}), got {tuple(target.shape)}")
        target = target.contiguous()
        ctx.perm, ctx.meta = perm, meta
        ctx.save_for_backward(t, target)
        nbytes = lib.shb_slab_l1_workspace()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
        loss = torch.empty((), dtype=torch.float32, device=t.device)
        _call(f"slab_l1_fwd[{rows}x{C}]", {"bytes": float(B) * rows * (C * target.element_size() + Cp * 2 * planes)},
              lib.shb_slab_l1_fwd, _p(t), _p(target), _dt(target), _p(inverse_perm(perm)), _p(ws), nbytes, _p(loss), B, rows, C,
              planes, _stream())
        _count(2)
        return loss

    @staticmethod
    def backward(ctx, g):
        t, target = ctx.saved_tensors
        rows, B, C, Cp, planes, act, masked = ctx.meta
        gs = g.to(torch.float32).contiguous()
        gt = torch.empty_like(t)
        _call(f"slab_l1_bwd[{rows}x{C}]", {"bytes": float(B) * rows * (C * target.element_size() + 2 * Cp * 2 * planes)},
              lib.shb_slab_l1_bwd, _p(t), _p(target), _dt(target), _p(inverse_perm(ctx.perm)), _p(gs), _p(gt), B, rows, C, int(act),
              int(bool(masked)), planes, _stream())
        _count()
        return gt, None, None, None


def l1_loss(s, target, perm=None):
    """F.l1_loss(to_rows(s, perm), target) without the row-major reconstruction (target: (B, rows, C), no gradient)."""
    if target.requires_grad:
        raise NotImplementedError("slab l1_loss: the target takes no gradient; use to_rows + functions.l1_loss")
    return SlabL1LossFn.apply(s.t, target, perm, (s.rows, s.B, s.C, s.Cp, s.planes, s.act, s.masked))


class SlabPoolFn(torch.autograd.Function):
    """y[r] = sum_k P[r,k] x[k] over whole slabs (models.py:127,148); backward = P^T, times act' of x's producer."""

    @staticmethod
    def forward(ctx, t, pm, meta):
        rows, B, C, Cp, planes, act, masked = meta
        if rows != pm.rows_in:
            raise ValueError(f"pool expects {pm.rows_in} rows, got {rows}")
        ctx.pm, ctx.meta = pm, meta
        ctx.save_for_backward(t)
        y = torch.empty(Slab.shape_for(pm.rows_out, B, Cp, planes), dtype=torch.bfloat16, device=t.device)
        _call(f"slab_pool[{pm.rows_in}>{pm.rows_out}x{C}]", fn._pool_meta(B, pm, Cp, 2 * planes, False), lib.shb_slab_pool, _p(t),
              _p(pm.rowptr), _p(pm.colidx), _p(pm.vals), _p(y), None, B, pm.rows_out, Cp, 0, 0, planes, _stream())
        _count()
        return y

    @staticmethod
    def backward(ctx, g):
        (t,) = ctx.saved_tensors
        pm = ctx.pm
        rows, B, C, Cp, planes, act, masked = ctx.meta
        gx = torch.empty_like(t)
        _call(f"slab_pool_bwd[{pm.rows_out}>{pm.rows_in}x{C}]", fn._pool_meta(B, pm, Cp, 2 * planes, True), lib.shb_slab_pool,
              _p(g.contiguous()), _p(pm.t_rowptr), _p(pm.t_colidx), _p(pm.t_vals), _p(gx), _p(t) if act != 0 else None, B,
              pm.rows_in, Cp, act, int(masked), planes, _stream())
        _count()
        return gx, None, None


def pool(s, pm):
    t = SlabPoolFn.apply(s.t, pm, (s.rows, s.B, s.C, s.Cp, s.planes, s.act, s.masked))
    return Slab(t, pm.rows_out, s.B, s.C, s.Cp, s.planes, 0, False)


def _gconv_program(geom, backward, S, Cs, Cd, planes, B):
    """Group program for one pass if the grouped kernel (csrc/shb_slab_gconv.cu) takes this layer shape, else None.
    Group size: the largest the accumulator allows (TMEM: 2 R pad16(Cd) <= 512 columns, 256 when two CTAs share an SM) that
    still leaves every CTA a few tiles -- a group is one tile per 128-sample chunk."""
    import ctypes

    if not GCONV_ENABLED:
        return None
    rmax, sps = ctypes.c_int(0), ctypes.c_int(0)
    if lib.shb_slab_gconv_plan(S, Cs, Cd, planes, ctypes.addressof(rmax), ctypes.addressof(sps)) != 0:
        return None
    R = rmax.value
    rows_dst = geom.rows_in if backward else geom.rows_out
    nb = (B + CHUNK - 1) // CHUNK
    while R > 1 and (rows_dst + R - 1) // R * nb < _GCONV_MIN_TILES:
        R //= 2
    if R < 4:   # little sharing left to exploit, and the row-per-tile kernel has the leaner epilogue
        return None
    return geom.group_program(backward, R, sps.value)


_GCONV_MIN_TILES = 592  # two tiles for each of 296 resident CTAs
GCONV_ENABLED = True     # scripts/bench_slab_layer.py flips this to time the row-per-tile kernel on the same layers


def _conv_pass(name, cmeta, geom, backward, src, img, bias, dst, ymul, B, rows_dst, S, Cs, Cd, Cd_real, act, act_mul,
               zero_last, planes):
    """One launch of the conv kernel family: the grouped kernel where it applies, else the row-per-tile kernel."""
    prog = _gconv_program(geom, backward, S, Cs, Cd, planes, B)
    if prog is not None:
        _call(name, cmeta, lib.shb_slab_gconv, _p(src), _p(prog.gptr), _p(prog.recs), _p(prog.gdst), _p(prog.gmask),
              prog.n_groups, prog.R, prog.SPS, _p(img), _p(bias), _p(dst), _p(ymul), B, rows_dst, S, Cs, Cd, Cd_real, act,
              act_mul, zero_last, planes, _stream())
    else:
        ptr, ent = (geom.ptr_b, geom.ent_b) if backward else (geom.ptr_f, geom.ent_f)
        _call(name, cmeta, lib.shb_slab_conv, _p(src), _p(ptr), _p(ent), _p(img), _p(bias), _p(dst), _p(ymul), B, rows_dst, S,
              Cs, Cd, Cd_real, act, act_mul, zero_last, planes, _stream())


class SlabConvFn(torch.autograd.Function):
    """y = mask * act(W . gather(x) + b) (models.py:34-53) on slab tensors.

    forward : weight-image kernel + ONE fused gather-GEMM kernel (TMA slabs -> tcgen05 -> bias/act/mask epilogue).
    backward: weight/bias gradient kernel (+ fixed-order reduce) and the input-gradient kernel, which also applies act' and
              the mask of x's producer.  The incoming gradient is already w.r.t. this layer's pre-activation (see module
              docstring)."""

    @staticmethod
    def forward(ctx, t, weight, bias, geom, act, meta, want_gx, images=None):
        rows, B, C, Cp, planes, xact, xmasked = meta
        _cuda(t, weight, bias)
        cout, k = weight.shape
        S = geom.S
        if rows != geom.rows_in or k != S * C:
            raise ValueError(f"shape mismatch: x rows {rows} channels {C}, weight {tuple(weight.shape)}, geometry "
                             f"rows_in={geom.rows_in} S={S}")
        if geom.table.device != t.device:
            raise RuntimeError("spiral tables live on a different device than x")
        cout_p = pad_channels(cout)
        if not lib.shb_slab_conv_supported(S, Cp, cout_p, planes) or not lib.shb_slab_conv_supported(S, cout_p, Cp, planes) \
                or not lib.shb_slab_wgrad_supported(S, Cp, cout_p, planes):
            raise NotImplementedError(f"SpiralConv shape S={S}, {C}->{cout} channels is outside what the slab kernels support")
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        dev = t.device
        if images is not None:   # kept current by the model (WeightImages): one launch for all layers, not one per call
            img_f, img_b = images
        else:
            img_f = torch.empty(lib.shb_slab_weight_image_bytes(S, Cp, cout_p, planes), dtype=torch.uint8, device=dev)
            img_b = torch.empty(lib.shb_slab_weight_image_bytes(S, cout_p, Cp, planes), dtype=torch.uint8, device=dev)
            _call("slab_weight_images", {"bytes": 4.0 * w32.numel() + img_f.numel() + img_b.numel()}, lib.shb_slab_weight_images,
                  _p(w32), _p(img_f), _p(img_b), S, C, cout, Cp, cout_p, planes, _stream())
            _count()
        y = torch.empty(Slab.shape_for(geom.rows_out, B, cout_p, planes), dtype=torch.bfloat16, device=dev)
        tag = f"[{rows}>{geom.rows_out}x{S}x{C}>{cout}]"
        cmeta = fn._conv_meta(B, rows, geom.rows_out, S, C, cout, 2 * planes)
        _conv_pass("slabconv_fwd" + tag, cmeta, geom, False, t, img_f, b32, y, None, B, geom.rows_out, S, Cp, cout_p, cout, act,
                   0, int(geom.zero_last_row), planes)
        _count(2)
        ctx.save_for_backward(t, img_b)
        ctx.geom, ctx.meta, ctx.tag, ctx.cmeta = geom, meta, tag, cmeta
        ctx.dims = (cout, cout_p, bias is not None, weight.dtype, bool(want_gx))
        return y

    @staticmethod
    def backward(ctx, gz):
        t, img_b = ctx.saved_tensors
        geom = ctx.geom
        rows, B, C, Cp, planes, xact, xmasked = ctx.meta
        cout, cout_p, has_bias, wdtype, want_gx = ctx.dims
        S = geom.S
        gz = gz.contiguous()
        dev = t.device
        gw = gb = gx = None
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            nbytes = lib.shb_slab_wgrad_workspace(S, Cp, cout_p, planes)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            gw = torch.empty((cout, S * C), dtype=torch.float32, device=dev)
            gb = torch.empty((cout,), dtype=torch.float32, device=dev) if (has_bias and ctx.needs_input_grad[2]) else None
            _call("slabconv_wgrad" + ctx.tag, ctx.cmeta, lib.shb_slab_wgrad, _p(t), _p(geom.table), _p(gz), _p(gw), _p(gb),
                  _p(ws), nbytes, B, geom.rows_out, S, C, Cp, cout, cout_p, int(geom.zero_last_row),
                  geom.rows_in - 1 if geom.src_dummy_zero else -1, planes, _stream())
            _count(2)
            if wdtype != torch.float32:
                gw = gw.to(wdtype)
            if not ctx.needs_input_grad[1]:
                gw = None
        if want_gx and ctx.needs_input_grad[0]:
            gx = torch.empty_like(t)
            _conv_pass("slabconv_dgrad" + ctx.tag, ctx.cmeta, geom, True, gz, img_b, None, gx, t if xact != 0 else None, B, rows, S,
                       cout_p, Cp, C, 0, xact, int(xmasked), planes)
            _count()
            if geom.dummy_split is not None and not xmasked:
                T, sptr, sent, prow, pcol, pval = geom.dummy_split
                part = torch.empty(Slab.shape_for(T, B, Cp, planes), dtype=torch.bfloat16, device=dev)
                _call("slabconv_dgrad_dummy" + ctx.tag, {"bytes": 2.0 * planes * sent.numel() * cout_p * B}, lib.shb_slab_conv,
                      _p(gz), _p(sptr), _p(sent), _p(img_b), None, _p(part), None, B, T, S, cout_p, Cp, C, 0, 0, 0, planes, _stream())
                off = (rows - 1) * gx[0].numel() * 2  # byte offset of the dummy row's slabs
                _call("slabconv_dgrad_dummy_sum" + ctx.tag, {"bytes": 2.0 * planes * (T + 1) * Cp * B}, lib.shb_slab_pool,
                      _p(part), _p(prow), _p(pcol), _p(pval), _p(gx) + off, (_p(t) + off) if xact != 0 else None, B, 1, Cp, xact,
                      0, planes, _stream())
                _count(2)
        return gx, gw, gb, None, None, None, None, None


class WeightImages:
    """bf16 operand images of a set of SpiralConv weights, refreshed in ONE launch whenever a weight changed (its autograd
    version or storage): a model calls ``refresh()`` at the start of a pass and hands ``of(weight)`` to spiral_conv.  The
    images of a pass must not be overwritten before its backward has run: they are saved for the input-gradient kernel, and
    a refresh after an optimizer step writes into the same buffers -- which is the order a training step has."""

    def __init__(self, layers, planes):
        """layers: [(weight Parameter (Cout, S*Cin), S)]"""
        import ctypes

        self.planes = planes
        self.layers = list(layers)
        self._state = None
        self._img = {}
        n = len(self.layers)
        self._arr = [(ctypes.c_void_p * n)() for _ in range(3)]
        self._dims = [(ctypes.c_int * n)() for _ in range(5)]
        for k, (w, S) in enumerate(self.layers):
            cout, kdim = w.shape
            cin = kdim // S
            cp, op = pad_channels(cin), pad_channels(cout)
            img_f = torch.empty(lib.shb_slab_weight_image_bytes(S, cp, op, planes), dtype=torch.uint8, device=w.device)
            img_b = torch.empty(lib.shb_slab_weight_image_bytes(S, op, cp, planes), dtype=torch.uint8, device=w.device)
            self._img[w] = (img_f, img_b)
            self._arr[1][k], self._arr[2][k] = img_f.data_ptr(), img_b.data_ptr()
            for d, v in zip(self._dims, (S, cin, cout, cp, op)):
                d[k] = v

    def refresh(self):
        state = tuple((w.data_ptr(), w._version) for w, _ in self.layers)
        if state == self._state:
            return
        keep = []
        for k, (w, _) in enumerate(self.layers):
            w32 = w.detach()
            if w32.dtype != torch.float32 or not w32.is_contiguous():
                w32 = w32.float().contiguous()
                keep.append(w32)
            self._arr[0][k] = w32.data_ptr()
        nbytes = float(sum(4 * w.numel() + a.numel() + b.numel() for (w, _), (a, b) in zip(self.layers, self._img.values())))
        _call("slab_weight_images", {"bytes": nbytes}, lib.shb_slab_weight_images_batch, len(self.layers), self._arr[0],
              self._arr[1], self._arr[2], *self._dims, self.planes, _stream())
        _count()
        self._state = state

    def of(self, weight):
        return self._img[weight]


def spiral_conv(s, weight, bias, geom, activation="elu", want_gx=True, images=None):
    """Slab in, Slab out.  `want_gx=False` skips the input gradient (first layer of an encoder fed with data); `images`:
    prebuilt operand images of `weight` (WeightImages.of)."""
    if activation not in ACT_ENUM:
        raise NotImplementedError(activation)
    act = ACT_ENUM[activation]
    t = SlabConvFn.apply(s.t, weight, bias, geom, act, (s.rows, s.B, s.C, s.Cp, s.planes, s.act, s.masked), want_gx, images)
    cout = weight.shape[0]
    return Slab(t, geom.rows_out, s.B, cout, pad_channels(cout), s.planes, act, geom.zero_last_row)


def conv_rows(x, weight, bias, geom, activation="elu", compute_dtype=None):
    """The stand-alone SpiralConv call (models.py:34-53) on a caller-layout tensor: rows -> slabs, the fused kernel, slabs ->
    rows.  compute_dtype float32 (default for float32 input): hi/lo split operands, 1e-4 parity; bfloat16: 2e-2."""
    cdt = x.dtype if compute_dtype is None else compute_dtype
    if cdt not in (torch.float32, torch.bfloat16):
        raise TypeError(f"semantichuman_b200 supports float32 and bfloat16 activations, got {cdt}")
    _cuda(x, weight, bias)
    if x.dim() != 3:
        raise ValueError("x must be (B, V+1, C)")
    if x.shape[1] != geom.rows_in or weight.shape[1] != geom.S * x.shape[2]:
        raise ValueError(f"shape mismatch: x {tuple(x.shape)}, weight {tuple(weight.shape)}, geometry rows_in={geom.rows_in} "
                         f"S={geom.S}")
    s = from_rows(x, None, 1 if cdt == torch.bfloat16 else 2)
    return to_rows(spiral_conv(s, weight, bias, geom, activation), None, cdt)


def pool_rows(x, pm):
    """The stand-alone Pool call (torch.matmul(D|U, x), models.py:127,148) on a caller-layout tensor."""
    _cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"semantichuman_b200 supports float32 and bfloat16 activations, got {x.dtype}")
    s = from_rows(x, None, 1 if x.dtype == torch.bfloat16 else 2)
    return to_rows(pool(s, pm), None, x.dtype)
