"""torch.autograd.Functions over the C ABI of libshb200 (include/shb200.h): losses, grouped heads, FC shadow GEMMs.
The SpiralConv / Pool operators live in slab.py.

torch is plumbing here: it owns device memory, streams and the autograd graph; every arithmetic step of the
path is a kernel of libshb200.  Tensors must be CUDA tensors -- there is no CPU path (north_star).
Forward runs on the calling thread, backward on the autograd engine's device thread; both take the stream from
``torch.cuda.current_stream()`` at call time (threading).
"""
import torch

from . import _capi
from ._capi import ACT_ENUM, check, lib

_DT = {torch.float32: _capi.F32, torch.bfloat16: _capi.BF16}

# launch counter: bench.py reports how many libshb200 kernels a step launches ("gpu_launches")
LAUNCHES = {"n": 0}


def _count(n=1):
    LAUNCHES["n"] += n


class KernelTimer:
    """Optional per-entry-point CUDA-event timing (bench.py's roofline leg).  Events are recorded on the stream the
    kernels are launched on; nothing is synchronised until summary()."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, meta, e0, e1 in self.records:
            d = out.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += meta.get("flops", 0.0)
            d["bytes"] += meta.get("bytes", 0.0)
        return out


TIMER = None  # set to a KernelTimer to time every C-ABI call


def _call(name, meta, fn_, *args):
    """Invoke one C-ABI entry point, raising on a non-zero code; optionally bracketed by CUDA events."""
    t = TIMER
    if t is None:
        check(fn_(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn_(*args), name)
    e1.record()
    t.records.append((name, meta, e0, e1))


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"semantichuman_b200 supports float32 and bfloat16 activations, got {t.dtype}") from None


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("semantichuman_b200 kernels need CUDA tensors; there is no CPU fallback")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _conv_meta(B, rows_in, rows_out, S, cin, cout, esize):
    """Algorithmic work of one SpiralConv pass in any direction: every activation row read once and
    written once; indices and weights amortised over the batch."""
    return {"flops": 2.0 * B * rows_out * S * cin * cout, "bytes": float(B) * (rows_in * cin + rows_out * cout) * esize}


def _pool_meta(B, pm, C, esize, transposed):
    """rows actually referenced + rows written, plus the CSR itself."""
    rows_read = pm.rows_out if pm.is_selection else (pm.rows_out if transposed else pm.rows_in)
    rows_written = pm.rows_in if transposed else pm.rows_out
    return {"flops": 2.0 * B * pm.nnz * C, "bytes": float(B) * (rows_read + rows_written) * C * esize + pm.nnz * 8.0}


class L1LossFn(torch.autograd.Function):
    """mean |a - b|  (F.l1_loss, train_funcs.py:135,501); fixed-order two-stage reduction."""

    @staticmethod
    def forward(ctx, a, b):
        _cuda(a, b)
        if a.shape != b.shape or a.dtype != b.dtype:
            raise ValueError("l1_loss operands must have the same shape and dtype")
        a, b = a.contiguous(), b.contiguous()
        n = a.numel()
        nbytes = lib.shb_l1_loss_workspace(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        out = torch.empty((), dtype=torch.float32, device=a.device)
        _call("l1_loss_fwd", {"bytes": 2.0 * n * a.element_size()}, lib.shb_l1_loss_fwd, _p(a), _p(b), n, _p(ws), nbytes,
              _p(out), _dt(a), _stream())
        _count(2)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.detach().float().contiguous()
        ga = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if ga is None and gb is None:
            return None, None
        nout = (ga is not None) + (gb is not None)
        _call("l1_loss_bwd", {"bytes": (2.0 + nout) * a.numel() * a.element_size()}, lib.shb_l1_loss_bwd, _p(a), _p(b),
              a.numel(), _p(g), _p(ga), _p(gb), _dt(a), _stream())
        _count()
        return ga, gb


class PartNormIndex:
    """Validated (part, measure) index lists of the part-measure latent loss, resident on the device.  Build once and pass as
    `P` to partnorm_loss (with Q=None): no per-call host validation, usable inside a CUDA-graph capture."""

    def __init__(self, P, Q, n_parts, n_measure, device):
        P = torch.as_tensor(P).detach().to("cpu", torch.int64).reshape(-1)
        Q = torch.as_tensor(Q).detach().to("cpu", torch.int64).reshape(-1)
        if P.numel() == 0 or P.numel() != Q.numel():
            raise ValueError("partnorm_loss: P and Q must be non-empty index lists of equal length")
        if int(P.min()) < 0 or int(P.max()) >= n_parts or int(Q.min()) < 0 or int(Q.max()) >= n_measure:
            raise ValueError("partnorm_loss: part / measure index out of range")
        if P.unique().numel() != P.numel():
            raise ValueError("partnorm_loss: a part may appear once in P (its gradient row is written once)")
        self.n_parts, self.n_measure = int(n_parts), int(n_measure)
        self.P = P.to(torch.int32).to(device)
        self.Q = Q.to(torch.int32).to(device)


class PartNormLossFn(torch.autograd.Function):
    """Part-measure latent loss (train_funcs.py:145-152); loss and d loss/d z in one launch."""

    @staticmethod
    def forward(ctx, z, measure, P, Q, relative):
        _cuda(z, measure)
        z = z.float().contiguous()
        measure = measure.float().contiguous()
        if z.dim() != 3 or measure.dim() != 2 or measure.shape[0] != z.shape[0]:
            raise ValueError("partnorm_loss expects z (B, n_parts, L) and measure (B, n_measure)")
        B, n_parts, L = z.shape
        # the kernel reads P / Q as int32 device arrays and writes gz rows P[i]: validated on the host (small index lists),
        # once per call or once for all in a PartNormIndex
        idx = P if isinstance(P, PartNormIndex) else PartNormIndex(P, Q, n_parts, measure.shape[1], z.device)
        if idx.n_parts != n_parts or idx.n_measure != measure.shape[1] or idx.P.device != z.device:
            raise ValueError("partnorm_loss: index lists were validated for a different shape or device")
        P, Q = idx.P, idx.Q
        out = torch.empty((), dtype=torch.float32, device=z.device)
        gz = torch.empty_like(z)
        _call("partnorm_loss_fwd_bwd", {"bytes": 2.0 * z.numel() * 4}, lib.shb_partnorm_loss_fwd_bwd, _p(z), _p(measure),
              _p(P), _p(Q), _p(out), _p(gz), B, n_parts, L, measure.shape[1], P.numel(), int(bool(relative)), _stream())
        _count()
        ctx.save_for_backward(gz)
        return out

    @staticmethod
    def backward(ctx, g):
        (gz,) = ctx.saved_tensors
        return gz * g, None, None, None, None


class GroupLayout:
    """Device-side description of a set of row groups (the body parts of the bone-guided model) and of the packed
    parameter buffers of their per-group nn.Linear layers: idx/gptr (int32), weight/bias offsets (int64)."""

    def __init__(self, groups, rows, channels, latent, gather, device):
        import numpy as np

        self.G, self.rows, self.C, self.L, self.gather = len(groups), int(rows), int(channels), int(latent), bool(gather)
        sizes = [len(g) for g in groups]
        self.sizes, self.max_rows = sizes, max(sizes)
        flat = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups])
        if flat.min() < 0 or flat.max() >= rows:
            raise ValueError("group row index out of range")
        self.disjoint = len(np.unique(flat)) == len(flat)
        self.w_numel = [n * channels * latent for n in sizes]
        self.b_numel = [latent if gather else n * channels for n in sizes]
        dev = torch.device(device)
        self.idx = torch.as_tensor(flat, dtype=torch.int32, device=dev)
        self.gptr = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
        self.woff = torch.as_tensor(np.concatenate([[0], np.cumsum(self.w_numel)[:-1]]), dtype=torch.int64, device=dev)
        self.boff = torch.as_tensor(np.concatenate([[0], np.cumsum(self.b_numel)[:-1]]), dtype=torch.int64, device=dev)

    def supported(self):
        return self.L <= 32

    def pack(self, layers):
        """Packed (weight, bias) buffers of the group's nn.Linear layers; differentiable (torch.cat)."""
        for lay, wn in zip(layers, self.w_numel):
            if lay.weight.numel() != wn:
                raise ValueError("per-group Linear shape does not match the group layout")
        return (torch.cat([lay.weight.reshape(-1) for lay in layers]), torch.cat([lay.bias.reshape(-1) for lay in layers]))


class GroupLinearGatherFn(torch.autograd.Function):
    """z[b,k,:] = Linear_k(x[b, idx_k, :].reshape(-1))  for every group k in one launch (models.py:234,252)."""

    @staticmethod
    def forward(ctx, x, wcat, bcat, lay):
        _cuda(x, wcat, bcat)
        x, wcat, bcat = x.float().contiguous(), wcat.float().contiguous(), bcat.float().contiguous()
        B = x.shape[0]
        if x.shape[1] != lay.rows or x.shape[2] != lay.C:
            raise ValueError(f"grouped linear expects (B, {lay.rows}, {lay.C}), got {tuple(x.shape)}")
        z = torch.empty((B, lay.G, lay.L), dtype=torch.float32, device=x.device)
        _call("group_linear_gather", {"bytes": 4.0 * (x.numel() + wcat.numel() + z.numel())}, lib.shb_group_linear_gather_fwd,
              _p(x), _p(lay.idx), _p(lay.gptr), _p(wcat), _p(lay.woff), _p(bcat), _p(lay.boff), _p(z), B, lay.rows, lay.C,
              lay.G, lay.L, _stream())
        _count()
        ctx.save_for_backward(x, wcat)
        ctx.lay = lay
        return z

    @staticmethod
    def backward(ctx, gz):
        x, wcat = ctx.saved_tensors
        lay = ctx.lay
        gz = gz.float().contiguous()
        B = x.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        if need_x and not lay.disjoint:
            raise RuntimeError("input gradient of a grouped linear needs non-overlapping groups")
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(wcat) if need_w else None
        gb = torch.empty(sum(lay.b_numel), dtype=torch.float32, device=x.device) if need_b else None
        _call("group_linear_gather_bwd", {"bytes": 4.0 * (2 * x.numel() + 2 * wcat.numel())}, lib.shb_group_linear_gather_bwd,
              _p(x), _p(lay.idx), _p(lay.gptr), _p(wcat), _p(lay.woff), _p(lay.boff), _p(gz), _p(gx), _p(gw), _p(gb), B,
              lay.rows, lay.C, lay.G, lay.L, lay.max_rows, _stream())
        _count(int(need_x) + int(need_w) + int(need_b))
        return gx, gw, gb, None


class GroupLinearScatterFn(torch.autograd.Function):
    """y[b, idx_k[p], :] = Linear_k(zz[b,k,:])[p*C:(p+1)*C]  for every group, plus `extra` written to the last row
    (models.py:269-273: per-part decode, permutation scatter, dummy-row concat) -- one launch."""

    @staticmethod
    def forward(ctx, zz, wcat, bcat, extra, lay):
        _cuda(zz, wcat, bcat, extra)
        zz, wcat, bcat = zz.float().contiguous(), wcat.float().contiguous(), bcat.float().contiguous()
        B = zz.shape[0]
        if zz.shape[1] != lay.G or zz.shape[2] != lay.L:
            raise ValueError(f"grouped linear expects (B, {lay.G}, {lay.L}), got {tuple(zz.shape)}")
        y = torch.zeros((B, lay.rows, lay.C), dtype=torch.float32, device=zz.device)
        _call("group_linear_scatter", {"bytes": 4.0 * (zz.numel() + wcat.numel() + y.numel())}, lib.shb_group_linear_scatter_fwd,
              _p(zz), _p(lay.idx), _p(lay.gptr), _p(wcat), _p(lay.woff), _p(bcat), _p(lay.boff), _p(y), B, lay.rows, lay.C,
              lay.G, lay.L, _stream())
        _count()
        y[:, -1:, :] = extra.to(y.dtype).expand(B, -1, -1)
        ctx.save_for_backward(zz, wcat)
        ctx.lay, ctx.extra_shape, ctx.extra_dtype = lay, extra.shape, extra.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        zz, wcat = ctx.saved_tensors
        lay = ctx.lay
        gy = gy.float().contiguous()
        B = zz.shape[0]
        need_z, need_w, need_b, need_e = ctx.needs_input_grad[:4]
        gzz = torch.empty_like(zz) if need_z else None
        gw = torch.empty_like(wcat) if (need_w or need_b) else None
        gb = torch.empty(sum(lay.b_numel), dtype=torch.float32, device=zz.device) if need_b else None
        _call("group_linear_scatter_bwd", {"bytes": 4.0 * (2 * gy.numel() + 2 * wcat.numel())}, lib.shb_group_linear_scatter_bwd,
              _p(zz), _p(lay.idx), _p(lay.gptr), _p(wcat), _p(lay.woff), _p(lay.boff), _p(gy), _p(gzz), _p(gw), _p(gb), B,
              lay.rows, lay.C, lay.G, lay.L, lay.max_rows, _stream())
        _count(int(need_z) + int(need_w or need_b))
        ge = None
        if need_e:
            ge = gy[:, -1:, :]
            if tuple(ctx.extra_shape) != tuple(ge.shape):
                ge = ge.sum_to_size(ctx.extra_shape)
            ge = ge.to(ctx.extra_dtype)
        return gzz, (gw if need_w else None), gb, ge, None


W_MODES = {"all_one": 0, "linear": 1, "sin": 2, "threshold": 3}  # cfg.TRAIN.w_mode, train_funcs.py:259-267


class PairLossLayout:
    """Device tables of the orientation-adaptive pairwise-distance loss: part vertex lists, bones, per-part weight modes."""

    def __init__(self, parts, skl_list, device, w_mode="linear", leaf_parts=(), part_weights=None):
        import numpy as np

        if w_mode not in W_MODES:
            raise NotImplementedError(w_mode)
        if len(parts) != len(skl_list):
            raise ValueError("one bone per part")
        sizes = [len(p) for p in parts]
        flat = np.concatenate([np.asarray(p, dtype=np.int64) for p in parts])
        if len(np.unique(flat)) != len(flat):
            raise ValueError("parts must not overlap")
        self.G, self.max_rows, self.n_max_vertex = len(parts), max(sizes), int(flat.max())
        self.pairs_per_sample = float(sum(n * n for n in sizes))
        bone = np.full((self.G, 3), -1, np.int32)
        for k, b in enumerate(skl_list):
            if len(b) not in (2, 3):
                raise ValueError("a bone is two keypoints, or one keypoint and the mean of two")
            bone[k, :len(b)] = b
        self.n_max_kps = int(bone.max())
        mode = np.full(self.G, W_MODES[w_mode], np.int32)
        mode[list(leaf_parts)] = 0  # leaf parts use all-one weights (train_funcs.py:259)
        pw = np.full(self.G, 1.0 / self.G, np.float32) if part_weights is None else np.asarray(part_weights, np.float32)
        dev = torch.device(device)
        self.idx = torch.as_tensor(flat, dtype=torch.int32, device=dev)
        self.gptr = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
        self.bone = torch.from_numpy(bone).to(dev)
        self.wmode = torch.from_numpy(mode).to(dev)
        self.pw = torch.from_numpy(pw).to(dev)


class PairLossFn(torch.autograd.Function):
    """train_funcs.py:243-284 as two kernels (loss, then d loss / d rec); the ground truth, keypoints and scale get no
    gradient (they are data in the reference's loop)."""

    @staticmethod
    def forward(ctx, tx, rec, kps, scale, lay, w_threshold, relative):
        _cuda(tx, rec, kps)
        tx, rec, kps = tx.float().contiguous(), rec.float().contiguous(), kps.float().contiguous()
        if tx.shape != rec.shape or tx.dim() != 3 or tx.shape[2] != 3:
            raise ValueError("tx and rec must both be (B, V, 3)")
        B, V, _ = tx.shape
        if lay.n_max_vertex >= V or lay.n_max_kps >= kps.shape[1] or kps.shape[0] != B:
            raise ValueError("part vertex / bone keypoint index out of range")
        if scale is not None:
            scale = scale.float().contiguous()
            if tuple(scale.shape) != (B, lay.G):
                raise ValueError("scale must be (B, n_parts)")
        nbytes = lib.shb_pair_loss_workspace(B, lay.G, lay.max_rows)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=tx.device)
        loss = torch.empty((), dtype=torch.float32, device=tx.device)
        # the gradient w.r.t. rec is accumulated (unscaled) by the same walk over the pairs when it will be needed
        gacc = None
        if ctx.needs_input_grad[1]:   # per tile-pair slots: every unordered pair is evaluated once, credited to both vertices
            gacc = torch.empty(lib.shb_pair_loss_grad_acc_bytes(B, lay.G, lay.max_rows) // 4, dtype=torch.float32, device=tx.device)
        pairs = float(B) * lay.pairs_per_sample
        _call("pair_loss", {"bytes": 24.0 * tx.numel() / 3, "flops": 27.5 * pairs}, lib.shb_pair_loss_fwd, _p(tx), _p(rec),
              _p(kps), _p(lay.idx), _p(lay.gptr), _p(lay.bone), _p(lay.wmode), _p(lay.pw), _p(scale), float(w_threshold),
              int(bool(relative)), _p(loss), _p(gacc), _p(ws), nbytes, B, V, kps.shape[1], lay.G, lay.max_rows, _stream())
        _count(2)
        ctx.save_for_backward(gacc if gacc is not None else tx.new_empty(0), ws)
        ctx.lay, ctx.shape = lay, (B, V)
        return loss

    @staticmethod
    def backward(ctx, g):
        gacc, ws = ctx.saved_tensors
        lay = ctx.lay
        B, V = ctx.shape
        grec = None
        if ctx.needs_input_grad[1]:
            grec = torch.empty((B, V, 3), dtype=torch.float32, device=gacc.device)
            gs = g.detach().float().contiguous()
            _call("pair_loss_bwd", {"bytes": 24.0 * B * V}, lib.shb_pair_loss_bwd, _p(gacc), _p(lay.idx), _p(lay.gptr), _p(gs),
                  _p(grec), _p(ws), ws.numel(), B, V, lay.G, lay.max_rows, _stream())
            _count()
        return None, grec, None, None, None, None, None


def pair_loss(tx, rec, kps, layout, scale=None, w_threshold=0.8, relative=True):
    """Orientation-adaptive pairwise-distance loss (train_funcs.py:243-284): `layout` = PairLossLayout(parts, skl_list, ...)."""
    return PairLossFn.apply(tx, rec, kps, scale, layout, w_threshold, relative)


def group_linear_gather(x, layers, lay):
    w, b = lay.pack(layers)
    return GroupLinearGatherFn.apply(x, w, b, lay)


def group_linear_scatter(zz, layers, extra, lay):
    w, b = lay.pack(layers)
    return GroupLinearScatterFn.apply(zz, w, b, extra, lay)


def cast_bf16(src, dst):
    """dst (bfloat16) = src (float32), one kernel; used for the FC weight shadows of the bf16 mode."""
    _cuda(src, dst)
    if src.dtype != torch.float32 or dst.dtype != torch.bfloat16 or src.numel() != dst.numel():
        raise TypeError("cast_bf16: float32 source and bfloat16 destination of equal size")
    _call("cast_bf16", {"bytes": 6.0 * src.numel()}, lib.shb_cast_bf16, _p(src.contiguous()), _p(dst), src.numel(), _stream())
    _count()
    return dst


def column_sums(g):
    """sum over the batch of a (B, N) float32 / bfloat16 CUDA tensor, in fp32 (shb_colsum): an nn.Linear's bias gradient."""
    _cuda(g)
    B, N = g.shape
    g = g.contiguous()
    out = torch.empty((N,), dtype=torch.float32, device=g.device)
    _call(f"colsum[{B}x{N}]", {"bytes": float(B) * N * g.element_size()}, lib.shb_colsum, _p(g),
          _DT[g.dtype], B, N, _p(out), _stream())
    _count()
    return out


class LinearShadowFn(torch.autograd.Function):
    """y = x W^T + b with bf16 operands taken from the weight SHADOWS (kept current by optim.Adam), gradients delivered to
    the fp32 master parameters in fp32 (cuBLAS accumulates in fp32 anyway; no bf16 gradient tensor, no cast kernels).
    The GEMMs are plain library GEMMs (models.py:129,142: nn.Linear)."""

    @staticmethod
    def forward(ctx, x, w, b, w_sh, b_sh):
        ctx.save_for_backward(x, w_sh)
        ctx.has_bias = b is not None
        ctx.w_param = w
        return torch.addmm(b_sh, x, w_sh.t()) if b is not None else torch.mm(x, w_sh.t())

    @staticmethod
    def backward(ctx, g):
        x, w_sh = ctx.saved_tensors
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.mm(g, w_sh)
        if ctx.needs_input_grad[1]:
            sink = getattr(ctx.w_param, "_shb_grad_sink", None)
            if sink is not None:
                # data-parallel run: the GEMM writes the weight gradient straight into its (fp32 or bf16) all-reduce bucket --
                # no zero-fill, no accumulate pass, no cast -- and the bucket goes on the wire at once (dp.GradSync)
                if sink.buf.dtype == g.dtype:
                    torch.mm(g.t(), x, out=sink.buf)
                else:
                    torch.mm(g.t(), x, out_dtype=sink.buf.dtype, out=sink.buf)
                sink.ready()
            else:
                gw = torch.mm(g.t(), x, out_dtype=torch.float32)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # (CPU tensors reach this function only in the gloo tests of dp.GradSync's host logic, which drive it with a toy
            # module: the models themselves refuse CPU tensors before they get here)
            gb = column_sums(g) if g.is_cuda else torch.sum(g, 0, dtype=torch.float32)
        return gx, gw, gb, None, None


def spiral_conv(x, weight, bias, geom, activation="elu", compute_dtype=None):
    """Stand-alone SpiralConv on a (B, V+1, C) tensor (see slab.conv_rows)."""
    from . import slab

    if activation not in ACT_ENUM:
        raise NotImplementedError(activation)
    return slab.conv_rows(x, weight, bias, geom, activation, compute_dtype)


def pool(x, pm):
    """Stand-alone Pool on a (B, rows, C) tensor (see slab.pool_rows)."""
    from . import slab

    return slab.pool_rows(x, pm)


def l1_loss(a, b):
    return L1LossFn.apply(a, b)


def partnorm_loss(z, measure, P, Q, relative=True):
    return PartNormLossFn.apply(z, measure, P, Q, relative)
