"""Training-step driver for the plain SpiralAutoencoder (train_funcs.py:495-513 as one callable).

    step = TrainStep(model, lr=1e-3, weight_decay=5e-5)     # Adam as main.py:262
    loss = step(x_device)                                    # fwd + l1 + bwd (+ grad all-reduce) + Adam; 0-d tensor

The reference loop calls optim.zero_grad / model / loss_fn / backward / optim.step itself and can keep doing so
with the drop-in model; this class is the same sequence packaged for the benchmark and for data-parallel runs
(gradient buckets + overlapped all-reduce from dp.GradSync), with pinned-host staging for the end-to-end path and,
optional CUDA-graph replay of the whole step (shapes are static; ~110 launches, the bucket all-reduces and their
Python glue collapse into one graph launch per rank).
"""
import torch

from . import functions as fn
from .dp import GradSync
from .optim import Adam


class TrainStep:
    def __init__(self, model, lr=1e-3, weight_decay=5e-5, optimizer=True, graph=False, comm_sms=0, grad_comm_dtype="auto",
                 fused_loss=True):
        """`fused_loss`: take the L1 loss through model.reconstruction_loss (slab-side loss, no row-major reconstruction) when
        the model offers it; False runs model(x) + functions.l1_loss, the reference loop's two calls.
        `grad_comm_dtype`: dtype of the gradient sinks of the two FC weights (dp.GradSync) --
        "auto": bfloat16 when the model computes in bf16 (the compute dtype must be set before this constructor), none in
        fp32 mode (every gradient travels in fp32 through p.grad); or torch.float32 / torch.bfloat16 / None explicitly."""
        self.model = model
        self.fused_loss = bool(fused_loss) and hasattr(model, "reconstruction_loss")
        if grad_comm_dtype == "auto":
            grad_comm_dtype = torch.bfloat16 if getattr(model, "compute_dtype", None) == torch.bfloat16 else None
        if not optimizer and not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            grad_comm_dtype = None   # gradients are the product (tests, inspection): leave them in p.grad
        self.graph_enabled = bool(graph)
        # The captured step runs on a side stream.  autograd's AccumulateGrad nodes (created, and kept alive, by the bucket
        # hooks GradSync registers) remember the stream that was current at their creation and run there: created under the
        # default stream they would fork every gradient accumulation -- and the all-reduce launched from its hook -- off the
        # capture stream and join again at the end of backward.  So the hooks are registered under the capture stream.
        self._side = torch.cuda.Stream() if (self.graph_enabled and next(model.parameters()).is_cuda) else None
        if self._side is not None:
            with torch.cuda.stream(self._side):
                self.sync = GradSync(model, comm_sms=comm_sms, sink_dtype=grad_comm_dtype)
        else:
            self.sync = GradSync(model, comm_sms=comm_sms, sink_dtype=grad_comm_dtype)
        # own multi-tensor Adam (device-side step count: replayable; writes the bf16 weight shadows in the same pass)
        shadows = model.shadow_map() if hasattr(model, "shadow_map") else None
        self.optim = Adam(model.parameters(), lr=lr, weight_decay=weight_decay, shadows=shadows) if optimizer else None
        # Early optimizer launches: a gradient sink (the FC weights: 99 % of the parameters) is final in the middle of the
        # backward pass, and nothing after its own backward node reads that weight again.  Its Adam update (HBM-bound) therefore
        # runs on a side stream BESIDE the encoder backward (shared-memory-port-bound), after the sink's all-reduce where there
        # is one; the launch after backward only covers the remaining parameters.
        self._opt_stream = None
        self._early = []
        if self.optim is not None and self.sync.sinks:
            self._opt_stream = torch.cuda.Stream()
            self.sync.on_sink_ready = self._early_step
        self._copy_stream = None
        self._staged = None
        self._graph = None
        self._gx = None
        self._gloss = None
        self.launches_per_step = None

    def _early_step(self, param, grad, work):
        """GradSync callback (runs inside backward, on the stream of the node that produced `grad`)."""
        if self.optim is None:   # optimizer removed after construction: just join the collective here
            if work is not None:
                work.wait()
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._opt_stream):
            self._opt_stream.wait_event(ev)      # the GEMM that wrote the sink (and everything before it, incl. the tick)
            if work is not None:
                work.wait()                      # ... and its all-reduce
            self.optim.step(grads={param: grad}, only=[param], tick=False)
        self._early.append(param)

    def _eager(self, x):
        self.sync.reset()
        self._early = []
        split = self._opt_stream is not None
        if split:
            self.optim.tick()   # before anything of this step: every optimizer launch below reads the advanced count
        if self.fused_loss:
            loss = self.model.reconstruction_loss(x)   # == the two lines below, without the row-major reconstruction
        else:
            xh, _z = self.model(x)
            loss = fn.l1_loss(x, xh)  # train_funcs.py:501  loss_fn(tx, tx_hat)
        loss.backward()
        self.sync.finish()
        if self.optim is not None:
            if split:
                torch.cuda.current_stream().wait_stream(self._opt_stream)
                self.optim.step(grads=self.sync.grad_map(), skip=self._early, tick=False)
            else:
                self.optim.step(grads=self.sync.grad_map())
        return loss

    def capture(self, example_x, warmup=3):
        """Capture one whole step into a CUDA graph (static input buffer; later calls copy into it and replay)."""
        if not self.graph_enabled:
            raise RuntimeError("construct TrainStep(graph=True) to capture")
        self._gx = example_x.clone()
        # warm-up and capture share one side stream: autograd's AccumulateGrad nodes (kept alive by the bucket hooks in
        # multi-process runs) remember the stream they were created on, and a mismatch would make the engine
        # synchronise with the default stream in the middle of the capture
        side = self._side
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager(self._gx)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = fn.LAUNCHES["n"]
        g = torch.cuda.CUDAGraph()
        # multi-process: the bucket all-reduces are captured too (NCCL forks/joins its stream inside the graph); the
        # process group's watchdog thread polls events concurrently, hence thread-local capture checking there
        mode = "global" if self.sync.world == 1 else "thread_local"
        with torch.cuda.graph(g, stream=side, capture_error_mode=mode):
            self._gloss = self._eager(self._gx)
        self.launches_per_step = fn.LAUNCHES["n"] - n0
        self._graph = g
        return self

    def release(self):
        """Drop the captured graph (back to eager).  Must run before dist.destroy_process_group(): NCCL will not
        tear a communicator down while a live graph still holds its kernels."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._gloss = None
            torch.cuda.synchronize()

    def __call__(self, x):
        if self._graph is not None and fn.TIMER is None and x.shape == self._gx.shape and x.dtype == self._gx.dtype:
            self._gx.copy_(x, non_blocking=True)
            self._graph.replay()
            fn.LAUNCHES["n"] += self.launches_per_step  # the replay launches the captured kernels again
            return self._gloss
        return self._eager(x)

    # ---- end-to-end path: inputs start in pinned host memory, the loss ends in host memory
    def stage(self, x_host_pinned):
        """Start the H2D copy of the next batch on a side stream (overlaps the current step)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        with torch.cuda.stream(self._copy_stream):
            xd = x_host_pinned.to(next(self.model.parameters()).device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = (xd, ev)

    def step_staged(self, loss_host_pinned):
        """Run one step on the staged batch; the loss is copied to pinned host memory asynchronously."""
        xd, ev = self._staged
        torch.cuda.current_stream().wait_event(ev)
        xd.record_stream(torch.cuda.current_stream())
        loss = self(xd)
        loss_host_pinned.copy_(loss.detach().reshape(loss_host_pinned.shape), non_blocking=True)
        return loss


class BoneGuidedStep:
    """One training step of the bone-guided model (train_funcs.py:128-392, edit_mode 'equal', exc_mode 'ori_m', w_part_mode
    '1/K'): three passes through the model -- reconstruction, interpolation (non-leaf part codes scaled by a random factor),
    exchange (keypoints swapped across the batch) -- six loss terms (reconstruction L1, part-measure latent loss, keypoint L1
    and the orientation-adaptive pairwise-distance loss for both edited passes), ONE backward, Adam; optionally replayed as one
    CUDA graph.  The per-sample edge / volume regularisers of :136-143, :322-332 are separate torch functions
    (aux_losses.py) and not part of this step.

        step = BoneGuidedStep(model, J_regressor, kps_keep, parts, skl_list, P, Q)
        loss = step(tx, tx_interp, tx_exc, measure)
    """

    DEFAULT_WEIGHTS = {"rec": 1.0, "zpartreg": 1e-2, "interp_kps": 1.0, "interp_euc": 1e-2, "exc_kps": 1.0, "exc_euc": 1e-2}

    def __init__(self, model, J_regressor, kps_keep, parts, skl_list, P, Q, leaf_parts=(0, 7, 10, 13, 16), weights=None,
                 w_mode="linear", w_threshold=0.8, relative=True, factor=(0.4, 0.8), lr=1e-3, weight_decay=5e-5,
                 optimizer=True, graph=False):
        dev = next(model.parameters()).device
        self.model = model
        self.J = torch.as_tensor(J_regressor, dtype=torch.float32, device=dev).contiguous()
        self.keep = torch.as_tensor(kps_keep, dtype=torch.long, device=dev)
        self.P, self.Q = list(P), list(Q)
        self._P_dev = torch.as_tensor(self.P, dtype=torch.long, device=dev)
        self._pn_index = None  # validated (P, Q) lists on the device, built at the first call (needs the measure width)
        self.layout = fn.PairLossLayout(parts, skl_list, dev, w_mode=w_mode, leaf_parts=leaf_parts)
        self.n_parts = len(parts)
        self.weights = dict(self.DEFAULT_WEIGHTS if weights is None else weights)
        self.w_threshold, self.relative, self.factor = float(w_threshold), bool(relative), factor
        self.optim = Adam(model.parameters(), lr=lr, weight_decay=weight_decay) if optimizer else None
        self.graph_enabled = bool(graph)
        self._graph = self._static = self._gloss = None
        self.terms = {}

    def _kps(self, v):
        return torch.matmul(self.J, v[:, :-1, :])

    def loss(self, tx, tx_interp, tx_exc, measure, factor=None):
        m, w = self.model, self.weights
        B = tx.shape[0]
        t = {}
        tx_hat, z, _ = m(tx, self._kps(tx)[:, self.keep])
        t["rec"] = fn.l1_loss(tx, tx_hat)
        if self._pn_index is None:
            self._pn_index = fn.PartNormIndex(self.P, self.Q, z.shape[1], measure.shape[1], z.device)
        t["zpartreg"] = fn.partnorm_loss(z, measure, self._pn_index, None, self.relative)
        # interpolation pass (train_funcs.py:213-228): one factor for every non-leaf part code
        if factor is None:
            factor = torch.rand(1, device=tx.device) * self.factor[0] + self.factor[1]
        factor = torch.as_tensor(factor, dtype=torch.float32, device=tx.device).reshape(1)
        scale = torch.ones(B, self.n_parts, device=tx.device)
        scale[:, self._P_dev] = factor
        kps_i = self._kps(tx_interp)
        new_kps = kps_i[:, self.keep]
        lat, lat_k, dummy = m.encode(tx_interp, new_kps)
        rec_i = m.decode(lat * scale[:, :, None], lat_k, dummy)
        t["interp_kps"] = fn.l1_loss(self._kps(rec_i)[:, self.keep].contiguous(), new_kps.contiguous())
        t["interp_euc"] = fn.pair_loss(tx_interp[:, :-1, :], rec_i[:, :-1, :], kps_i, self.layout, scale=scale,
                                       w_threshold=self.w_threshold, relative=self.relative)
        # exchange pass (:296-300, :319): keypoints of the mirror sample
        kps_e = self._kps(tx_exc)
        new_kps_e = torch.flip(kps_e, dims=[0])[:, self.keep]
        lat, lat_k, dummy = m.encode(tx_exc, new_kps_e)
        rec_e = m.decode(lat, lat_k, dummy)
        t["exc_kps"] = fn.l1_loss(self._kps(rec_e)[:, self.keep].contiguous(), new_kps_e.contiguous())
        t["exc_euc"] = fn.pair_loss(tx_exc[:, :-1, :], rec_e[:, :-1, :], kps_e, self.layout, scale=None,
                                    w_threshold=self.w_threshold, relative=self.relative)
        self.terms = t
        total = None
        for k, v in t.items():
            total = w[k] * v if total is None else total + w[k] * v
        return total

    def _eager(self, tx, tx_interp, tx_exc, measure, factor=None):
        for p in self.model.parameters():
            p.grad = None
        loss = self.loss(tx, tx_interp, tx_exc, measure, factor)
        loss.backward()
        if self.optim is not None:
            self.optim.step()
        return loss

    def capture(self, tx, tx_interp, tx_exc, measure, warmup=3):
        if not self.graph_enabled:
            raise RuntimeError("construct BoneGuidedStep(graph=True) to capture")
        self._static = [t.clone() for t in (tx, tx_interp, tx_exc, measure)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager(*self._static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = fn.LAUNCHES["n"]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._gloss = self._eager(*self._static)
        self.launches_per_step = fn.LAUNCHES["n"] - n0
        self._graph = g
        return self

    def __call__(self, tx, tx_interp, tx_exc, measure, factor=None):
        if self._graph is not None and factor is None and fn.TIMER is None:
            for dst, src in zip(self._static, (tx, tx_interp, tx_exc, measure)):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
            fn.LAUNCHES["n"] += self.launches_per_step
            return self._gloss
        return self._eager(tx, tx_interp, tx_exc, measure, factor)
