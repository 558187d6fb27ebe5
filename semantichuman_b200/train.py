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
    def __init__(self, model, lr=1e-3, weight_decay=5e-5, optimizer=True, graph=False, comm_sms=0):
        self.model = model
        self.sync = GradSync(model, comm_sms=comm_sms)
        self.graph_enabled = bool(graph)
        # own multi-tensor Adam (device-side step count: replayable; writes the bf16 weight shadows in the same pass)
        shadows = model.shadow_map() if hasattr(model, "shadow_map") else None
        self.optim = Adam(model.parameters(), lr=lr, weight_decay=weight_decay, shadows=shadows) if optimizer else None
        self._copy_stream = None
        self._staged = None
        self._graph = None
        self._gx = None
        self._gloss = None
        self.launches_per_step = None

    def _eager(self, x):
        self.sync.reset()
        xh, _z = self.model(x)
        loss = fn.l1_loss(x, xh)  # train_funcs.py:501  loss_fn(tx, tx_hat)
        loss.backward()
        self.sync.finish()
        if self.optim is not None:
            self.optim.step()
        return loss

    def capture(self, example_x, warmup=3):
        """Capture one whole step into a CUDA graph (static input buffer; later calls copy into it and replay)."""
        if not self.graph_enabled:
            raise RuntimeError("construct TrainStep(graph=True) to capture")
        self._gx = example_x.clone()
        # warm-up and capture share one side stream: autograd's AccumulateGrad nodes (kept alive by the bucket hooks in
        # multi-process runs) remember the stream they were created on, and a mismatch would make the engine
        # synchronise with the default stream in the middle of the capture
        side = torch.cuda.Stream()
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
