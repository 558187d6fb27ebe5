"""Batch-sharded data parallelism for the training step (BASELINE.json config 4).

The reference is single-device (main.py:197-201); this is the one thing the build ADDS around the path.  One
process per GPU (torchrun), every rank holds a full replica, the global batch is split evenly by sample, and the
only exchange is a gradient all-reduce (NCCL over NVLink/NVSwitch, ReduceOp.AVG == global-batch mean loss because
shards are equal).  Gradients live in four flat fp32 buckets ordered by when backward finishes them:

    decoder convs (ready first) -> fc_latent_dec (57 MB) -> fc_latent_enc (57 MB) -> encoder convs (last)

Each bucket's all-reduce is launched from an autograd post-accumulate hook the moment its last gradient lands, so
the two 57 MB FC buckets travel while the rest of the backward is still computing (the decoder-side one starts a whole
FC backward earlier than a single 114 MB bucket would).  ``finish()`` joins before the optimizer.

Gradient sinks (``sink_dtype``): a parameter the model lists in ``direct_grad_params()`` (the two FC weights in bf16 mode)
gets a bucket of its own that the weight-gradient GEMM writes DIRECTLY (functions.LinearShadowFn) -- no zero-fill of the
bucket, no read-modify-write accumulate pass by autograd (together 4x the bucket size in HBM traffic per step) -- and that
may be bf16, halving the bytes on NVLink.  Such a gradient never becomes ``p.grad`` (torch requires grad dtype == param dtype):
the optimizer takes it from ``GradSync.sinks`` (``optim.Adam.step(grads=sync.grad_map())``).  One backward per ``reset()``:
a sink is overwritten, not accumulated.
"""
import torch
import torch.distributed as dist


def shard_batch(global_batch, rank, world):
    """Even split by sample; raises if the global batch does not divide (equal shards keep AVG exact)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


N_BUCKETS = 4


def default_comm_sms(world):
    """CTAs for NCCL / SMs left free of the persistent kernels while a large bucket is in flight.  Measured on one 8 x B200 box
    (profiles/README.md, round 2 scaling): 16 is as good as NCCL's defaults at 2 and 4 ranks and 8 is worse; at 8 ranks 16 starves
    the all-reduce (2.16 ms/step) and 32 is best (2.01 ms; NCCL's defaults 2.03)."""
    return 32 if world >= 8 else 16


def bind_host_to_device(device_index):
    """Pin the calling process to the CPUs of the GPU's own NUMA node (sysfs local_cpulist of its PCI function) BEFORE it
    allocates pinned staging buffers: a process scheduled on the other socket stages every batch across the socket link, and the
    per-step host-to-device copy (21 MB at batch 256) then no longer hides behind the 1.8 ms step.  Returns the CPU set used,
    or None when the topology is not exposed or the allowed CPUs do not include any local one (nothing is changed then)."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            text = f.read().strip()
        local = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            local.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = local & allowed
        if not use or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except (OSError, AttributeError, ValueError, RuntimeError, AssertionError):  # no such device / no sysfs topology
        return None


def init_data_parallel(device, comm_sms=None):
    """Create the NCCL process group of a data-parallel run with NCCL limited to `comm_sms` CTAs (0: NCCL's defaults).
    GradSync(model, comm_sms=...) then keeps that many SMs free of the persistent SpiralConv kernels WHILE a large gradient
    bucket is in flight (measured on 2 B200: with all 148 SMs -- and all of their shared memory -- held by one-CTA-per-SM
    kernels the all-reduce cannot start beside them and its 0.3 ms stay exposed; reserving SMs for the whole step costs more
    than it hides).  comm_sms=None: default_comm_sms(world size from the environment)."""
    if comm_sms is None:
        import os

        comm_sms = default_comm_sms(int(os.environ.get("WORLD_SIZE", "1")))
    if comm_sms and comm_sms > 0:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = int(comm_sms)
        opts.config.min_ctas = min(int(comm_sms), 4)
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=device)
    return dist.get_world_size()


def _bucket_of(name):
    """Backward order: decoder convs -> decoder-side latent layers -> encoder-side latent layers -> encoder convs."""
    if name.startswith("dconv."):
        return 0
    if name.startswith("fc_latent_dec"):
        return 1
    if name.startswith(("fc_latent_enc", "kps_enc_list")):
        return 2
    return 3


class _Sink:
    """Direct destination of one parameter's gradient: the producing kernel writes `buf`, then calls ready()."""

    __slots__ = ("buf", "_sync", "_bucket")

    def __init__(self, buf, sync, bucket):
        self.buf, self._sync, self._bucket = buf, sync, bucket

    def ready(self):
        self._sync._launch(self._bucket)


class GradSync:
    """Flat-bucket gradient all-reduce overlapped with backward."""

    def __init__(self, model, process_group=None, comm_sms=0, sink_dtype=None):
        self.pg = process_group
        self.comm_sms = int(comm_sms)   # SMs kept free of persistent kernels while a large bucket is being reduced
        self._capped = False
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # NCCL averages inside the collective; gloo (CPU tests of this logic) has no AVG: sum, then scale in finish()
        self._avg_in_collective = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self._params = [p for _, p in named]
        self.buckets = []
        self._handles = []
        self.sinks = {}   # parameter -> _Sink
        for p in self._params:  # a sink left by an earlier GradSync of this model must not outlive it
            if hasattr(p, "_shb_grad_sink"):
                del p._shb_grad_sink
        self._model = model
        self.on_sink_ready = None   # callable(param, reduced-gradient tensor, all-reduce Work or None): see train.TrainStep
        direct = []
        if sink_dtype is not None and hasattr(model, "direct_grad_params"):
            if sink_dtype not in (torch.float32, torch.bfloat16):
                raise TypeError("gradient sinks are float32 or bfloat16")
            direct = [p for p in model.direct_grad_params() if p.requires_grad]
        if self.world == 1:
            # nothing to exchange: no flat buckets, gradients are plain per-parameter tensors (reset() drops them) -- except
            # the sinks, which exist for what they save locally (no accumulate pass) and for the early optimizer step
            for p in direct:
                flat = torch.zeros(p.numel(), dtype=sink_dtype, device=p.device)
                bucket = {"flat": flat, "params": [p], "pending": 0, "work": None, "sink": True}
                self.buckets.append(bucket)
                self.sinks[p] = p._shb_grad_sink = _Sink(flat.view_as(p), self, bucket)
            return
        for b in range(N_BUCKETS):
            for n, p in named:  # a sink is a bucket of its own, placed where its parameter's bucket would be
                if _bucket_of(n) == b and any(p is q for q in direct):
                    flat = torch.zeros(p.numel(), dtype=sink_dtype, device=p.device)
                    bucket = {"flat": flat, "params": [p], "pending": 0, "work": None, "sink": True}
                    self.buckets.append(bucket)
                    self.sinks[p] = p._shb_grad_sink = _Sink(flat.view_as(p), self, bucket)
            members = [(n, p) for n, p in named if _bucket_of(n) == b and not any(p is q for q in direct)]
            if not members:
                continue
            # every member starts on a 16-byte boundary (the optimizer kernel moves float4s); the padding stays zero
            total = sum((p.numel() + 3) // 4 * 4 for _, p in members)
            flat = torch.zeros(total, dtype=members[0][1].dtype, device=members[0][1].device)
            off = 0
            for _, p in members:
                p.grad = flat[off:off + p.numel()].view_as(p)  # autograd accumulates in place into the view
                off += (p.numel() + 3) // 4 * 4
            bucket = {"flat": flat, "params": [p for _, p in members], "pending": 0, "work": None, "sink": False}
            self.buckets.append(bucket)
            for _, p in members:
                p.register_post_accumulate_grad_hook(self._make_hook(bucket))
        self.reset()

    def _launch(self, bucket):
        bucket["pending"] = 0
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
            bucket["work"] = dist.all_reduce(bucket["flat"], op=op, group=self.pg, async_op=True)
            if self.comm_sms > 0 and bucket["flat"].numel() >= (1 << 20) and not self._capped:
                self._cap(True)  # kernels enqueued from here to finish() leave room for NCCL's CTAs
        if bucket["sink"] and self.on_sink_ready is not None and (self.world == 1 or self._avg_in_collective):
            p = bucket["params"][0]
            self.on_sink_ready(p, self.sinks[p].buf, bucket["work"])
            bucket["work"] = None   # the callback joined the collective on its own stream

    def _make_hook(self, bucket):
        def hook(_param):
            bucket["pending"] -= 1
            if bucket["pending"] == 0 and self.world > 1:
                self._launch(bucket)
        return hook

    def grad_map(self):
        """{parameter: reduced gradient} of the sink parameters (for optim.Adam.step(grads=...)); empty without sinks."""
        return {p: s.buf for p, s in self.sinks.items() if not s._bucket.get("inactive", False)}

    def _cap(self, on):
        from ._capi import check, lib

        check(lib.shb_set_persistent_sms(148 - self.comm_sms if on else 148), "shb_set_persistent_sms")
        self._capped = on

    def reset(self):
        """Call before each backward (instead of optimizer.zero_grad(set_to_none=True), which would drop the views)."""
        # a sink only works while its parameter's gradient is produced by a kernel that knows about sinks (the model's current
        # compute mode decides); otherwise the gradient arrives through autograd as p.grad
        live = self._model.direct_grad_params() if self.sinks else []
        for b in self.buckets:
            if b["sink"]:
                b["inactive"] = not any(b["params"][0] is q for q in live)
                if b["inactive"] and self.world > 1:
                    raise RuntimeError("the model's compute dtype changed after GradSync was built: its gradient sinks no longer "
                                       "receive their gradients; construct a new GradSync / TrainStep")
        if self.world == 1:
            for p in self._params:
                p.grad = None
            for b in self.buckets:
                b["pending"] = 0 if b["inactive"] else 1
            return
        for b in self.buckets:
            if not b["sink"]:
                b["flat"].zero_()   # autograd accumulates into these; a sink is overwritten by its producer
            b["pending"] = len(b["params"])
            b["work"] = None

    def finish(self):
        """Join the outstanding all-reduces (the current stream waits; the host does not block)."""
        if self._capped:
            self._cap(False)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
                if not self._avg_in_collective:
                    b["flat"].div_(self.world)
            elif b["pending"] != 0:
                raise RuntimeError("GradSync.finish(): a bucket never completed -- was backward run, and reset() called?")

    def grad_bytes(self):
        if self.world == 1:
            return sum((self.sinks[p].buf if p in self.sinks else p).numel() * (self.sinks[p].buf if p in self.sinks else p).element_size()
                       for p in self._params)
        return sum(b["flat"].numel() * b["flat"].element_size() for b in self.buckets)
