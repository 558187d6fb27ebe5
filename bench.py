#!/usr/bin/env python
"""Benchmark of the hot path: SpiralAutoencoder training step on 6890-vertex meshes (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype fp32|bf16] [--batch B_per_gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on the box's host cores

One step = forward + L1 loss + backward (+ gradient all-reduce when N>1) + Adam (main.py:262 hyper-parameters) over
one batch of synthetic meshes; weights are deterministic random-init; per-GPU batch 256 (weak scaling: N=8 is
BASELINE.json's global batch 2048).  Prints ONE JSON line on rank 0.  The headline is north_star's bf16 mode (bf16
operands, fp32 accumulation and master weights, parity 2e-2); the fp32 mode (bf16 hi + lo operands on the same tensor-core
kernels, parity 1e-4 -- the reference's own precision) is timed in the same run and reported under "other_mode"; run with
--dtype fp32 to make it the headline line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FENC = [[3, 16, 32, 64, 128], [[], [], [], [], []]]  # configure/cfgs.py:11
FDEC = [[128, 64, 32, 32, 16], [[], [], [], [], 3]]  # configure/cfgs.py:12
NZ = 256  # traincfg.yaml:10
METRIC = "train meshes/sec (6890-vert SpiralAE fwd+bwd)"
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that can come out on top, from the
# ncu --set full captures summarised under profiles/ (B=256, bf16).  key = the timer tag bench.py reports as roofline.kernel
NCU_TRAFFIC_BYTES = {
    "slabconv_wgrad[6891>6891x14x32>16]": (169.78e6 + 4.53e6, "profiles/r02z_ncu_l0.csv"),
    "slabconv_fwd[6891>6891x14x32>16]": (116.14e6 + 32.75e6, "profiles/r02z_ncu_l0.csv"),
    "slabconv_dgrad[6891>6891x14x32>16]": (171.88e6 + 77.82e6, "profiles/r02z_ncu_l0.csv"),
    "slabconv_wgrad[863>863x8x128>64]": (84.95e6 + 3.54e6, "profiles/r02z_ncu_l3.csv"),
    "slabconv_fwd[863>863x8x128>64]": (56.72e6 + 8.20e6, "profiles/r02z_ncu_l3.csv"),
    "slab_pool[3446>6891x32]": (86.11e6 + 74.89e6, "profiles/r02z_ncu_pool.csv"),
    "slab_pool_bwd[6891>3446x32]": (209.94e6 + 42.82e6, "profiles/r02z_ncu_pool.csv"),
}
N_INPUT_BATCHES = 8  # distinct resident batches rotated through the timed loop


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (B200_PROFILING.md recipe:
    clocks.sm, clocks.max.sm, clocks_event_reasons.*), every 50 ms on a host thread."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index):
        import threading

        self.samples, self.reasons, self.err = [], set(), None
        self.stop_flag = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err, self.max_mhz = None, repr(e), None
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except AttributeError:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
            self.stop_flag.wait(0.05)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [], "error": self.err}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(self.reasons)}


def base_config(args, world):
    return {"workload": "SpiralAutoencoder training step: fwd + L1 loss + bwd + Adam, synthetic 6890-vertex "
                        "SMPL-topology template (hier_2222), default filters, nz=256 (BASELINE.json configs[2]; "
                        "N=8 is configs[3]'s global batch 2048)",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "params": 28559811, "optimizer": "Adam(lr=1e-3, wd=5e-5)", "parallelism": f"dp{world}",
            "comm_sms": (getattr(args, "comm_sms", 0) if world > 1 else 0),
            "cuda_graph": bool(not getattr(args, "no_graph", False)),
            "host_bound_to_gpu_numa_node": bool(getattr(args, "host_cpus", None)),
            "cache": f"{N_INPUT_BATCHES} distinct input batches rotated; per-step working set (activations + 114 MB "
                     "weights + Adam state) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(steps, warmup, batch, threads=None, budget_s=240.0):
    """The reference's own training step on the host cores: its unmodified ``models.SpiralAutoencoder`` (staged under
    oracle/_ref by oracle/build_ref.py -> kind "reference") driven as train_funcs.py:495-510 drives it (zero_grad, forward,
    l1 loss, backward, Adam of main.py:262), or -- where the staged files are absent -- the oracle port of the same code
    (kind "port").  Nothing of the product package is imported here."""
    import torch.nn.functional as F

    from oracle import build_ref
    from tests.golden.loader import Hierarchy
    from tests.golden.synthetic import fill_deterministic_, synthetic_meshes

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    h = Hierarchy("2222")
    Dd, Ud = h.dense_DU()  # dense padded D/U exactly as the reference multiplies them (main.py:183-205)
    spirals = h.spirals()
    xs = [synthetic_meshes(h.verts0, batch, seed=i) for i in range(2)]
    if build_ref.available():
        kind = "reference"
        ref = build_ref.import_reference_models()
        dev = torch.device("cpu")
        model = ref.SpiralAutoencoder(FENC, FDEC, NZ, h.sizes, h.spiral_sizes, spirals, Dd, Ud, device=dev)
        fill_deterministic_(model, seed=2)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-5)

        def one(x):
            opt.zero_grad()
            xh, _ = model(x)
            loss = F.l1_loss(x, xh)
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        kind = "port"
        from oracle import spiral_oracle as so

        enc, dec = so.conv_plan(FENC, FDEC, 4)

        class Lin(torch.nn.Module):
            def __init__(self, k, n):
                super().__init__()
                self.conv = torch.nn.Linear(k, n)

        class Params(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = torch.nn.ModuleList(Lin(h.spiral_sizes[l] * ci, co) for (l, ci, co, _) in enc)
                self.fc_latent_enc = torch.nn.Linear((h.sizes[-1] + 1) * 128, NZ)
                self.fc_latent_dec = torch.nn.Linear(NZ, (h.sizes[-1] + 1) * 128)
                self.dconv = torch.nn.ModuleList(Lin(h.spiral_sizes[l] * ci, co) for (l, ci, co, _) in dec)

        m = fill_deterministic_(Params(), seed=2)
        params = dict(m.named_parameters())
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=5e-5)

        def one(x):
            opt.zero_grad()
            xh, _ = so.autoencoder_forward(params, x, FENC, FDEC, h.sizes, spirals, Dd, Ud)
            loss = so.l1_loss(x, xh)
            loss.backward()
            opt.step()
            return float(loss.detach())
    times = []
    start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one(xs[i % 2])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
        # bounded sample: a slow host stops after the time budget (at least two timed steps); `steps` reports what ran
        if len(times) >= 2 and time.perf_counter() - start > budget_s:
            break
    total = sum(times)
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times), "cores": threads,
            "batch": batch, "steps": len(times), "kind": kind}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, batch=args.batch)
    world = args.gpus
    what = ("the reference's unmodified models.SpiralAutoencoder (oracle/_ref)" if r["kind"] == "reference"
            else "oracle port of the reference's PyTorch CPU path")
    cfg = base_config(args, world)
    cfg["cuda_graph"] = False
    cfg["global_batch"] = args.batch  # one CPU process: the per-GPU batch of the own arm, N times over at N GPUs
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "meshes/s", "n_gpus": world,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": "meshes/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": f"batch {r['batch']} per step, {r['steps']} steps: {what} (dense D/U bmm, index gather, "
                                       "autograd, Adam) in fp32 on the host cores"},
            "e2e": {"value": r["value"], "unit": "meshes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ own arm
def run_own(args):
    import torch.distributed as dist

    import semantichuman_b200 as shb
    from semantichuman_b200 import functions as fn
    from tests.golden.loader import Hierarchy
    from tests.golden.synthetic import fill_deterministic_, synthetic_meshes
    from semantichuman_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    from semantichuman_b200.dp import bind_host_to_device, init_data_parallel

    # Pin the process to the GPU's NUMA node before any pinned allocation -- except in the process that also times the CPU
    # baseline on "all host cores" (N = 1, rank 0): its worker threads would inherit the narrowed mask.
    runs_cpu_leg = world == 1 and not args.no_cpu_baseline
    args.host_cpus = None if (args.no_numa_bind or runs_cpu_leg) else bind_host_to_device(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_data_parallel(dev, comm_sms=args.comm_sms)
    dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[args.dtype]

    h = Hierarchy("2222")
    Dsp, Usp = h.sparse_DU()
    model = shb.SpiralAutoencoder(FENC, FDEC, latent_size=NZ, sizes=h.sizes, spiral_sizes=h.spiral_sizes,
                                  spirals=h.spirals(dev), D=Dsp, U=Usp, device=dev)
    fill_deterministic_(model, seed=2)
    model = model.to(dev).set_compute_dtype(dtype)
    use_graph = not args.no_graph  # whole step (incl. the gradient all-reduces at N>1) replayed as one CUDA graph per rank
    step = TrainStep(model, graph=use_graph, comm_sms=args.comm_sms if world > 1 else 0)
    B = args.batch
    host = [synthetic_meshes(h.verts0, B, seed=1000 * rank + i).pin_memory() for i in range(N_INPUT_BATCHES)]
    resident = [x.to(dev) for x in host]
    loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if use_graph:
        step.capture(resident[0])  # 3 real steps on a side stream, then the recorded one
    for i in range(args.warmup):
        step(resident[i % N_INPUT_BATCHES])
    barrier()

    # ---- timed region 1: inputs resident in HBM (the headline `value`; nothing but the step inside)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = fn.LAUNCHES["n"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = step(resident[i % N_INPUT_BATCHES])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = fn.LAUNCHES["n"] - launches0
    last_loss = float(loss.item())

    # ---- timed region 1b: the same K steps again with a CUDA-event pair around every libshb200 entry point (on the
    # launching stream): per-kernel durations for the roofline leg.  Its ~0.2 ms/step of event overhead stays out of `value`.
    timer = fn.KernelTimer() if rank == 0 else None
    fn.TIMER = timer
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    i0.record()
    for i in range(args.steps):
        step(resident[i % N_INPUT_BATCHES])
    i1.record()
    barrier()
    fn.TIMER = None
    ms_instr = i0.elapsed_time(i1)
    clocks = sampler.stop() if sampler else None

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    # The host runs E2E_LAG steps ahead of the loss it reads (step i is enqueued, then the loss of step i - E2E_LAG is read on
    # the host, as train_funcs.py:513 reads it every step): every batch is still copied host->device and every loss
    # device->host inside the timed region, but a host thread descheduled for a millisecond no longer idles the GPU.
    E2E_LAG = 2
    for i in range(2):
        step.stage(host[i % N_INPUT_BATCHES])
        step.step_staged(loss_host[i:i + 1])
    barrier()
    done = [torch.cuda.Event() for _ in range(E2E_LAG + 1)]
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    step.stage(host[0])
    seen = 0.0
    for i in range(args.steps):
        k = i % (E2E_LAG + 1)
        step.step_staged(loss_host[k:k + 1])
        done[k].record()
        if i + 1 < args.steps:
            step.stage(host[(i + 1) % N_INPUT_BATCHES])
        if i >= E2E_LAG:
            j = (i - E2E_LAG) % (E2E_LAG + 1)
            done[j].synchronize()
            seen += float(loss_host[j])
    for i in range(max(args.steps - E2E_LAG, 0), args.steps):
        j = i % (E2E_LAG + 1)
        done[j].synchronize()
        seen += float(loss_host[j])
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))

    step.release()  # the graph holds NCCL kernels: drop it before the process group goes away
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    per = timer.summary()
    top_name, top = max(per.items(), key=lambda kv: kv[1]["ms"])
    avg_ms = top["ms"] / top["launches"]
    ai = top["flops"] / max(top["bytes"], 1.0)
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    if ai > ridge and args.dtype == "bf16":
        achieved = top["flops"] / top["launches"] / (avg_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"]}
    else:
        achieved = top["bytes"] / top["launches"] / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"]}
    traffic = NCU_TRAFFIC_BYTES.get(top_name) if args.dtype == "bf16" and B == 256 else None
    roof.update({"instrumented_ms_per_step": ms_instr / args.steps,
                 "traffic": None if traffic is None else traffic[0],
                 "traffic_source": None if traffic is None else traffic[1], "kernel": top_name, "avg_launch_ms": avg_ms,
                 # share of the (clean, graph-replayed) step this kernel accounts for; the instrumented eager pass is
                 # longer than the clean step (event pairs, host gaps), so its own wall time is not the denominator
                 "share_of_step": (top["ms"] / args.steps) / (ms / args.steps),
                 "kernel_family_share": sum(v["ms"] for k, v in per.items() if k.split("[")[0] == top_name.split("[")[0])
                 / args.steps / (ms / args.steps),
                 "peak_source": peaks["source"] + " (sustained figures: kernel timed inside a long step)",
                 "arith_intensity_flop_per_byte": ai})
    kernels = sorted(((k, v["ms"] / args.steps) for k, v in per.items()), key=lambda kv: -kv[1])
    families = {}
    for k, v in kernels:
        families[k.split("[")[0]] = round(families.get(k.split("[")[0], 0.0) + v, 4)
    if args.kernels_out:
        json.dump({k: {"ms_per_step": v, "launches_per_step": per[k]["launches"] / args.steps,
                       "alg_bytes_per_launch": per[k]["bytes"] / max(per[k]["launches"], 1),
                       "flops_per_launch": per[k]["flops"] / max(per[k]["launches"], 1)} for k, v in kernels},
                  open(args.kernels_out, "w"), indent=1)
    step_flops = sum(v["flops"] for v in per.values()) / args.steps
    step_bytes = sum(v["bytes"] for v in per.values()) / args.steps

    # the other precision mode, a few steps, same process (reported beside the headline; not the headline)
    other = None
    if world == 1 and not args.no_other_mode:
        odt = torch.float32 if dtype == torch.bfloat16 else torch.bfloat16
        model.set_compute_dtype(odt)
        for i in range(3):
            step._eager(resident[i % N_INPUT_BATCHES])
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_other = 10
        o0.record()
        for i in range(n_other):
            step._eager(resident[i % N_INPUT_BATCHES])
        o1.record()
        torch.cuda.synchronize()
        oms = o0.elapsed_time(o1) / n_other
        other = {"dtype": "f32" if odt == torch.float32 else "bf16", "ms_per_step": oms, "value": B / (oms * 1e-3),
                 "unit": "meshes/s", "steps": n_other,
                 "note": "fp32 mode: activations and weights as bf16 hi + lo, three tcgen05 products per term (hi.hi + lo.hi + hi.lo), fp32 accumulation; "
                         "1e-4 parity (the reference's own precision)" if odt == torch.float32 else
                         "bf16 mode: bf16 operands, fp32 accumulation and master weights; 2e-2 parity"}
        model.set_compute_dtype(dtype)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(steps=args.cpu_steps, warmup=1, batch=args.cpu_batch)
        cpu = {"value": r["value"], "unit": "meshes/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"batch {r['batch']} per step (a bounded sample of the batch-{B} workload), {r['steps']} timed steps after "
                         f"1 warm-up ({r['ms_per_step']:.0f} ms/step): "
                         + ("the reference's unmodified models.SpiralAutoencoder (oracle/_ref)" if r["kind"] == "reference"
                            else "oracle port of the reference's PyTorch CPU path") + ", fp32, all host cores"}

    total = B * world
    line = {"metric": METRIC, "value": total * args.steps / (ms * 1e-3), "unit": "meshes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.dtype == "fp32" else "bf16",
            "data": "synthetic", "config": base_config(args, world),
            "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "meshes/s",
                    "h2d_bytes_per_step": host[0].numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "other_mode": other, "loss": last_loss, "loss_e2e_mean": seen / args.steps,
            "kernels_ms_per_step": {k: round(v, 4) for k, v in kernels[:12]},
            "kernel_families_ms_per_step": families,
            # whole-step rates of the same algorithmic counts (per GPU; graph-replayed step time, not the instrumented pass)
            "algorithmic_per_step": {"gflop_convs_pools": step_flops / 1e9, "gbytes": step_bytes / 1e9,
                                     "gbytes_per_s": step_bytes / 1e9 / (ms / args.steps * 1e-3),
                                     "tflop_per_s": step_flops / 1e12 / (ms / args.steps * 1e-3),
                                     "frac_of_hbm_peak": step_bytes / 1e9 / (ms / args.steps * 1e-3) / peaks["hbm_gbs"]}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="shb200", choices=["shb200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"],
                    help="bf16 (default): bf16 operands / fp32 accumulate on the tcgen05 kernels, north_star's 2e-2 mode; "
                         "fp32: bf16 hi + lo operands (three products per term) on the same kernels, the 1e-4 mode")
    ap.add_argument("--no-other-mode", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step as a CUDA graph")
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--comm-sms", type=int, default=-1,
                    help="N>1: CTAs NCCL may use = SMs the persistent kernels leave free while a gradient bucket is in flight; "
                         "0 = NCCL defaults, no reservation; -1 (default) = dp.default_comm_sms(world)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the process to the CPUs of the GPU's NUMA node (dp.bind_host_to_device)")
    ap.add_argument("--kernels-out", default=None, help="write the full per-kernel table of the instrumented pass here")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--cpu-batch", type=int, default=64, help="batch of the in-run cpu_baseline sample (10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.comm_sms < 0:
        args.comm_sms = 32 if args.gpus >= 8 else 16  # == semantichuman_b200.dp.default_comm_sms (the reference arm imports no product code)
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_own(args)


if __name__ == "__main__":
    main()
