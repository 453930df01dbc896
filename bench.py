#!/usr/bin/env python
"""Benchmark of the hot path: ResNet-50 + FPN forward, batch 16 per GPU, synthetic 800x1333 images
zero-padded to 800x1344 (the reference's size_divisor=32 convention), bf16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py
--gpus N ...` (one process per GPU).  Images are independent units (eval-mode BN), so the batch is
sharded across ranks with NO data-path collective ("scaling": "weak": 16 images per GPU);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

Prints ONE JSON line (rank 0).  Headline keys (BASELINE.json config 2):
  `value`     whole-job images/s with inputs resident in HBM (device events, max over ranks)
  `e2e`       the same through the public module API from pinned HOST buffers: H2D copy of every step's batch and
              a D2H read of the coarsest pyramid level inside the timed region (the consumers of P2..P5 -- the
              detection heads -- live on the device); `e2e_full_copy` also copies ALL five levels back
  `roofline`  the tcgen05 implicit-GEMM kernel family, per-launch device time measured live with CUDA events
              (tdet_plan_run_timed) against the BURST bf16 peak of MEASURED_PEAKS.json
  `sustained` the same step looped for >= 3 s (power-capped clocks) with the median SM clock
  `cpu_baseline` the reference's algorithm (CPU fp32) timed on this box's host cores on a bounded sample
Secondary legs in the SAME line, so that the driver's 1/2/4/8-GPU runs carry them:
  `r101_b64`  BASELINE.json config 3: ResNet-101 + FPN forward, batch 64 sharded over the N GPUs
  `train`     BASELINE.json config 4: ResNet-50 + FPN forward+backward (frozen BN, stem + stage 1 frozen),
              batch 8 per GPU, per-stage gradient buckets all-reduced by NCCL under the backward kernels

`--impl reference` times the reference's CPU implementation of the path with all host threads on the same
config: the UNMODIFIED reference through oracle/reference_shim.py where its tree exists (the dev container),
else the bit-identical oracle port (the reference is pure Python and cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SM_RESERVE = int(os.environ.get("TDET_SM_RESERVE", "8"))  # SMs left to NCCL while gradients are all-reduced
METRIC = "ResNet-50-FPN img/s @800x1333 bf16 at 1/2/4/8 B200; % tensor-pipe peak"
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        p["_source"] = "measured (MEASURED_PEAKS.json)"
        return p
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback (B200_PROFILING.md)"
    return p


def padded_width(w, divisor=32):
    return (w + divisor - 1) // divisor * divisor


def make_batch(batch, h, w, seed, dtype):
    """randn images (post-normalisation statistics) zero-padded right to a multiple of 32."""
    g = torch.Generator().manual_seed(seed)
    wp, hp = padded_width(w), padded_width(h)
    x = torch.zeros(batch, 3, hp, wp, dtype=dtype)
    x[:, :, :h, :w] = torch.randn(batch, 3, h, w, generator=g).to(dtype)
    return x


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def mark(self):
        return len(self.samples)

    def summary(self, first=0, last=None):
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples[first:last]:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the reference's algorithm on the host cores)
# ---------------------------------------------------------------------------------------------------

def cpu_forward_fn(depth, with_fpn):
    """(callable(x) -> outputs, kind): the unmodified reference modules where the tree is importable
    (oracle/reference_shim.py: /root/reference, dev container only), else the bit-identical oracle port."""
    from oracle import reference_shim
    from oracle import resnet_fpn_oracle as orc
    if reference_shim.available():
        try:
            pair = reference_shim.build_pair(depth, seed=0)
            if pair is not None:
                bb, neck = pair
                if with_fpn:
                    return (lambda x: neck(bb(x))), "reference"
                return (lambda x: bb(x)), "reference"
        except Exception as e:  # a broken tree must not take the bench down: fall back to the port
            sys.stderr.write("reference shim failed (%s): timing the oracle port\n" % (e,))
    g = torch.Generator().manual_seed(0)
    bsd = orc.make_resnet_state(depth, generator=g)
    exp = orc.EXPANSION[orc.ARCH[depth][0]]
    nsd = orc.make_fpn_state([64 * 2 ** i * exp for i in range(4)], 256, 5, generator=g)
    if with_fpn:
        return (lambda x: orc.resnet_fpn_forward(bsd, nsd, x, depth)), "port"
    return (lambda x: orc.resnet_forward(bsd, x, depth)), "port"


def time_cpu(depth, h, w, with_fpn, iters, warmup, threads):
    """Per-iteration seconds of the CPU forward: fp32, eval, no_grad.  with_fpn: the
    R-depth + FPN path on the image padded to a multiple of 32; else BASELINE.json config 1 exactly (backbone
    only, raw h x w)."""
    torch.set_num_threads(threads)
    fn, kind = cpu_forward_fn(depth, with_fpn)
    g = torch.Generator().manual_seed(0)
    x = make_batch(1, h, w, 0, torch.float32) if with_fpn else torch.randn(1, 3, h, w, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            fn(x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, kind


def median(v):
    s = sorted(v)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    times, kind = time_cpu(args.depth, args.height, args.width, True, max(args.steps, 1), max(args.warmup, 1), threads)
    total = sum(times)
    value = len(times) / total
    med = 1.0 / median(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-%d + FPN forward, 800x1333 (padded 800x1344)" % args.depth,
                   "sample": "each step = 1 image on the host CPU (bounded sample of the batch-16 workload)"},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": threads, "kind": kind,
                         "median_img_s": med, "best_img_s": 1.0 / min(times),
                         "sample": "%d x (1 image 3x800x1344 fp32, %s, %d threads)" % (
                             len(times), "unmodified reference modules" if kind == "reference"
                             else "oracle port of resnet.py+fpn.py", threads)},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------

def build_pair(depth, dev, train=False):
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    torch.manual_seed(0)
    exp = 4 if depth >= 50 else 1
    kw = dict(frozen_stages=1, bn_eval=True, bn_frozen=True) if train else {}
    bb = obj_from_dict(dict(type="ResNet", depth=depth, **kw), parent=models.backbone)
    bb.init_weights()
    neck = obj_from_dict(dict(type="FPN", in_channels=[64 * 2 ** i * exp for i in range(4)], out_channels=256,
                              num_outs=5), parent=models.necks)
    neck.init_weights()
    bb, neck = bb.to(dev), neck.to(dev)
    if train:
        return bb.train(), neck.train()
    return bb.eval(), neck.eval()


class Ctx(object):
    """Rank / device / collective plumbing of one bench process."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # the training leg's all-reduce overlaps persistent one-CTA-per-SM kernels: NCCL stays on a few SMs
            # and exactly those are left free (BucketAllReduce.sm_reserve), or every conv would run a second wave
            os.environ.setdefault("NCCL_MAX_CTAS", str(SM_RESERVE))
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = torch.tensor(list(values), device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class quiet_host(object):
    """Python's cyclic garbage collector paused for a timed region (collected just before): a generation-2 pass over
    the modules / plans of earlier legs takes 20-200 ms of HOST time and, landing inside a 10-step loop, showed up as
    9.3 ms training steps averaging 11-33 ms at random (per-step events: median 9.3, one outlier)."""

    def __enter__(self):
        import gc
        gc.collect()
        self.was_enabled = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self.was_enabled:
            gc.enable()


def timed_steps(ctx, fn, steps, warmup):
    """W untimed steps, then exactly `steps` steps between barrier + synchronize, device events, max over ranks."""
    for _ in range(max(warmup, 3)):
        fn()
    with quiet_host():
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        ctx.barrier()
    return ctx.max_over_ranks([e0.elapsed_time(e1)])[0], out


def sustained_leg(ctx, fn, images_per_step, min_seconds, sampler):
    """Loops the step until >= min_seconds of device time have passed (the power cap then sets the clock)."""
    fn()
    with quiet_host():
        ctx.barrier()
        first = sampler.mark() if sampler else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        steps = 0
        while True:
            for _ in range(20):
                fn()
            steps += 20
            torch.cuda.synchronize()
            if time.perf_counter() - t0 >= min_seconds:
                break
        e1.record()
        ctx.barrier()
    ms = ctx.max_over_ranks([e0.elapsed_time(e1)])[0]
    clocks = sampler.summary(first, sampler.mark()) if sampler else None
    return {"value": ctx.world * images_per_step * steps / (ms / 1e3), "unit": "img/s", "seconds": ms / 1e3,
            "steps": steps, "ms_per_step": ms / steps, "clocks": clocks}


def e2e_leg(ctx, step, x_host, x_dev, out_like, steps, copy_levels):
    """End to end from pinned host memory: double-buffered H2D on a copy stream, forward, D2H of the listed
    pyramid levels.  Returns (img/s, h2d bytes, d2h bytes)."""
    dev = ctx.dev
    B = x_host.shape[0]
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    host_sets = [[torch.empty(out_like[i].shape, dtype=out_like[i].dtype).contiguous(
        memory_format=torch.channels_last).pin_memory() for i in copy_levels] for _ in range(2 if len(copy_levels) > 1 else 1)]
    host_out = host_sets[0]
    main_stream = torch.cuda.current_stream(dev)
    # more than one level back (the PCIe-bound full copy): the D2H runs on a stream of its own from device-side
    # snapshots of the pyramid (two sets, 0.25 ms of HBM copies per step), so it overlaps the NEXT steps' compute and
    # H2D instead of serialising with them; the host buffers of a step are complete when its `landed` event fires
    overlap_d2h = len(copy_levels) > 1
    d2h_stream = torch.cuda.Stream(device=dev) if overlap_d2h else None
    snap = [[torch.empty_like(out_like[i]) for i in copy_levels] for _ in range(2)] if overlap_d2h else None
    snapped = [torch.cuda.Event(), torch.cuda.Event()]
    landed = [torch.cuda.Event(), torch.cuda.Event()]

    def loop(n_steps):
        for b in range(2):
            consumed[b].record(main_stream)
            if overlap_d2h:
                landed[b].record(d2h_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[0])
            stage[0].copy_(x_host, non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n_steps):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < n_steps:
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[nxt])
                    stage[nxt].copy_(x_host, non_blocking=True)
                    ready[nxt].record(copy_stream)
            main_stream.wait_event(ready[cur])
            o = step(stage[cur])
            consumed[cur].record(main_stream)
            if overlap_d2h:
                main_stream.wait_event(landed[cur])      # snapshot set `cur` has left for the host (step i - 2)
                for d, lvl in zip(snap[cur], copy_levels):
                    d.copy_(o[lvl], non_blocking=True)
                snapped[cur].record(main_stream)
                with torch.cuda.stream(d2h_stream):
                    d2h_stream.wait_event(snapped[cur])
                    for h, d in zip(host_sets[cur], snap[cur]):
                        h.copy_(d, non_blocking=True)
                    landed[cur].record(d2h_stream)
            else:
                for h, lvl in zip(host_out, copy_levels):
                    h.copy_(o[lvl], non_blocking=True)
        if overlap_d2h:
            for b in range(2):
                main_stream.wait_event(landed[b])        # the timed region ends when the last results are on the host

    loop(3)
    with quiet_host():
        ctx.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        f0.record()
        loop(steps)
        f1.record()
        ctx.barrier()
        t_wall = time.perf_counter() - t_wall0
    ms = max(ctx.max_over_ranks([max(f0.elapsed_time(f1), 0.0), t_wall * 1e3]))  # events vs wall clock: the slower
    return (ctx.world * B * steps / (ms / 1e3), x_host.numel() * x_host.element_size(),
            sum(h.numel() * h.element_size() for h in host_out))


def train_leg(ctx, args, steps, warmup, launch_table=""):
    """BASELINE.json config 4: R50+FPN forward+backward with fixed random upstream gradients on P2..P6,
    weights/images resident, per-stage flat fp32 gradient buckets all-reduced on a side stream."""
    from torch_detection_b200 import training
    from oracle import resnet_fpn_oracle as orc  # FLOP model only
    dev, world = ctx.dev, ctx.world
    bb, neck = build_pair(args.depth, dev, train=True)
    sync = training.BucketAllReduce(defer=True, sm_reserve=SM_RESERVE if world > 1 else 0)
    bb.set_grad_sync(sync)
    neck.set_grad_sync(sync)
    B = args.train_batch
    x = make_batch(B, args.height, args.width, 100 + ctx.rank, torch.bfloat16).to(dev)
    params = [p for p in list(bb.parameters()) + list(neck.parameters()) if p.requires_grad]
    outs = neck(bb(x))
    g = torch.Generator().manual_seed(7)
    grads = [(torch.randn(o.shape, generator=g) * 1e-3).to(torch.bfloat16).to(dev).contiguous(
        memory_format=torch.channels_last) for o in outs]
    del outs

    def step():
        for p in params:
            p.grad = None
        o = neck(bb(x))
        torch.autograd.backward(list(o), grads)
        sync.finish()

    ms_total, _ = timed_steps(ctx, step, steps, warmup)
    # per-step device times (events between steps): a hiccup shows up as max >> median
    n_probe = max(steps, 30)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_probe + 1)]
    host_t = [0.0] * (n_probe + 1)
    with quiet_host():
        torch.cuda.synchronize()
        evs[0].record()
        host_t[0] = time.perf_counter()
        segs = [0] * (n_probe + 1)
        segs[0] = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
        for i in range(n_probe):
            step()
            evs[i + 1].record()
            host_t[i + 1] = time.perf_counter()
            segs[i + 1] = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
        torch.cuda.synchronize()
    dev_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_probe)]
    host_step = [(host_t[i + 1] - host_t[i]) * 1e3 for i in range(n_probe)]
    per_step = sorted(dev_step)
    worst = max(range(n_probe), key=lambda i: dev_step[i])
    # host time to ENQUEUE one step (no synchronisation inside): if it approaches ms_per_step the leg is launch-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    enqueue_ms = (time.perf_counter() - t0) / 3 * 1e3
    torch.cuda.synchronize()
    b0, n0 = sync.bytes_reduced, sync.buckets_reduced
    step()
    allreduce_mb = (sync.bytes_reduced - b0) / 1e6
    buckets = sync.buckets_reduced - n0
    # cost of the collective: the same steps with the all-reduce switched off (buckets still copied out), and
    # the device time of the collectives themselves on the side stream
    exposed_ms = coll_ms = None
    if world > 1:
        sync.profile(True)
        for _ in range(3):
            step()
        coll_ms = sync.collective_ms() / 3.0
        sync.profile(False)
        sync.enabled = False
        ms_off, _ = timed_steps(ctx, step, steps, 2)
        sync.enabled = True
        exposed_ms = ms_total / steps - ms_off / steps
        coll_ms = ctx.max_over_ranks([coll_ms])[0]
    res = None
    if ctx.rank == 0:
        Hp, Wp = x.shape[2], x.shape[3]
        fwd = orc.conv_flops(args.depth, Hp, Wp)[0]
        launches = {"fwd": bb._last_run[0].num_launches + neck._last_run[0].num_launches,
                    "bwd": bb._last_bwd_run[0].num_launches + neck._last_bwd_run[0].num_launches}
        bwd_fl = bb._last_bwd_run[0].flops + neck._last_bwd_run[0].flops
        value = world * B * steps / (ms_total / 1e3)
        res = {
            "metric": "ResNet-%d-FPN train (fwd+bwd) img/s @800x1333 bf16, frozen BN + stem + stage 1" % args.depth,
            "img_s": value, "value": value, "unit": "img/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms_total / steps, "batch_per_gpu": B, "host_enqueue_ms_per_step": enqueue_ms,
            "per_step_ms": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1],
                            "n": n_probe, "worst_step": worst, "host_ms_around_worst": host_step[max(worst - 2, 0):worst + 2],
                            "host_ms_max": max(host_step),
                            "new_allocator_segments_per_step": [segs[i + 1] - segs[i] for i in range(n_probe)],
                            "slow_host_step": max(range(n_probe), key=lambda i: host_step[i])},
            "allreduce_mb": allreduce_mb if world > 1 else 0.0, "buckets_per_step": buckets,
            "allreduce_device_ms": coll_ms, "allreduce_exposed_ms": exposed_ms,
            "overlap_ms": (coll_ms - max(exposed_ms, 0.0)) if coll_ms is not None else None,
            "sm_reserve": SM_RESERVE if world > 1 else 0,
            "gflop_per_image": {"forward": fwd / 1e9, "backward_executed": bwd_fl / B / 1e9},
            "tflops_per_gpu": (value / world) * (fwd + bwd_fl / B) / 1e12,
            "launches_per_step": launches,
            "gpu_launches": (launches["fwd"] + launches["bwd"]) * steps,
        }
        if launch_table:
            table = []
            for mod in (neck, bb):
                plan, ext = mod._last_bwd_run
                for inf, t in zip(plan.launch_info(), plan.run_timed(ext)):
                    inf = dict(inf)
                    inf["module"] = type(mod).__name__ + ".backward"
                    inf["ms"] = t
                    table.append(inf)
            with open(launch_table, "w") as f:
                json.dump(table, f, indent=1)
            wg = [t for t in table if t["kind"] == 5]
            dg = [t for t in table if t["kind"] == 3]
            res["backward_kernels"] = {
                "wgrad": {"launches": len(wg), "ms": sum(t["ms"] for t in wg),
                          "tflops": sum(t["flops"] for t in wg) / max(sum(t["ms"] for t in wg), 1e-9) / 1e9},
                "dgrad": {"launches": len(dg), "ms": sum(t["ms"] for t in dg),
                          "tflops": sum(t["flops"] for t in dg) / max(sum(t["ms"] for t in dg), 1e-9) / 1e9},
                "other_ms": sum(t["ms"] for t in table if t["kind"] not in (3, 5))}
    del bb, neck, sync, x, grads, params
    torch.cuda.empty_cache()
    from torch_detection_b200 import engine
    if world > 1:
        engine.set_sm_reserve(dev, 0)
    return res


def r101_leg(ctx, args, steps, warmup):
    """BASELINE.json config 3: ResNet-101 + FPN forward, batch 64 sharded over the N GPUs (64 / N per GPU)."""
    from oracle import resnet_fpn_oracle as orc  # FLOP model only
    per_gpu = max(64 // ctx.world, 1)
    bb, neck = build_pair(101, ctx.dev)
    x = make_batch(per_gpu, args.height, args.width, 300 + ctx.rank, torch.bfloat16).to(ctx.dev)

    def step():
        with torch.no_grad():
            return neck(bb(x))

    ms, _ = timed_steps(ctx, step, steps, warmup)
    value = ctx.world * per_gpu * steps / (ms / 1e3)
    flops = orc.conv_flops(101, x.shape[2], x.shape[3])[0]
    res = {"img_s": value, "unit": "img/s", "global_batch": per_gpu * ctx.world, "batch_per_gpu": per_gpu,
           "ms_per_step": ms / steps, "steps": steps, "tflops_per_gpu": (value / ctx.world) * flops / 1e12,
           "workload": "config 3: ResNet-101 + FPN forward, batch 64 sharded over %d GPU(s), %dx%d" % (
               ctx.world, x.shape[2], x.shape[3])}
    del bb, neck, x
    if os.environ.get("TDET_BENCH_KEEP_CACHE", "0") == "0":
        torch.cuda.empty_cache()
    return res


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: route everything else that writes to fd 1 (NCCL's version
    banner, library chatter) to stderr and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--train-batch", type=int, default=8, help="images per GPU of the training leg (config 4)")
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true",
                    help="skip the sustained / full-copy / R101 batch-64 / training legs (quick A/B runs)")
    ap.add_argument("--legs", default="full,sustained,r101,train",
                    help="which of the extra legs run (comma list of full, sustained, r101, train)")
    ap.add_argument("--launch-table", default="", help="write the per-launch timing table (JSON) here")
    ap.add_argument("--io-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="fp32 = the fp32-I/O mode (split-precision kernels, <= 1e-4 vs the fp32 reference); secondary line")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train = BASELINE.json config 4 alone: forward+backward, frozen BN, frozen stem+stage 1, "
                         "batch 8 per GPU, bucketed NCCL gradient all-reduce overlapped with backward")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from oracle import resnet_fpn_oracle as orc  # FLOP model + cpu_baseline only
    ctx = Ctx(args)
    rank, world, dev = ctx.rank, ctx.world, ctx.dev

    if args.mode == "train":
        if args.batch != 16:
            args.train_batch = args.batch
        res = train_leg(ctx, args, args.steps, args.warmup, args.launch_table)
        if rank == 0:
            res.update({"warmup": max(args.warmup, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                        "dtype": "bf16", "data": "synthetic",
                        "config": {"workload": "config 4: ResNet-%d + FPN forward+backward, batch %d per GPU, 800x1344"
                                               % (args.depth, args.train_batch)}})
            emit(res)
        ctx.close()
        return 0

    # ---- model (reference build API, reference init, seed 0) -------------------------------------
    bb, neck = build_pair(args.depth, dev)
    B, H, W = args.batch, args.height, args.width
    io_dtype = torch.float32 if args.io_dtype == "fp32" else torch.bfloat16
    x_host = make_batch(B, H, W, 100 + rank, io_dtype).pin_memory()
    x_dev = x_host.to(dev)
    Hp, Wp = x_host.shape[2], x_host.shape[3]

    def step(x):
        with torch.no_grad():
            return neck(bb(x))

    # ---- device-resident throughput (the headline `value`) ---------------------------------------------
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    ctx.barrier()
    c_first = sampler.mark() if sampler else 0
    ms_total, outs = timed_steps(ctx, lambda: step(x_dev), args.steps, 0)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end from pinned host buffers -----------------------------------------------------------
    e2e_value, h2d, d2h = e2e_leg(ctx, step, x_host, x_dev, outs, args.steps, [len(outs) - 1])
    c_last = sampler.mark() if sampler else 0
    full = None
    sustained = None
    legs = set() if args.no_extra_legs else set(v.strip() for v in args.legs.split(",") if v.strip())
    if "full" in legs:
        fv, fh, fd = e2e_leg(ctx, step, x_host, x_dev, outs, max(args.steps // 2, 5), list(range(len(outs))))
        full = {"value": fv, "unit": "img/s", "h2d_bytes_per_step": fh, "d2h_bytes_per_step": fd,
                "note": "as e2e, but ALL of P2..P6 are copied back to pinned host memory every step (PCIe-bound): D2H on its own stream from device-side snapshots, overlapped with the following steps"}
    if "sustained" in legs:
        sustained = sustained_leg(ctx, lambda: step(x_dev), B, 3.0, sampler)

    # ---- roofline of the tcgen05 GEMM kernel family, measured live with CUDA events --------------------
    roof = None
    launches = []
    if rank == 0:
        peaks = load_peaks()
        step(x_dev)
        for mod in (bb, neck):
            plan, ext = mod._last_run
            info = plan.launch_info()
            acc = [0.0] * len(info)
            reps = 3
            for _ in range(reps):
                for i, t in enumerate(plan.run_timed(ext)):
                    acc[i] += t / reps
            for i, (inf, t) in enumerate(zip(info, acc)):
                inf = dict(inf)
                inf["module"] = type(mod).__name__
                inf["ms"] = t
                launches.append(inf)
        gemm = [l for l in launches if l["kind"] in (1, 3, 18)]  # conv, stem, fused bottleneck tail
        dom = [l for l in gemm if l["tile_n"] == 256 and l["kind"] in (1, 3)] or gemm
        dom_ms = sum(l["ms"] for l in dom)
        dom_fl = sum(l["flops"] for l in dom)
        all_ms = sum(l["ms"] for l in launches)
        gemm_ms = sum(l["ms"] for l in gemm)
        gemm_fl = sum(l["flops"] for l in gemm)
        peak = float(peaks["bf16_tflops"])  # burst: the denominator BASELINE.md section 3 names
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12
        traffic, traffic_note = None, None
        for tname in ("r2_traffic.json", "r1_s5_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.isfile(tpath) and args.depth == 50 and args.batch == 16:
                with open(tpath) as f:
                    tj = json.load(f)
                traffic = tj["roofline_traffic_bytes"]
                traffic_note = ("DRAM read+write of the family's largest launch (FPN P2 3x3, algorithmic 1102.2 MB) from "
                                + tj["source"])
                break
        whole = (value / world) * orc.conv_flops(args.depth, Hp, Wp)[0] / 1e12
        roof = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
            "kernel": "256-wide tcgen05 implicit-GEMM launches (conv_gemm_kernel<256,...> incl. CTA pairs and dual-source, "
                      "and conv_swap_kernel: 128x256x16 MMAs; all launches of one step: %d launches, %.1f%% of step time)"
                      % (len(dom), 100.0 * dom_ms / all_ms),
            "peak_source": peaks["_source"] + ", burst bf16 (cuBLAS 8192^3, best of 10)",
            "frac_of_sustained_peak": achieved / float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])),
            "all_gemm_launches": {"tflops": gemm_fl / (gemm_ms * 1e-3) / 1e12, "ms": gemm_ms,
                                  "share_of_step": gemm_ms / all_ms},
            "whole_step_tflops": whole, "whole_step_frac": whole / peak,
            "launch_ms_sum": all_ms,
        }
        if args.launch_table:
            with open(args.launch_table, "w") as f:
                json.dump(launches, f, indent=1)
    n_launch = bb._last_run[0].num_launches + neck._last_run[0].num_launches
    del bb, neck, outs, x_dev
    torch.cuda.empty_cache()

    # ---- secondary legs: config 3 (R101, batch 64 sharded) and config 4 (training, NCCL all-reduce) ------
    r101 = train = None
    import gc
    gc.collect()                 # the modules hold reference cycles: free their arenas before the next leg builds its own
    torch.cuda.empty_cache()
    if "train" in legs and args.io_dtype == "bf16":
        m0 = sampler.mark() if sampler else 0
        train = train_leg(ctx, args, 10, 3)
        if train is not None and sampler:
            train["clocks"] = sampler.summary(m0, sampler.mark())
        gc.collect()
        torch.cuda.empty_cache()
    if "r101" in legs and args.io_dtype == "bf16":
        r101 = r101_leg(ctx, args, 5, 3)
    if sampler:
        sampler.stop()

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t_fpn, kind = time_cpu(args.depth, H, W, True, 10, 2, threads)
        t_bb, _ = time_cpu(args.depth, H, W, False, 10, 2, threads)
        t_one, _ = time_cpu(args.depth, H, W, False, 2, 1, 1)
        gf_fpn = orc.conv_flops(args.depth, Hp, Wp)[0] / 1e9
        cpu = {"value": 1.0 / median(t_fpn), "unit": "img/s", "cores": threads, "kind": kind,
               "best_img_s": 1.0 / min(t_fpn), "gflops": gf_fpn / median(t_fpn),
               "sample": "median of 10 x (1 image 3x%dx%d fp32, ResNet-%d+FPN, %s, %d threads) after 2 warm-up"
                         % (Hp, Wp, args.depth, "unmodified reference" if kind == "reference" else "oracle port", threads),
               "config1_backbone_only": {
                   "workload": "BASELINE.json config 1: ResNet-%d backbone, 1x3x%dx%d fp32" % (args.depth, H, W),
                   "median_img_s": 1.0 / median(t_bb), "best_img_s": 1.0 / min(t_bb), "cores": threads, "iters": len(t_bb),
                   "one_thread_img_s": 1.0 / median(t_one), "one_thread_iters": len(t_one)}}

    if rank == 0:
        flops_img = orc.conv_flops(args.depth, Hp, Wp)[0]
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.io_dtype == "bf16" else "bf16 hi+lo pairs (split precision, fp32 I/O)",
            "data": "synthetic",
            "config": {"workload": "ResNet-%d + FPN forward, batch %d per GPU, %dx%d zero-padded to %dx%d, "
                                   "%s NCHW in, P2..P6 %s out" % (args.depth, B, H, W, Hp, Wp, args.io_dtype,
                                                                   args.io_dtype),
                       "images_per_gpu": B, "global_batch": B * world, "gflop_per_image": flops_img / 1e9,
                       "parallelism": "batch sharded, no data-path collective",
                       "l2": "per-step working set (~1.4 GB/img of activations) >> 126 MB L2, no explicit flush",
                       "host": "python's cyclic gc collected before and paused during every timed region"},
            "tflops_per_gpu": (value / world) * flops_img / 1e12,
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "public API neck(backbone(x)); pinned host batch -> H2D (copy stream, double buffered) -> "
                            "forward -> D2H of P6; P2..P5 stay on the device, where their consumers (the detection "
                            "heads) run -- e2e_full_copy moves all five levels"},
            "e2e_full_copy": full,
            "sustained": sustained,
            "gpu_launches": n_launch * args.steps,
            "launches_per_step": n_launch,
            "clocks": sampler.summary(c_first, c_last) if sampler else None,
            "roofline": roof, "cpu_baseline": cpu,
            "r101_b64": r101, "train": train,
        }
        emit(line)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
