#!/usr/bin/env python
"""Benchmark of the hot path: ResNet-50 + FPN forward, batch 16 per GPU, synthetic 800x1333 images
zero-padded to 800x1344 (the reference's size_divisor=32 convention), bf16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py
--gpus N ...` (one process per GPU).  Images are independent units (eval-mode BN), so the batch is
sharded across ranks with NO data-path collective ("scaling": "weak": 16 images per GPU);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM;
`e2e` = the same through the public module API from pinned HOST buffers (H2D copy of every step's
batch and a D2H read of the coarsest pyramid level inside the timed region);
`roofline` = the tcgen05 implicit-GEMM kernel family, per-launch device time measured live with CUDA
events (tdet_plan_run_timed) against MEASURED_PEAKS.json; `cpu_baseline` = the oracle (the reference's
algorithm, CPU fp32) timed on this box's host cores on a bounded sample (rank 0, N=1 only).

`--impl reference` times the reference's CPU implementation of the path (the oracle port: the
reference tree itself cannot travel to the GPU box) with all host threads on the same config.
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

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def time_cpu_oracle(depth, h, w, with_fpn, iters, warmup, threads):
    """Reference algorithm (oracle port) on host cores: fp32, eval, no_grad."""
    from oracle import resnet_fpn_oracle as orc
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    bsd = orc.make_resnet_state(depth, generator=g)
    exp = orc.EXPANSION[orc.ARCH[depth][0]]
    in_ch = [64 * 2 ** i * exp for i in range(4)]
    nsd = orc.make_fpn_state(in_ch, 256, 5, generator=g)
    x = make_batch(1, h, w, 0, torch.float32) if with_fpn else \
        torch.randn(1, 3, h, w, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            if with_fpn:
                orc.resnet_fpn_forward(bsd, nsd, x, depth)
            else:
                orc.resnet_forward(bsd, x, depth)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    times = time_cpu_oracle(args.depth, args.height, args.width, True, args.steps, args.warmup, threads)
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ResNet-%d + FPN forward, 800x1333 (padded 800x1344)" % args.depth,
                   "sample": "each step = 1 image on the host CPU (bounded sample of the batch-16 workload)"},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": threads, "kind": "port",
                         "sample": "%d x (1 image 3x800x1344 fp32, oracle port of resnet.py+fpn.py)" % len(times)},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_train(args):
    """Config 4 (secondary bench line): R50+FPN forward+backward with fixed random upstream gradients on
    P2..P6, weights/images resident, per-stage flat fp32 gradient buckets all-reduced on a side stream."""
    import torch.distributed as dist
    from torch_detection_b200 import models, training
    from torch_detection_b200.utils import obj_from_dict
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the all-reduce overlaps persistent one-CTA-per-SM kernels: keep NCCL on a few SMs and leave
        # exactly those free (BucketAllReduce.sm_reserve), or every conv kernel would run a second wave
        os.environ.setdefault("NCCL_MAX_CTAS", str(SM_RESERVE))
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    exp = 4 if args.depth >= 50 else 1
    bb = obj_from_dict(dict(type="ResNet", depth=args.depth, frozen_stages=1, bn_eval=True, bn_frozen=True),
                       parent=models.backbone)
    bb.init_weights()
    neck = obj_from_dict(dict(type="FPN", in_channels=[64 * 2 ** i * exp for i in range(4)], out_channels=256,
                              num_outs=5), parent=models.necks)
    neck.init_weights()
    bb, neck = bb.to(dev).train(), neck.to(dev).train()
    sync = training.BucketAllReduce(defer=True, sm_reserve=SM_RESERVE if world > 1 else 0)
    bb.set_grad_sync(sync)
    neck.set_grad_sync(sync)
    B = args.batch if args.batch != 16 else 8
    x = make_batch(B, args.height, args.width, 100 + rank, torch.bfloat16).to(dev)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    if rank == 0:
        from oracle import resnet_fpn_oracle as orc
        Hp, Wp = x.shape[2], x.shape[3]
        fwd = orc.conv_flops(args.depth, Hp, Wp)[0]
        launches = {"fwd": bb._last_run[0].num_launches + neck._last_run[0].num_launches,
                    "bwd": bb._last_bwd_run[0].num_launches + neck._last_bwd_run[0].num_launches}
        bwd_fl = bb._last_bwd_run[0].flops + neck._last_bwd_run[0].flops
        table = []
        for mod in (neck, bb):
            plan, ext = mod._last_bwd_run
            info = plan.launch_info()
            for inf, t in zip(info, plan.run_timed(ext)):
                inf = dict(inf)
                inf["module"] = type(mod).__name__ + ".backward"
                inf["ms"] = t
                table.append(inf)
        if args.launch_table:
            with open(args.launch_table, "w") as f:
                json.dump(table, f, indent=1)
        value = world * B * args.steps / (ms_total / 1e3)
        wg = [t for t in table if t["kind"] == 5]
        dg = [t for t in table if t["kind"] == 3]
        emit(({
            "metric": "ResNet-%d-FPN train (fwd+bwd) img/s @800x1333 bf16, frozen BN + stem + stage 1" % args.depth,
            "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "config 4: ResNet-%d + FPN forward+backward, batch %d per GPU, %dx%d" %
                                   (args.depth, B, Hp, Wp),
                       "grad_allreduce": "%d flat fp32 buckets, %.1f MB per step, NCCL on a side stream" %
                                         (sync.buckets_reduced // (args.steps + max(args.warmup, 3)),
                                          sync.bytes_reduced / (args.steps + max(args.warmup, 3)) / 1e6)},
            "gflop_per_image": {"forward": fwd / 1e9, "backward_executed": bwd_fl / B / 1e9},
            "tflops_per_gpu": (value / world) * (fwd + bwd_fl / B) / 1e12,
            "launches_per_step": launches, "gpu_launches": (launches["fwd"] + launches["bwd"]) * args.steps,
            "backward_kernels": {
                "wgrad": {"launches": len(wg), "ms": sum(t["ms"] for t in wg),
                          "tflops": sum(t["flops"] for t in wg) / max(sum(t["ms"] for t in wg), 1e-9) / 1e9},
                "dgrad": {"launches": len(dg), "ms": sum(t["ms"] for t in dg),
                          "tflops": sum(t["flops"] for t in dg) / max(sum(t["ms"] for t in dg), 1e-9) / 1e9},
                "other_ms": sum(t["ms"] for t in table if t["kind"] not in (3, 5))},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--width", type=int, default=1333)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch-table", default="", help="write the per-launch timing table (JSON) here")
    ap.add_argument("--io-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="fp32 = the fp32-I/O mode (split-precision kernels, <= 1e-4 vs the fp32 reference); secondary line")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train = BASELINE.json config 4: forward+backward, frozen BN, frozen stem+stage 1, "
                         "batch 8 per GPU, bucketed NCCL gradient all-reduce overlapped with backward")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)

    import torch.distributed as dist
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    from oracle import resnet_fpn_oracle as orc  # FLOP model + cpu_baseline only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- model (reference build API, reference init, seed 0) -------------------------------------
    torch.manual_seed(0)
    exp = 4 if args.depth >= 50 else 1
    in_ch = [64 * 2 ** i * exp for i in range(4)]
    bb = obj_from_dict(dict(type="ResNet", depth=args.depth), parent=models.backbone)
    bb.init_weights()
    neck = obj_from_dict(dict(type="FPN", in_channels=in_ch, out_channels=256, num_outs=5),
                         parent=models.necks)
    neck.init_weights()
    bb = bb.to(dev).eval()
    neck = neck.to(dev).eval()

    B, H, W = args.batch, args.height, args.width
    io_dtype = torch.float32 if args.io_dtype == "fp32" else torch.bfloat16
    x_host = make_batch(B, H, W, 100 + rank, io_dtype).pin_memory()
    x_dev = x_host.to(dev)
    Hp, Wp = x_host.shape[2], x_host.shape[3]

    def step(x):
        with torch.no_grad():
            return neck(bb(x))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        outs = step(x_dev)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- end to end from pinned host buffers (double-buffered H2D on a copy stream) --------------------
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    p6_host = torch.empty((B, 256, outs[-1].shape[2], outs[-1].shape[3]), dtype=io_dtype).contiguous(
        memory_format=torch.channels_last).pin_memory()
    main_stream = torch.cuda.current_stream(dev)

    def e2e_loop(n_steps):
        for b in range(2):
            consumed[b].record(main_stream)
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
            p6_host.copy_(o[-1], non_blocking=True)
        return o

    e2e_loop(3)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    f0.record()
    e2e_loop(args.steps)
    f1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms2 = torch.tensor([max(f0.elapsed_time(f1), 0.0), t_wall * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.max().item())  # device events vs wall clock: take the slower
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    if sampler:
        sampler.stop()

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
        gemm = [l for l in launches if l["kind"] in (1, 3)]
        dom = [l for l in gemm if l["tile_n"] == 256] or gemm
        dom_ms = sum(l["ms"] for l in dom)
        dom_fl = sum(l["flops"] for l in dom)
        all_ms = sum(l["ms"] for l in launches)
        gemm_ms = sum(l["ms"] for l in gemm)
        gemm_fl = sum(l["flops"] for l in gemm)
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12
        traffic, traffic_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r1_s5_traffic.json")
        if os.path.isfile(tpath) and args.depth == 50 and args.batch == 16:
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["roofline_traffic_bytes"]
            traffic_note = ("DRAM read+write of the family's largest launch (FPN P2 3x3, algorithmic 1102.2 MB) from "
                            + tj["source"])
        roof = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
            "kernel": "256-wide tcgen05 implicit-GEMM launches (conv_gemm_kernel<256,...> incl. CTA pairs, and "
                      "conv_swap_kernel: 128x256x16 MMAs; all launches of one step: %d launches, %.1f%% of step time)"
                      % (len(dom), 100.0 * dom_ms / all_ms),
            "peak_source": peaks["_source"] + ", sustained bf16 (kernel timed inside a long step)",
            "frac_of_burst_peak": achieved / float(peaks["bf16_tflops"]),
            "all_gemm_launches": {"tflops": gemm_fl / (gemm_ms * 1e-3) / 1e12, "ms": gemm_ms,
                                  "share_of_step": gemm_ms / all_ms},
            "whole_step_tflops": (value / world) * orc.conv_flops(args.depth, Hp, Wp)[0] / 1e12,
            "launch_ms_sum": all_ms,
        }
        if args.launch_table:
            with open(args.launch_table, "w") as f:
                json.dump(launches, f, indent=1)

    # ---- CPU baseline beside it (rank 0, single-GPU run only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times = time_cpu_oracle(args.depth, H, W, True, 6, 1, threads)
        cpu = {"value": len(times) / sum(times), "unit": "img/s", "cores": threads, "kind": "port",
               "sample": "6 x (1 image 3x%dx%d fp32, ResNet-%d+FPN, oracle port, %d threads); best %.3f s"
                         % (Hp, Wp, args.depth, threads, min(times))}

    if rank == 0:
        n_launch = bb._last_run[0].num_launches + neck._last_run[0].num_launches
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
                       "l2": "per-step working set (~1.4 GB/img of activations) >> 126 MB L2, no explicit flush"},
            "tflops_per_gpu": (value / world) * flops_img / 1e12,
            "e2e": {"value": e2e_value, "unit": "img/s",
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
                    "d2h_bytes_per_step": p6_host.numel() * p6_host.element_size(),
                    "note": "public API neck(backbone(x)); pinned bf16 host batch -> H2D (copy stream, "
                            "double buffered) -> forward -> D2H of P6"},
            "gpu_launches": n_launch * args.steps,
            "launches_per_step": n_launch,
            "clocks": sampler.summary() if sampler else None,
            "roofline": roof, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
