#!/usr/bin/env python
"""bench.py -- KITTI-shape scenes/sec of the full 3DSSD SA backbone (BASELINE.json metric, configs[1]).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (the reference arm, see below)

A "step" is one pass of the whole backbone (layer1..layer4 of configs/kitti/3dssd/3dssd.yaml, incl. the vote
layer) over one batch of 8 synthetic KITTI-shaped scenes [8,16384,4] per GPU (weak scaling: the batch is
sharded by scene, no data-path collective; one NCCL all-gather of the per-scene detection blocks ends a step).

value  : scenes/s, inputs resident in HBM; every step is one CUDA-graph replay, `--pipeline` (default 8) steps are
         in flight on separate streams (a step chains latency-bound FPS stages and throughput-bound MLP stages, so
         the FPS of step i+1 overlaps the MLP of step i); timed with ONE CUDA-event pair around all K steps, L2
         flushed before every step, max over ranks.  config.latency_ms_single_step is the un-overlapped step time.
e2e    : same metric through the public API with HOST buffers: pinned H2D copy of the batch + backbone +
         D2H read of the detection block inside the timed region.
roofline: dominant kernel (D-FPS layer 1) timed live with CUDA events on its launch stream.
cpu_baseline: the CPU restatement (oracle/, fp32 BLAS MLP) on a bounded sample of the same workload.

--impl reference: the reference's implementation of this path is CUDA (lib/utils/tf_ops/*_g.cu) -- it has no
CPU code for sampling/grouping (SURVEY.md finding 4).  The arm therefore runs the reference's OWN kernels,
compiled unmodified into oracle/_ref/libref_ops.so, on the same GPU, with PyTorch fp32 ops standing in
one-for-one for the TF stock ops (oracle/ref_layers.py); if that library is absent it falls back to the CPU
port.  Its cpu_baseline object carries the CPU port timing in both cases.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "KITTI-shape scenes/sec (B×16384 pts) full SA backbone, 1/2/4/8×B200"
UNIT = "scenes/s"
SCENES_PER_GPU = 8
NPOINTS = 16384


# ----------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on a bounded sample
# ----------------------------------------------------------------------------------------------------------
def cpu_baseline(arch, params, pts_np, max_scenes=2):
    """Times the CPU restatement of the whole backbone (oracle/, all host threads; MLP through numpy's fp32
    BLAS instead of the double-precision checker loop) on the first `max_scenes` scenes of the workload."""
    from oracle import layers as olayers
    from oracle import ops as oops
    oops.build()
    cores = oops.get_threads()
    saved = oops.linear_bn_relu

    def fast_linear(x, w, bias=None, bn=None, relu=True):
        y = x.reshape(-1, x.shape[-1]) @ w
        if bias is not None:
            y = y + bias
        if bn is not None:
            g, be, mu, var = bn
            inv = g / np.sqrt(var + np.float32(1e-3))
            y = y * inv + (be - mu * inv)
        if relu:
            np.maximum(y, 0, out=y)
        return y.reshape(x.shape[:-1] + (w.shape[1],)).astype(np.float32, copy=False)

    oops.linear_bn_relu = fast_linear
    try:
        sample = np.ascontiguousarray(pts_np[:max_scenes])
        t0 = time.perf_counter()
        olayers.backbone_forward(arch, sample, params, ffps_mode="matrix")
        dt = time.perf_counter() - t0
    finally:
        oops.linear_bn_relu = saved
    return {"value": sample.shape[0] / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the %d scenes of one step, full backbone, oracle/ C restatement + numpy fp32 BLAS MLP, %.1f s"
                      % (sample.shape[0], pts_np.shape[0], dt)}


def ncu_dram_bytes(name):
    """dram read+write bytes per launch from a committed ncu export (profiles/<name>), or None."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, seen = 0.0, 0
    for line in open(path):
        parts = line.split()
        if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[2] in unit:
            tot += float(parts[1]) * unit[parts[2]]
            seen += 1
    return tot if seen == 2 else None


def tensor_roofline(torch, pkg, dev, peaks):
    """Largest GEMM of the step (layer 4, scale 2, last conv: 65536 x 512 -> 1024, max-pooled over 32 neighbours) timed
    alone: achieved = bf16 MMA flops actually issued (3 per logical MMA, hi/lo split) / time, against the measured
    cuBLAS bf16 peak."""
    P = pkg.params
    rng = np.random.default_rng(1)
    prm = {}
    P._conv_init(rng, prm, "s", 512, 1024, True)
    f = P.fold(prm, "s", True, dev)
    x = torch.randn((2048, 32, 512), device=dev)
    hi, lo = pkg.split_rows(x)
    for _ in range(3):
        pkg.linear_tc(hi, lo, f, pool=32)
    torch.cuda.synchronize()
    ev = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pkg.linear_tc(hi, lo, f, pool=32); e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    peak = float(peaks.get("bf16_tflops", peaks.get("bf16_tfs", 1700.0))) if peaks else 1700.0
    ach = 3 * 2.0 * 65536 * 512 * 1024 / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu_dram_bytes("r01_ncu_tc_l4_pooled.txt"),
            "kernel": "linear_tc_kernel (layer 4 scale 2 last conv, 65536x512x1024, pooled)", "kernel_ms": ms,
            "note": "bf16 MMA flops issued = 3 per logical fp32-grade MMA (hi.hi + lo.hi + hi.lo); fp32-equivalent rate = achieved / 3"}


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ffps-mode", default="direct", choices=["direct", "matrix", "fused"])
    ap.add_argument("--mlp-mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--gather-in-kernel", type=int, default=1, help="1: first conv of unfused SA scales gathers its operand itself")
    ap.add_argument("--hoist-first", type=int, default=2, help="first conv of an SA scale evaluated per point (hoisted), not per grouped row: 0 off, 1 layer-by-layer scales, 2 all")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (for ncu captures)")
    ap.add_argument("--pipeline", type=int, default=8, help="steps in flight (independent CUDA graphs on separate streams)")
    ap.add_argument("--fps-cluster", type=int, default=0, help="tuning: force the FPS cluster size (0 = heuristic)")
    ap.add_argument("--fps-cluster-cap", type=int, default=-1,
                    help="cap on the heuristic FPS cluster size; default 4 when steps are pipelined (frees SMs), else none")
    ap.add_argument("--fps-variant", type=int, default=0, help="tuning: 0 = auto, 1 = force the xyz-in-packet D-FPS kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-scenes", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return                                   # under torchrun rank 0 alone runs the reference arm

    import torch
    pkg = importlib.import_module("3dssd_b200")
    synth = importlib.import_module("3dssd_b200.synth")
    if args.impl == "reference":
        rank, world, local = 0, int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    else:
        rank, world, local = pkg.dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    arch = pkg.config.ARCH_3DSSD
    params = pkg.params.init_params(arch, 1, seed=0)
    head_params = pkg.params.init_head_params(pkg.config.layer_channels(arch, 1)[-1], seed=1)
    pts_np = synth.kitti_like(SCENES_PER_GPU, NPOINTS, seed=1000 + rank * SCENES_PER_GPU)

    if args.impl == "reference":
        return run_reference(args, torch, pkg, arch, params, pts_np, rank, world, dev)

    import torch.distributed as dist
    if args.fps_cluster_cap < 0:
        args.fps_cluster_cap = 4 if (args.pipeline > 1 and not args.no_graph) else 0
    pkg.lib().ssd3d_tune_set_fps_cluster_cap(args.fps_cluster_cap)
    pkg.lib().ssd3d_tune_set_fps_cluster(args.fps_cluster)
    pkg.lib().ssd3d_tune_set_fps_variant(args.fps_variant)
    # detection head + decode + GPU BEV-NMS produce the per-scene detection block that is gathered / copied to the host
    head = pkg.DetectionHead(params=head_params, device=dev)
    net = pkg.SABackbone(arch, params, in_channels=1, device=dev, ffps_mode=args.ffps_mode, mlp_mode=args.mlp_mode, head=head,
                         gather_in_kernel=bool(args.gather_in_kernel), hoist_first=args.hoist_first)
    pts = torch.from_numpy(pts_np).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    # launches per step, counted at the ctypes boundary (every C-ABI call below launches exactly one kernel)
    counter = {"n": 0}
    L = pkg.lib()
    counted = [n for n in pkg.EXPORTS if n not in ("ssd3d_version", "ssd3d_last_error", "ssd3d_fps_needs_temp",
                                                    "ssd3d_tune_set_fps_cluster")]
    originals = {n: getattr(L, n) for n in counted}

    class Counting:
        def __init__(self, fn):
            self.fn = fn

        def __call__(self, *a):
            counter["n"] += 1
            return self.fn(*a)

    for n in counted:
        setattr(L, n, Counting(originals[n]))
    out = net.forward(pts)
    blk, cnt = net.detections(out[0], out[1])
    torch.cuda.synchronize()
    launches_per_step = counter["n"]
    for n in counted:
        setattr(L, n, originals[n])

    # ---- steps in flight: P independent step pipelines (own CUDA graph + static buffers + stream each).  A step is a
    # chain of latency-bound stages (FPS) and throughput-bound stages (MLP); with two steps in flight the FPS of step
    # i+1 runs on SMs the MLP of step i leaves idle.  P=1 gives plain back-to-back steps.
    P = 1 if args.no_graph else max(1, args.pipeline)
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    runners = []
    for _ in range(P):
        if args.no_graph:
            static_in = pts.clone()

            def replay(points=None, static_in=static_in):
                if points is not None:
                    static_in.copy_(points, non_blocking=True)
                o = net.forward(static_in)
                return o, net.detections(o[0], o[1])
            runners.append(replay)
        else:
            runners.append(net.capture(pts))
    main = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_in = torch.from_numpy(pts_np).pin_memory()
    host_out = [torch.empty((SCENES_PER_GPU, 100, 9), dtype=torch.float32).pin_memory() for _ in range(P)]
    host_cnt = [torch.empty((SCENES_PER_GPU,), dtype=torch.int32).pin_memory() for _ in range(P)]

    def one_step(i, e2e):
        k = i % P
        outg, (b, c) = runners[k](host_in if e2e else None)      # e2e: pinned H2D into the graph's static input first
        if world > 1:
            b, c = pkg.dist.gather_detections(b, c)
            b, c = b[rank * SCENES_PER_GPU:(rank + 1) * SCENES_PER_GPU], c[rank * SCENES_PER_GPU:(rank + 1) * SCENES_PER_GPU]
        if e2e:
            host_out[k].copy_(b, non_blocking=True)
            host_cnt[k].copy_(c, non_blocking=True)

    def timed_run(nsteps, e2e, pipelined):
        """K steps; returns (total device ms from one event pair around the whole region, list of per-step ms)."""
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per = []
        start.record(main)
        for s_ in streams:
            s_.wait_event(start)
        for i in range(nsteps):
            s_ = streams[i % P] if pipelined else streams[0]
            with torch.cuda.stream(s_):
                flush.zero_()                                      # L2 flush between steps (inside the bracket)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s_)
                one_step(i if pipelined else 0, e2e)
                e1.record(s_)
                per.append((e0, e1))
        for s_ in streams:
            main.wait_stream(s_)
        end.record(main)
        barrier()
        return start.elapsed_time(end), [a.elapsed_time(b) for a, b in per]

    timed_run(args.warmup, False, True)
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    t_wall0 = time.perf_counter()
    total_ms, _ = timed_run(args.steps, False, True)
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    # single-step latency (no overlap between steps), for the record
    _, lat = timed_run(max(3, min(args.steps, 10)), False, False)
    latency_ms = float(np.median(lat))

    # ---- e2e: host buffers through the public API ---------------------------------------------------------
    timed_run(args.warmup, True, True)
    barrier()
    e2e_ms, _ = timed_run(args.steps, True, True)

    # ---- roofline of the dominant kernel: D-FPS layer 1, timed alone with events on its stream --------------
    xyz = pts[..., :3].contiguous()
    for _ in range(3):
        pkg.farthest_point_sample(4096, xyz)
    torch.cuda.synchronize()
    kev = []
    for _ in range(max(5, min(args.steps, 20))):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pkg.farthest_point_sample(4096, xyz)
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, k_ms = (float(v) for v in t.tolist())

    if rank == 0:
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fps_bytes = SCENES_PER_GPU * (4096 - 1) * NPOINTS * 16          # B*(M-1)*N*(4c+4), SURVEY.md 8d
        achieved = fps_bytes / (k_ms * 1e-3) / 1e9
        scenes = SCENES_PER_GPU * world * args.steps
        line = {
            "metric": METRIC, "value": scenes / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: full 3DSSD SA backbone (3dssd.yaml layer1-4 + vote) + detection head, decode and "
                                   "BEV NMS, synthetic KITTI 16384x4 clouds, batch 8 per GPU", "scenes_per_gpu": SCENES_PER_GPU,
                       "global_batch": SCENES_PER_GPU * world, "points": NPOINTS, "parallelism": "scene-sharded dp%d" % world,
                       "ffps": args.ffps_mode, "mlp": args.mlp_mode, "cuda_graph": not args.no_graph, "l2": "flushed (256 MiB write) before every timed step",
                       "steps_in_flight": P, "latency_ms_single_step": latency_ms, "fps_cluster_cap": args.fps_cluster_cap,
                       "timing": "one CUDA-event pair around all K steps (L2 flushes included), max over ranks; each step is a "
                                 "CUDA-graph replay, %d step pipelines on separate streams" % P,
                       "wall_s_bracket": wall},
            "e2e": {"value": scenes / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(host_in.numel() * 4), "d2h_bytes_per_step": int(host_out[0].numel() * 4 + host_cnt[0].numel() * 4)},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_dram_bytes("r01_ncu_fps_l1.txt"), "kernel": "fps3_direct_kernel (D-FPS layer 1, 16384->4096, B=8, cluster of %d CTAs per scene)" % (min(8, args.fps_cluster_cap) if args.fps_cluster_cap else 8),
                         "kernel_ms": k_ms,
                         "note": "effective-stream bytes B*(M-1)*N*16 (what the reference streams per round, SURVEY 8d); "
                                 "the kernel keeps them on-chip, so frac can exceed 1 and DRAM traffic is ~2 MB; peak = "
                                 + ("MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "6650 GB/s (of fallback)")},
        }
        line["roofline_tensor"] = tensor_roofline(torch, pkg, dev, peaks)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(arch, params, pts_np, args.cpu_scenes)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args, torch, pkg, arch, params, pts_np, rank, world, dev):
    """Reference arm (see module docstring).  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    from oracle import ref_ops
    cpu = cpu_baseline(arch, params, pts_np, args.cpu_scenes)
    if not ref_ops.available():
        line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": world,
                "steps": 1, "warmup": 0, "ms_per_step": 1e3 * SCENES_PER_GPU / cpu["value"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1] (CPU port: oracle/_ref not built)"},
                "cpu_baseline": cpu,
                "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return
    from oracle import ref_layers
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    pts = torch.from_numpy(pts_np).to(dev)
    cache = {}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for _ in range(max(1, min(args.warmup, 3))):
        ref_layers.backbone_forward(arch, pts, params, cache)
    torch.cuda.synchronize()
    steps = args.steps
    sampler = ClockSampler(dev.index or 0).start()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ref_layers.backbone_forward(arch, pts, params, cache)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    clocks = sampler.stop()
    val = SCENES_PER_GPU * steps / (tot * 1e-3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": args.warmup, "ms_per_step": tot / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: full 3DSSD SA backbone, synthetic KITTI 16384x4 clouds, batch 8",
                       "reference_device": "B200: the reference's own CUDA kernels (oracle/_ref/libref_ops.so, unmodified "
                                           "sources, nvcc -O2 sm_100) + PyTorch fp32 ops one-for-one for the TF stock ops; "
                                           "the reference has no CPU implementation of this path",
                       "l2": "flushed between timed steps"},
            "clocks": clocks,
            "cpu_baseline": cpu,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
