#!/usr/bin/env python
"""Benchmark of the VoteNet point-cloud hot path (BASELINE.json: scenes/sec of the
Pointnet2Backbone forward, 40k points, batch 16, at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one Pointnet2Backbone forward (SA1-4 + FP1-2, eval mode) over one batch of 16
synthetic 40 000-point scenes with xyz + colour + normal + height (C = 7): BASELINE.json
configs[1].  Scenes shard across GPUs by batch index, 16 per GPU (weak scaling, the
reference's DDP recipe README.md:60-70); the forward needs no collective.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_POINTS = 40000
BATCH = 16
FEATURES = 7
METRIC = "scenes/sec VoteNet backbone fwd (40k pts, B=16)"
UNIT = "scenes/s"
WORKLOAD = "Pointnet2Backbone forward (SA1-4 2048/1024/512/256, FP1-2), 40000 pts xyz+color+normal+height (C=7), batch 16 per GPU, eval"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


# --------------------------------------------------------------------------- clocks ---

class ClockSampler(object):
    """nvidia-smi sampled every 20 ms; started before warm-up (the tool takes ~0.5 s to come
    up), reported over the samples taken between mark_begin() and mark_end(), i.e. DURING the
    timed region (B200_PROFILING.md recipe)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.path = None
        self.index = index
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        import datetime
        try:
            rows = []
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), f))
                except ValueError:
                    continue
            inside = [r for r in rows if self.t_begin is not None and self.t_begin - 0.02 <= r[0] <= self.t_end + 0.02]
            out["window"] = "timed region" if inside else "whole run (timed region shorter than one sample)"
            for ts, a, b_, f in (inside or rows):
                sm.append(a)
                mx.append(b_)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------ reference (CPU) ---

def rank_sm_clock(index):
    """Current SM clock (MHz) of GPU `index` through NVML; -1 when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        return float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    except Exception:
        return -1.0


def cpu_reference_run(steps, warmup, sample_scenes, threads=None):
    """The reference path re-expressed on the host (the reference ops are CUDA-only,
    sampling.cpp:83): C oracle ops + torch-CPU conv/BN, all host threads.  Each step is a
    bounded sample of the workload: `sample_scenes` scenes of the same 40k-point config."""
    import torch
    from bridgeqa_b200 import detector, synthetic
    from oracle import cpu_ops, modules_cpu
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_ops.set_num_threads(cores)
    net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=FEATURES), seed=0)
    sd = net.state_dict()
    pc = synthetic.make_batch(sample_scenes, NUM_POINTS, FEATURES).numpy()
    for _ in range(warmup):
        modules_cpu.backbone(pc[:1], sd)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        modules_cpu.backbone(pc, sd)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": sample_scenes * steps / total, "ms_per_step": 1e3 * total / steps,
            "cores": cores, "sample": "%d scene(s) of the workload per step, %d step(s)" % (sample_scenes, steps)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 3))
    r = cpu_reference_run(steps, min(args.warmup, 1), sample_scenes=4)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample_scenes_per_step": 4},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference pointnet2 ops are CUDA-only; this is the line-by-line host port (oracle/) "
                "timed on the box's CPU cores, per BASELINE.json north_star",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------- product arm ---

def algorithmic_work(name, dims):
    """Algorithmic bytes (and flops) of one C-ABI call, from its integer arguments.
    Formulas: DESIGN.md 'Kernels' / SURVEY.md 8d."""
    if name in ("bqa_furthest_point_sampling", "bqa_furthest_point_sampling_grid", "bqa_furthest_point_sampling_grid_lean",
                "bqa_furthest_point_sampling_cond"):
        b, n, m = dims[:3]
        return {"bytes": b * (12 * n + 4 * m + 12 * m), "bound": "hbm", "iters": m - 1, "units": b}
    if name == "bqa_ball_query":
        b, n, m, ns = dims[:4]
        return {"bytes": b * (12 * n + 12 * m + 4 * m * ns), "bound": "hbm", "pair_tests": b * n * m}
    if name == "bqa_ball_query_grid_search":
        b, n, m_total, j0, m, ns = dims[:6]
        return {"bytes": b * (12 * n + 12 * m + 4 * m * ns), "bound": "hbm", "pair_tests": b * n * m}
    if name == "bqa_bn_train_stats":
        b, c, l = dims[:3]
        return {"bytes": 4 * b * c * l, "bound": "hbm"}
    if name == "bqa_conv1x1_tf32_forward":        # read x, write y (the weights stay in L2); HBM-bound
        b, cin, cout, p = dims[:4]
        return {"bytes": 4 * b * p * (cin + cout), "flops": 2 * b * p * cin * cout, "bound": "hbm"}
    if name == "bqa_conv1x1_tf32_wgrad":          # read x and dy once
        b, cin, cout, p = dims[:4]
        return {"bytes": 4 * b * p * (cin + cout), "flops": 2 * b * p * cin * cout, "bound": "hbm"}
    if name == "bqa_bn_relu_forward":
        b, c, l = dims[:3]
        return {"bytes": 8 * b * c * l, "bound": "hbm"}
    if name == "bqa_bn_relu_max_forward":
        b, c, npt, ns = dims[:4]
        return {"bytes": 4 * b * c * npt * ns + 8 * b * c * npt, "bound": "hbm"}
    if name == "bqa_bn_relu_backward":          # stats pass reads dx, y; apply pass reads dx, y, writes dy
        b, c, l = dims[:3]
        return {"bytes": 20 * b * c * l, "bound": "hbm"}
    if name == "bqa_bn_relu_max_backward":      # dense pass: read y, write dy (dout / argmax are 1/nsample of that)
        b, c, npt, ns = dims[:4]
        return {"bytes": 8 * b * c * npt * ns + 12 * b * c * npt, "bound": "hbm"}
    if name == "bqa_group_concat_point_major":
        b, n, c, stride, npt, ns = dims[:6]
        return {"bytes": b * npt * ns * (4 + 8 * (c + 3)), "bound": "hbm"}
    if name == "bqa_fps_prefix_check":
        b, n, m = dims[:3]
        return {"bytes": b * (12 * n + 4), "bound": "hbm"}
    if name == "bqa_ball_query_grid_build":
        b, n = dims[:2]
        return {"bytes": b * (12 * n + 16 * n), "bound": "hbm"}       # read xyz, write the cell-sorted float4 copy
    if name in ("bqa_group_points", "bqa_group_points_grad", "bqa_group_points_grad_ws"):
        b, c, n, npt, ns = dims[:5]
        return {"bytes": b * (4 * npt * ns + 8 * c * npt * ns), "bound": "hbm"}
    if name in ("bqa_gather_points", "bqa_gather_points_grad"):
        b, c, n, m = dims[:4]
        return {"bytes": b * (4 * m + 8 * c * m), "bound": "hbm"}
    if name == "bqa_three_nn":
        b, n, m = dims[:3]
        return {"bytes": b * (12 * n + 12 * m + 24 * n), "bound": "hbm"}
    if name in ("bqa_three_interpolate", "bqa_three_interpolate_grad"):
        b, c, m, n = dims[:4]
        return {"bytes": b * (24 * n + 4 * c * m + 4 * c * n), "bound": "hbm"}
    if name == "bqa_transpose_to_point_major":
        b, c, n = dims[:3]
        return {"bytes": 8 * b * c * n, "bound": "hbm"}
    if name == "bqa_rows_to_16":                  # read c fp32 columns of every row, write stride 16-bit
        rows, c, row_stride, first, stride = dims[:5]
        return {"bytes": rows * (4 * c + 2 * stride), "bound": "hbm"}
    if name == "bqa_to_point_major_16":
        b, c, n, stride = dims[:4]
        return {"bytes": b * n * (4 * c + 2 * stride), "bound": "hbm"}
    if name in ("bqa_sa_mlp_max_forward", "bqa_sa_mlp_max_forward_v2"):
        b, n, npt, ns, c = dims[:5]
        c1, c2, c3 = dims[7:10]          # dims[5:7] = feat_stride, normalize_xyz
        flops = 2 * b * npt * ns * ((c + 3) * c1 + c1 * c2 + c2 * c3)
        byts = b * (4 * npt * ns + 4 * npt * ns * (c + 3) + 4 * c3 * npt)
        return {"bytes": byts, "flops": flops, "bound": "tensor"}
    if name == "bqa_fp_mlp_forward":
        b, n, m, ck, cs = dims[:5]
        c1, c2 = dims[7:9]               # dims[5:7] = row strides
        flops = 2 * b * n * ((ck + cs) * c1 + c1 * c2)
        byts = 4 * b * (ck * m + cs * n + c2 * n + 3 * n + 3 * m)
        return {"bytes": byts, "flops": flops, "bound": "tensor"}
    return {"bytes": 0, "bound": "hbm"}


def sms_occupied(name, dims, sms=148):
    """SMs a kernel holds while it runs (its CTAs are persistent or one per SM-sized chunk): used for the
    SM-time budget of the in-flight regime, where the step is bound by total SM-time, not by latency."""
    if name == "bqa_furthest_point_sampling_grid_lean":
        b, n = dims[:2]
        if 512 <= n <= 53440:          # one-SM kernel (fps_stream.cu): one CTA per scene
            return min(sms, b)
        return min(sms, b * -(-n // (768 * 18)))
    if name == "bqa_furthest_point_sampling_grid":
        b, n = dims[:2]
        return min(sms, b * (6 if n > 20480 else max(1, -(-n // 10240))))
    if name in ("bqa_sa_mlp_max_forward", "bqa_sa_mlp_max_forward_v2"):
        b, n, npt, ns = dims[:4]
        return min(sms, b * npt * ns // 128)
    if name == "bqa_fp_mlp_forward":
        b, n = dims[:2]
        return min(sms, b * -(-n // 128))
    return None


def measure_train(args, torch, dist, device, world, rank, steps, warmup, per_kernel=True, sample_clocks=True):
    """configs[3]: K training steps of the detector (C = 132, train-mode BN): fwd + bwd + gradient
    all-reduce (launched bucket by bucket from inside the backward pass when world > 1).  Returns the
    measurements as a dict (device time, max over ranks)."""
    from bridgeqa_b200 import _native, detector, distributed as D, fused as _fused, synthetic, training
    feats = 132
    net = synthetic.fill_state_dict(detector.VoteNetDetector(feats), seed=0).to(device)
    loss_fn = training.ProjectionLoss().to(device)
    bsz = args.train_batch
    pc = synthetic.make_batch(bsz, NUM_POINTS, feats, first_scene=rank * bsz).to(device)
    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True          # torch's (and the reference's) default
    torch.backends.cuda.matmul.allow_tf32 = True
    from bridgeqa_b200 import train_fused
    if args.torch_bn:
        train_fused.set_enabled(False)
    graphed = None
    if not getattr(args, "train_eager", False) and not args.no_prefetch and not args.torch_bn:
        # fwd + bwd replayed from two alternating CUDA graphs (the next batch's sampling is produced by the
        # current replay); the gradient buckets are all-reduced after the replay
        graphed = training.GraphedTrainStep(net, loss_fn, pc)
        reducer = graphed.reducer
    else:
        reducer = D.OverlappedGradReducer(net) if world > 1 else None

    def one_step(next_pc):
        if graphed is not None:
            return graphed(pc, next_pc)
        return training.train_step(net, loss_fn, pc, next_point_clouds=next_pc, reducer=reducer)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # the next batch's SA1 sampling (coordinates only) is issued under this step's backward
    nxt = None if args.no_prefetch else pc
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if sample_clocks else None
    if rank == 0 and clocks:
        clocks.start()
    for _ in range(warmup):
        one_step(nxt)
    barrier()
    l0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.mark_begin()
    e0.record()
    for _ in range(steps):
        loss = one_step(nxt)
    # K steps = K samplings inside the timed region: the first consumed the warm-up's prefetch,
    # the last one's prefetch has to finish before the clock stops
    torch.cuda.current_stream(device).wait_stream(_fused.side_stream(device, "prefetch"))
    e1.record()
    barrier()
    clk = None
    if clocks:
        clocks.mark_end()
        clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    launches = _native.launch_count() - l0
    if graphed is not None:
        launches = graphed.launches_per_step * steps     # kernel nodes of this library the replays re-issue
    # the exchange alone: the reducer's buckets, all-reduced back to back (what the backward pass hides)
    ar_ms, ar_bytes = 0.0, 0
    if reducer is not None and world > 1:
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(3):
            if rep == 1:
                torch.cuda.synchronize()
                a0.record()
            for flat, _members in reducer.buckets:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = a0.elapsed_time(a1) / 2.0
        ar_bytes = reducer.bytes
    nparam = sum(p.numel() for p in net.parameters())
    res = {"ms": ms, "steps": steps, "warmup": warmup, "launches": launches, "loss": float(loss), "clk": clk,
           "bsz": bsz, "nparam": nparam, "allreduce_ms": ar_ms, "allreduce_bytes": ar_bytes,
           "issue": ("2 alternating CUDA graphs (training.GraphedTrainStep); buckets all-reduced after the replay"
                     if graphed is not None else "eager; buckets all-reduced from the backward hooks"),
           "conv": "tcgen05 TF32 (conv_tf32.cu)" if train_fused.conv_enabled() else "cuDNN TF32",
           "fused_bn_relu": bool(train_fused.enabled())}
    if per_kernel:
        # per-kernel pass (one more step, CUDA events around every C-ABI call)
        from bridgeqa_b200 import profiler
        with profiler.KernelTimer() as kt:
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            training.train_step(net, loss_fn, pc, reducer=None if graphed is not None else reducer)
            k1.record()
            barrier()
        res["kern"] = kt.summary()
        res["kpass_ms"] = k0.elapsed_time(k1)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
    if args.torch_bn:
        train_fused.set_enabled(True)
    del net, pc, reducer, graphed
    torch.cuda.empty_cache()
    return res


def run_train_mode(args, torch, dist, device, world, rank, real_stdout):
    """--mode train: the configs[3] line on its own."""
    res = measure_train(args, torch, dist, device, world, rank, args.steps, args.warmup)
    ms, kern, kpass_ms, bsz, nparam, clk, launches, loss = (res["ms"], res["kern"], res["kpass_ms"], res["bsz"],
                                                            res["nparam"], res["clk"], res["launches"], res["loss"])
    if rank == 0:
        peaks = measured_peaks()
        groups = {}
        for key, d in kern.items():
            w = algorithmic_work(d["name"], d["dims"])
            g = groups.setdefault(d["name"], {"kernel": d["name"], "calls_per_step": 0, "ms": 0.0, "bytes": 0})
            g["calls_per_step"] += d["calls"]
            g["ms"] += d["ms"]
            g["bytes"] += w.get("bytes", 0) * d["calls"]
        kernels = []
        for g in groups.values():
            row = {"kernel": g["kernel"], "calls_per_step": g["calls_per_step"], "ms": round(g["ms"], 4),
                   "share": round(g["ms"] / kpass_ms, 4)}
            if g["bytes"]:
                ach = g["bytes"] / (g["ms"] / 1e3) / 1e9
                row.update(bound="hbm", achieved=round(ach, 1), peak=peaks["hbm_gbs"], unit="GB/s",
                           frac=round(ach / peaks["hbm_gbs"], 4))
            kernels.append(row)
        kernels.sort(key=lambda r: -r["ms"])
        line = {"metric": "scenes/sec DET train step (fwd+bwd+grad all-reduce), 40k pts, C=132",
                "value": bsz * world * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
                "data": "synthetic",
                "config": {"workload": "VoteNetDetector fwd+bwd, train-mode BN, %d scenes/GPU: sm_100a operators "
                                       "(sampling, grid ball query, grouping, 1x1 convs, BatchNorm+ReLU+max "
                                       "fwd/bwd), NCCL all-reduce of %d fp32 gradients in 2 buckets launched "
                                       "from the backward pass" % (bsz, nparam),
                           "convs": res["conv"], "fused_bn_relu": res["fused_bn_relu"], "issue": res["issue"],
                           "sampling_prefetch": not args.no_prefetch},
                "allreduce": {"ms_alone": round(res["allreduce_ms"], 4), "bytes": res["allreduce_bytes"]},
                "clocks": clk,
                "kernels": kernels, "kernel_pass_ms": round(kpass_ms, 3),
                "gpu_launches": launches, "loss": loss}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="force the un-fused operator path")
    ap.add_argument("--tf32", action="store_true", help="allow TF32 in torch convs of the un-fused path")
    ap.add_argument("--workload", default="backbone", choices=["backbone", "detector"],
                    help="backbone = configs[1] (headline); detector = configs[2]: backbone + voting + "
                         "proposal with 132-d point features")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of "
                    "replaying the captured CUDA graph")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward = BASELINE headline (configs[1]); train = DET train step, configs[3]")
    ap.add_argument("--train-batch", type=int, default=16, help="scenes per GPU in --mode train")
    ap.add_argument("--no-prefetch", action="store_true", help="--mode train: sample each batch in front of its "
                    "forward instead of under the previous step's backward")
    ap.add_argument("--torch-bn", action="store_true", help="--mode train: torch BatchNorm/ReLU/max_pool modules "
                    "instead of the sm_100a streaming kernels (the reference's module structure)")
    ap.add_argument("--in-flight", type=int, default=12,
                    help="batches in flight (graphs.InFlight): consecutive steps are issued on a ring of this many "
                         "streams so the next batch's sampling chain runs under this batch's SA/FP kernels; 1 = serial")
    ap.add_argument("--bind-cpu", action="store_true",
                    help="bind this rank (and the pinned buffers it then allocates) to the CPUs NVML reports as "
                         "local to its GPU -- for the end-to-end number at 8 GPUs, which is bound by host <-> device "
                         "copies (experimental: not measured at 8 GPUs yet)")
    ap.add_argument("--e2e-input", default="staged16", choices=["staged16", "fp32"],
                    help="end-to-end leg: what the loader hands over on the host.  staged16 = staging.StagedCloud "
                         "(fp32 xyz + 16-bit point-major features, the buffer the fused SA1 kernel gathers from); "
                         "fp32 = the reference's (B,N,3+C) fp32 cloud")
    ap.add_argument("--e2e-readback", default="consumer", choices=["consumer", "features16", "features32"],
                    help="end-to-end leg: what is read back to the host every step.  consumer = what the next stage "
                         "on the HOST needs (seed coordinates / indices + per-scene feature checksum; the seed "
                         "features are consumed on the device by voting / proposal); features16 / features32 add "
                         "fp2_features as fp16 / fp32")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` block (configs[3]: a few DET training "
                    "steps of the C=132 detector, fwd + bwd + gradient all-reduce) of the default line")
    ap.add_argument("--train-eager", action="store_true", help="issue the training step launch by launch (gradient "
                    "all-reduce from the backward hooks) instead of replaying training.GraphedTrainStep")
    ap.add_argument("--train-steps", type=int, default=5, help="timed steps of the `train` block (2 warm-up steps)")
    ap.add_argument("--no-ref-ext", action="store_true", help="skip the `ref_ext` leg (stock reference modules on the "
                    "reference's own CUDA extension, timed in a subprocess on the same GPU)")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="operand format of the fused tensor-core kernels (fp32 accumulate)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner)
    # are sent to stderr until the line is ready
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import bridgeqa_b200
    from bridgeqa_b200 import _native, detector, profiler, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpu_binding = None
    if args.bind_cpu:
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            cpu_binding = sorted(os.sched_getaffinity(0))
        except Exception as e:      # cpuset of the container may not contain the GPU's CPUs
            cpu_binding = "failed: %r" % (e,)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    args.warmup = max(args.warmup, 3)
    if args.unfused:
        bridgeqa_b200.set_fused(False)
    bridgeqa_b200.set_precision(args.precision)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)

    _native.lib()   # fail loudly here if the CUDA library is missing
    if args.mode == "train":
        return run_train_mode(args, torch, dist, device, world, rank, real_stdout)
    features = FEATURES if args.workload == "backbone" else 132
    if args.workload == "backbone":
        net = synthetic.fill_state_dict(detector.Pointnet2Backbone(input_feature_dim=features), seed=0)
        out_key = "fp2_features"
    else:
        net = synthetic.fill_state_dict(detector.VoteNetDetector(features), seed=0)
        out_key = "bbox_corner"
    net = net.to(device).eval()
    if not args.no_graph:
        net.enable_cuda_graph(bind_inputs=True)   # public API (graphs.py): replay instead of ~35 launches/step

    # inputs: 16 scenes per rank, resident in HBM in ROT rotated variants so that a step's
    # input was last touched ROT-1 steps (and >> 126 MB of other traffic) ago
    ROT = 8 if args.workload == "backbone" else 4     # a 132-d batch is 346 MB: larger than L2 on its own
    if args.in_flight > ROT:
        ROT = args.in_flight                          # one resident input (and graph) per forward in flight
    host_batch = synthetic.make_batch(BATCH, NUM_POINTS, features, first_scene=rank * BATCH)
    host_pinned = [torch.roll(host_batch, shifts=997 * i, dims=1).contiguous().pin_memory() for i in range(ROT)]
    dev_inputs = [h.to(device) for h in host_pinned]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)

    depth = max(1, min(args.in_flight, ROT))
    queue = net.in_flight(depth)       # public API (graphs.InFlight): `depth` forwards in flight

    def step(i, q=None):
        if q is not None:              # step i and step i + ROT share a graph: ROT % depth == 0 keeps them on one stream
            return q.submit({"point_clouds": dev_inputs[i % ROT]})
        with torch.no_grad():
            return net({"point_clouds": dev_inputs[i % ROT]})[out_key]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    rank_sm_clock(local_rank)          # NVML initialised here, not inside the timed region
    for i in range(max(args.warmup, ROT if not args.no_graph else 0)):   # every input buffer's graph exists
        step(i)
    for i in range(max(args.warmup, ROT) if depth > 1 else 0):
        step(i, queue)
    queue.drain()
    barrier()

    # ---- timed region: exactly K steps, device-timed ----
    clocks.mark_begin()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_host0 = time.perf_counter()
    my_sm_mhz = -1.0
    for i in range(args.steps):
        step(i, queue if depth > 1 else None)
        if i == args.steps // 2:
            my_sm_mhz = rank_sm_clock(local_rank)     # this rank's GPU, mid-run (NVML, host side only)
    host_issue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    queue.drain()                      # the timing stream waits for every forward in flight
    ev1.record()
    barrier()
    clocks.mark_end()
    elapsed_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None

    # ---- the same K steps one batch at a time (latency of a batch; not the headline) ----
    serial_ms = elapsed_ms
    if depth > 1:
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(args.steps):
            step(i)
        s1.record()
        barrier()
        serial_ms = s0.elapsed_time(s1)

    # ---- per-kernel pass: the same K steps again with a CUDA-event pair around every C-ABI
    # call (on the stream it is launched on).  Kept out of the pass above because ~50 extra
    # event records per step cost host time that a 2.5 ms step notices.
    if not args.no_graph:
        net.enable_cuda_graph(False)       # per-call events need the launches issued one by one
    launches0 = _native.launch_count()
    # (same sampling variant as the timed region: the queue's forwards run the throughput one)
    with profiler.KernelTimer() as kt, bridgeqa_b200.fused.lean_sampling(queue.lean):
        kp0 = torch.cuda.Event(enable_timing=True)
        kp1 = torch.cuda.Event(enable_timing=True)
        kp0.record()
        for i in range(args.steps):
            step(i)
        kp1.record()
        barrier()
    kern = kt.summary()
    kernel_pass_ms = kp0.elapsed_time(kp1)
    # what an event pair around a C-ABI call reads when the call launches nothing (zero rows): the host-side gap
    # between the two records under eager issue.  Rows of kernels[] shorter than a few times this carry it.
    import ctypes as _ct
    floor = []
    for _ in range(21):
        fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        fa.record()
        _native.call("bqa_rows_to_16", _ct.c_longlong(0), 7, 10, 3, 8, 1, None, None,
                     _native.stream_ptr(device))
        fb.record()
        torch.cuda.synchronize()
        floor.append(fa.elapsed_time(fb))
    event_pair_floor_ms = sorted(floor)[len(floor) // 2]
    launches = _native.launch_count() - launches0   # == kernel nodes the graph replays for K steps
    if not args.no_graph:
        net.enable_cuda_graph(True, bind_inputs=True)

    # ---- e2e: host (pinned) -> device -> forward -> host, copies inside the timed region ----
    from bridgeqa_b200 import staging
    if args.workload == "detector":
        rb_keys, rb_half = ["bbox_corner", "objectness_scores", "sem_cls_scores", "aggregated_vote_features"], ()
    else:
        rb_keys, rb_half = ["fp2_xyz", "fp2_inds", "fp2_checksum"], ()
        if args.e2e_readback == "features16":
            rb_keys, rb_half = rb_keys + ["fp2_features"], ("fp2_features",)
        elif args.e2e_readback == "features32":
            rb_keys = rb_keys + ["fp2_features"]
    if args.e2e_input == "staged16":
        # the loader emits the staged format (host work outside the timed region, like any loader work)
        host_e2e = [staging.stage_host(h, pin=True) for h in host_pinned]
        h2d_bytes = host_e2e[0].nbytes()
    else:
        host_e2e = host_pinned
        h2d_bytes = host_pinned[0].numel() * 4
    out_host = {}

    # pipelined: NB = depth + 1 device input buffers, each with its own graph and static outputs
    # (bind_inputs).  The H2D copy of step i runs on copy_in as soon as the forward that last read
    # its buffer (step i - NB) is done; the forward goes through the in-flight queue; the D2H read
    # of its results runs on copy_out straight out of the graph's output buffers, and the next
    # forward on that buffer (step i + NB) waits for the read.  Every copy is inside the timed region.
    copy_in, copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    main = torch.cuda.current_stream(device)
    NB = depth + 1
    if args.e2e_input == "staged16":
        dev_buf = [staging.StagedCloud.empty_like(host_e2e[0], device) for _ in range(NB)]
    else:
        dev_buf = [torch.empty_like(dev_inputs[0]) for _ in range(NB)]

    def e2e_run(k):
        fwd_done = [None] * NB
        read_done = [None] * NB
        start = torch.cuda.Event()
        start.record(main)
        copy_in.wait_event(start)
        for i in range(k):
            j = i % NB
            with torch.cuda.stream(copy_in):
                if fwd_done[j] is not None:
                    copy_in.wait_event(fwd_done[j])                # forward i-NB no longer reads the buffer
                dev_buf[j].copy_(host_e2e[i % ROT], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_in)
            ticket = queue.submit({"point_clouds": dev_buf[j]},
                                  after=[ready] + ([read_done[j]] if read_done[j] is not None else []))
            fwd_done[j] = ticket.done
            with torch.cuda.stream(copy_out), torch.no_grad():
                dd = ticket.wait(copy_out)
                if "fp2_checksum" in rb_keys:
                    # per-scene checksum of the seed features (B, 256): the "metric" a host-side consumer of
                    # the backbone would log; the features themselves stay on the device for voting / proposal
                    dd = dict(dd)
                    dd["fp2_checksum"] = dd["fp2_features"].mean(dim=2)
                staging.read_back(dd, rb_keys, out=out_host, half=rb_half)
                read_done[j] = torch.cuda.Event()
                read_done[j].record(copy_out)
        main.wait_stream(copy_out)
        main.wait_stream(copy_in)

    e2e_run(2 * NB)
    barrier()
    # the host link alone: the same pinned input copied back to back, no kernels (explains an end-to-end
    # number that is bound by the copy: the boxes of the pool differ by 2x here)
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_in):
        ca.record(copy_in)
        for i in range(8):
            dev_buf[i % NB].copy_(host_e2e[i % ROT], non_blocking=True)
        cb.record(copy_in)
    copy_in.synchronize()
    h2d_alone_gbps = 8 * h2d_bytes / (ca.elapsed_time(cb) * 1e-3) / 1e9
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host.values())
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # max over ranks
    t = torch.tensor([elapsed_ms, e2e_ms, my_sm_mhz, serial_ms], dtype=torch.float64, device=device)
    by_rank = [elapsed_ms / args.steps]
    mhz_by_rank = [my_sm_mhz]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        by_rank = [float(x[0]) / args.steps for x in every]       # device time of each rank's K steps
        mhz_by_rank = [float(x[2]) for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, serial_ms = float(t[0]), float(t[1]), float(t[3])

    # configs[3] in the same line: a few training steps on every rank (the all-reduce is a collective)
    train_res = None
    if not args.no_train and args.workload == "backbone" and args.train_steps > 0:
        try:
            train_res = measure_train(args, torch, dist, device, world, rank, args.train_steps, 2,
                                      per_kernel=False, sample_clocks=False)
        except Exception as e:       # the forward line must survive a failing extra
            train_res = {"error": repr(e)[:300]}
    if rank == 0:
        peaks = measured_peaks()
        scenes = BATCH * world * args.steps
        value = scenes / (elapsed_ms / 1e3)
        kernels = []
        kernel_time_ms = sum(d["ms"] for d in kern.values()) or 1.0
        for key, d in kern.items():
            w = algorithmic_work(d["name"], d["dims"])
            ms = d["ms"] / d["calls"]
            # share: of the pass's elapsed time (streams overlap, the host leaves gaps);
            # share_of_kernel_time: of the summed kernel times -- the figure an ncu launch list gives
            row = {"kernel": d["name"], "dims": d["dims"], "calls_per_step": d["calls"] / args.steps,
                   "ms": round(ms, 5), "share": round(d["ms"] / kernel_pass_ms, 4),
                   "share_of_kernel_time": round(d["ms"] / kernel_time_ms, 4), "bound": w["bound"]}
            if w["bound"] == "tensor":
                ach = w["flops"] / (ms / 1e3) / 1e12
                peak = peaks["bf16_tflops_sustained"]
                row.update(achieved=round(ach, 2), peak=peak, unit="TFLOP/s", frac=round(ach / peak, 4))
            else:
                ach = w["bytes"] / (ms / 1e3) / 1e9
                row.update(achieved=round(ach, 2), peak=peaks["hbm_gbs"], unit="GB/s",
                           frac=round(ach / peaks["hbm_gbs"], 5))
            if "iters" in w and w["iters"] > 0:
                row["us_per_iter"] = round(1e3 * ms / w["iters"], 4)
            if "pair_tests" in w:
                row["gpairs_per_s"] = round(w["pair_tests"] / (ms / 1e3) / 1e9, 1)
            kernels.append(row)
        kernels.sort(key=lambda r: -r["share"])
        top = kernels[0] if kernels else None
        roofline = None
        traffic = None
        prof_name = None
        try:   # dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture
            prof_name = "r2_kernels_ncu.json" if os.path.exists(os.path.join(ROOT, "profiles", "r2_kernels_ncu.json")) \
                else "r1_kernels_ncu.json"
            prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
            names = {"bqa_furthest_point_sampling": "fps_cluster_kernel<14", "bqa_furthest_point_sampling_grid": "fps_sorted_kernel<14",
                     "bqa_furthest_point_sampling_grid_lean": "fps_stream_kernel"}
            if top and top["kernel"] in names and top["dims"][:3] == [BATCH, NUM_POINTS, 2048]:
                hit = [k for k in prof if k["kernel"].startswith(names[top["kernel"]])]
                if hit:
                    traffic = hit[0]["dram_read_bytes"] + hit[0]["dram_write_bytes"]
        except Exception:
            traffic = None
        if top:
            roofline = {"kernel": "%s%s" % (top["kernel"], top["dims"]), "bound": top["bound"],
                        "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                        "frac": top["frac"], "traffic": traffic,
                        "traffic_source": ("committed ncu --set full capture of the same kernel and shape "
                                           "(profiles/%s), NOT measured by this run" % prof_name) if traffic else None,
                        "algorithmic_bytes": algorithmic_work(top["kernel"], top["dims"])["bytes"],
                        "peak_source": peaks["source"],
                        "share_of_step": top["share"], "share_of_kernel_time": top["share_of_kernel_time"],
                        "ms": top["ms"]}
            if "us_per_iter" in top:
                roofline["us_per_iter"] = top["us_per_iter"]
                roofline["note"] = ("FPS is a serial chain of npoint-1 cluster-wide argmax steps: "
                                    "latency-bound, HBM fraction is reported for the contract only")
        line = {
            "metric": METRIC if args.workload == "backbone" else
            "scenes/sec VoteNet full detector fwd (40k pts, 132-d features, B=16)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16" if bridgeqa_b200.fused.precision() == "fp16" else "bf16")
                     if any(r["kernel"] == "bqa_sa_mlp_max_forward" for r in kernels) else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "backbone" else
                       "full detector forward: backbone + VotingModule + ProposalModule (256 proposals, r=0.3, "
                       "nsample=16, on-device box decode), 40000 pts with 132-d features, batch 16 per GPU, eval",
                       "global_batch": BATCH * world, "num_points": NUM_POINTS,
                       "parallelism": "scenes sharded by batch index, %d/GPU, no collective in forward" % BATCH,
                       "l2": "inputs rotate over %d resident batches (%.0f MB) + >1 GB intermediate traffic per step"
                             % (ROT, ROT * h2d_bytes / 1e6),
                       "fused": any(r["bound"] == "tensor" for r in kernels), "torch_tf32": bool(args.tf32),
                       "cuda_graph": not args.no_graph,
                       "in_flight": depth, "lean_sampling": bool(queue.lean), "cpu_binding": cpu_binding,
                       "in_flight_note": "consecutive steps are issued on a ring of %d streams (graphs.InFlight): each step "
                                         "is still one forward over one batch of 16 scenes; the next batch's sampling chain "
                                         "(latency-bound; throughput variant of the kernel: 48 SMs) runs under this batch's SA/FP kernels; "
                                         "serial.ms_per_step is one batch at a time" % depth},
            "serial": {"ms_per_step": serial_ms / args.steps, "value": scenes / (serial_ms / 1e3), "unit": UNIT,
                       "note": "same K steps with one batch in flight (latency variant of the sampling kernel) = latency of a batch"},
            "e2e": {"value": scenes / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "h2d_alone_gb_per_s": round(h2d_alone_gbps, 2),
                    "h2d_alone_ms_per_step": round(h2d_bytes / h2d_alone_gbps / 1e6, 4),
                    "input": ("staging.StagedCloud: fp32 xyz + 16-bit point-major features (pinned host)"
                              if args.e2e_input == "staged16" else "(B,N,3+C) fp32 cloud (pinned host)"),
                    "read_back": {k_: list(v_.shape) + [str(v_.dtype)] for k_, v_ in out_host.items()}},
            "ms_per_step_by_rank": [round(v, 4) for v in by_rank],
            "sm_mhz_by_rank": mhz_by_rank,
            "gpu_launches": launches,
            "host_issue_ms_per_step": round(host_issue_ms, 3),
            "kernel_pass": {"ms_per_step": round(kernel_pass_ms / args.steps, 4),
                            "event_pair_floor_ms": round(event_pair_floor_ms, 4),
                            "note": "kernels[] and roofline come from a second pass of the same K steps with "
                                    "CUDA events around every C-ABI call; shares are relative to that pass.  "
                                    "Issued eagerly, the host leaves event_pair_floor_ms between the two records even "
                                    "when nothing is launched: rows of a few hundredths of a ms are upper bounds "
                                    "(ncu durations of the same kernels: profiles/r2_kernels_ncu.json)"},
            "roofline": roofline,
            "kernels": kernels,
            "clocks": clk,
        }
        # SM-time budget of the timed (in-flight) regime: the step is bound by total SM-time there, so each
        # kernel's cost is (SMs it holds) x (its duration); durations are the eager pass's (same kernels)
        budget = []
        for r in kernels:
            occ = sms_occupied(r["kernel"], r["dims"])
            if occ:
                budget.append({"kernel": r["kernel"], "dims": r["dims"][:5], "sms": occ,
                               "sm_ms": round(occ * r["ms"] * r["calls_per_step"], 3)})
        line["sm_time_budget"] = {
            "step_sm_ms": round(148 * elapsed_ms / args.steps, 2),
            "note": "timed regime: ms_per_step x 148 SMs; rows = SMs held x kernel duration for the kernels that "
                    "hold whole SMs (sampling clusters, persistent SA / FP CTAs); the rest are short kernels",
            "kernels": sorted(budget, key=lambda r_: -r_["sm_ms"])}
        if train_res is not None:
            if "error" in train_res:
                line["train"] = train_res
            else:
                tms = train_res["ms"] / train_res["steps"]
                line["train"] = {
                    "metric": "scenes/sec DET train step (fwd+bwd+grad all-reduce), 40k pts, C=132 (configs[3])",
                    "value": train_res["bsz"] * world / (tms / 1e3), "unit": UNIT, "ms_per_step": round(tms, 4),
                    "steps": train_res["steps"], "warmup": train_res["warmup"], "scenes_per_gpu": train_res["bsz"],
                    "n_gpus": world, "scaling": "weak", "dtype": "tf32", "convs": train_res["conv"],
                    "fused_bn_relu": train_res["fused_bn_relu"], "gpu_launches": train_res["launches"],
                    "issue": train_res["issue"],
                    "allreduce": {"ms_alone": round(train_res["allreduce_ms"], 4), "bytes": train_res["allreduce_bytes"],
                                  "how": "2 flat fp32 buckets, NCCL all-reduce launched from the backward pass "
                                         "(distributed.OverlappedGradReducer); ms_alone = the same buckets reduced "
                                         "back to back outside a step"},
                    "loss": train_res["loss"],
                    "note": "device time, max over ranks"}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=1, warmup=0, sample_scenes=4)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"],
                                    "kind": "port", "sample": r["sample"]}
        if world == 1 and not args.no_ref_ext and args.workload == "backbone":
            # the reference's OWN code on this GPU: unmodified Python layer + its CUDA extension (oracle/_ref),
            # in a subprocess so that none of it is loaded into this process
            try:
                torch.cuda.synchronize()
                rr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_ext_forward.py"),
                                     "--steps", "3", "--warmup", "1", "--features", str(features)],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=240)
                ref = json.loads(rr.stdout.strip().splitlines()[-1])
                if "value" in ref:
                    ref["speedup_device_resident"] = round(value / ref["value"], 2)
                if "checksum_fp2_features" in ref:
                    # same weights (seed 0), same batch (scenes 0-15): mean |fp2_features| of this repo's forward
                    with torch.no_grad():
                        mine = float(net({"point_clouds": dev_inputs[0]})["fp2_features"].double().abs().mean())
                    ref["checksum_fp2_features_this_repo"] = mine
                    ref["checksum_rel_diff"] = abs(mine - ref["checksum_fp2_features"]) / abs(ref["checksum_fp2_features"])
                line["ref_ext"] = ref
            except Exception as e:
                line["ref_ext"] = {"unavailable": repr(e)[:200]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
