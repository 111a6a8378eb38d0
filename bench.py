#!/usr/bin/env python
"""Benchmark of the Ada-MVS cascade hot path (BASELINE.json metric: depth maps/sec, 5-view 768x384).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one forward of the drop-in ``Infer_AdaMVSNet`` (predict_whu.py's class) over a batch of B
reference views (1 ref + 4 src each, 3x384x768, ndepths 48/32/8, fp32) per GPU.  Views are
independent, so N GPUs run N such batches with no data-path collective (weak scaling); NCCL is used
only for the barrier and the max-over-ranks of the device time.

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput, ``e2e`` the same metric
through the public forward() with pinned-host inputs and a device->host read of the results inside
the timed region.  ``roofline`` is measured live with CUDA events around the fused cost-volume kernel
(K2, HBM-bound — the kernel BASELINE.json's metric names); ``kernels`` lists every C-ABI kernel's share
of the step and its own bound.  ``cpu_baseline`` / ``--impl reference`` time the oracle's CPU port of
the reference (the reference itself is Python and does not travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, V = 384, 768, 5
NDEPTHS = (48, 32, 8)
RATIOS = (4.0, 2.0, 1.0)
NUM_DEPTH = 192
MS_NDEPTHS = (128, 32, 8)          # BASELINE config 5: MS-REDNet, D = 128 planes at the first stage
MS_NUM_DEPTH = 512
METRIC = "depth maps/sec, 5-view 768x384"
UNIT = "depth_maps/s"
MS_WORKLOAD = ("configs[4]: MS-REDNet (Infer_CascadeREDNet) 5-view 768x384, ndepths 128/32/8, fp32, random-init weights; "
               "reference views sharded over ranks")
WORKLOAD = ("configs[1]: Ada-MVS 5-view (1 ref + 4 src) 768x384 cascade inference, fp32, random-init weights; "
            "reference views are independent and are sharded over ranks (configs[2])")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def costvolume_algorithmic_bytes(B, C, D, h, w, Vs=4, view_weights=True, weight_px=None):
    """every feature map read once, the hypothesis source once, the (Ada-MVS) view weights once at
    their source resolution (`weight_px` pixels per view: the stage-1 maps), the aggregated volume written once (fp32)."""
    wpx = (h * w if weight_px is None else weight_px) if view_weights else 0
    return 4 * B * ((1 + Vs) * C * h * w + h * w + Vs * wpx + C * D * h * w)


def regnet_flops(B, C, D, h, w):
    """Conv FLOPs of one recurrent-regulariser sweep (2*9*Cin*Cout per output pixel; transposed
    convs per input pixel)."""
    full, half = h * w, (h // 2) * (w // 2)
    per_plane = 18 * (C * 8 * full + 16 * 16 * full + 16 * 8 * full + 8 * 16 * half + 32 * 32 * half
                      + 32 * 16 * half + 16 * 8 * half + 8 * 1 * full)
    return B * D * per_plane


def regnet_tc_flops(B, C, D, h, w):
    """The part of regnet_flops that runs on the tensor cores (the five stride-1 convolutions; conv2 and the tail are FFMA)."""
    full, half = h * w, (h // 2) * (w // 2)
    return B * D * 18 * (C * 8 * full + 16 * 16 * full + 16 * 8 * full + 32 * 32 * half + 32 * 16 * half)


def msred_flops(B, C, D, h, w):
    """Conv FLOPs of one MS-REDNet regulariser sweep (reference models/msrednet.py:355-372)."""
    px = h * w
    macs = (C * 16 * px / 4 + 16 * 32 * px / 16 + 32 * 64 * px / 64                      # conv1..3 (stride 2)
            + (128 * 128 + 128 * 64) * px / 64 + 64 * 32 * px / 64                       # GRU4, upconv3
            + (64 * 64 + 64 * 32) * px / 16 + 32 * 16 * px / 16                           # GRU3, upconv2
            + (32 * 32 + 32 * 16) * px / 4 + 16 * 8 * px / 4                              # GRU2, upconv1
            + ((C + 8) * 16 + (C + 8) * 8) * px + 8 * px)                                 # GRU1, output layer
    return int(B * D * 18 * macs)


class ClockSampler:
    """nvidia-smi sampler running for the duration of the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_maps_per_s(n_maps=3, warmup=1, model="adamvs"):
    """The oracle's CPU port of the predict class' forward, all host threads, B=1 (reference's own batch size)."""
    import torch
    from adamvs_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    imgs, proj, dv = synth.make_sample(1, H, W, V, seed=0)
    if model == "adamvs":
        from oracle import adamvs_oracle as O
        sd = synth.fill_state_dict(synth.state_dict_shapes(NDEPTHS[0]), 0)
        f = O.feature_net(sd, imgs[:, 0])
        sd = synth.calibrate_state_dict(sd, {k: float(f[k].std()) for k in f}, 60.0)
        run = lambda: O.infer_adamvs_forward(sd, imgs, proj, dv, num_depth=NUM_DEPTH, ndepths=NDEPTHS, ratios=RATIOS)
    else:
        from oracle import msrednet_oracle as MO
        sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 0)
        f = MO.feature_net(sd, imgs[:, 0])
        sd = synth.calibrate_msred_state_dict(sd, {k: float(f[k].std()) for k in f}, 4.0)
        run = lambda: MO.infer_cascade_rednet_forward(sd, imgs, proj, dv, num_depth=MS_NUM_DEPTH, ndepths=MS_NDEPTHS, ratios=RATIOS)
    times = []
    for i in range(warmup + n_maps):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return 1.0 / statistics.median(times), torch.get_num_threads(), times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    steps = min(steps, 10)                                   # bounded sample: one depth map per step
    mps, cores, times = cpu_reference_maps_per_s(n_maps=steps, warmup=min(warm, 2), model=args.model)
    line = {
        "impl": "reference", "metric": METRIC, "value": mps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(warm, 2), "ms_per_step": 1e3 / mps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.model == "adamvs" else MS_WORKLOAD,
                   "class": "Infer_AdaMVSNet" if args.model == "adamvs" else "Infer_CascadeREDNet",
                   "ndepths": list(NDEPTHS if args.model == "adamvs" else MS_NDEPTHS),
                   "num_depth": NUM_DEPTH if args.model == "adamvs" else MS_NUM_DEPTH,
                   "views": V, "arm": "oracle torch-CPU port of the reference's PyTorch path, B=1 per step (the "
                                      "reference's own batch size), all host threads"},
        "cpu_baseline": {"value": mps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} depth maps (1 per step), median, after {min(warm, 2)} warm-up"},
        "e2e": {"value": mps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("ADAMVS_BENCH_BATCH", "32")),
                    help="reference views per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="adamvs", choices=["adamvs", "msrednet"],
                    help="adamvs = the headline (configs[1]/[2]); msrednet = BASELINE config 5")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from adamvs_b200 import ops, synth
    from models.adamvs import Infer_AdaMVSNet

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    min_warm = 1 if os.environ.get("ADAMVS_BENCH_PROFILING") else 3      # profiler passes only; never a bench value
    steps, warm, B = max(1, args.steps), max(min_warm, args.warmup), max(1, args.batch)

    # ---- model: seeded random-init weights, calibrated so that probabilities are not uniform
    msred = args.model == "msrednet"
    with open(os.devnull, "w") as devnull:
        stdout, sys.stdout = sys.stdout, devnull
        try:
            if msred:
                from models.msrednet import Infer_CascadeREDNet
                model = Infer_CascadeREDNet(num_depth=MS_NUM_DEPTH, ndepths=list(MS_NDEPTHS), depth_interals_ratio=list(RATIOS))
            else:
                model = Infer_AdaMVSNet(num_depth=NUM_DEPTH, ndepths=list(NDEPTHS), depth_intervals_ratio=list(RATIOS))
        finally:
            sys.stdout = stdout
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes() if msred else synth.state_dict_shapes(NDEPTHS[0]), 0)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    NSETS = 3                                                # distinct input batches, cycled (not L2-hot)
    host = []
    for s in range(NSETS):
        imgs, proj, dv = synth.make_sample(B, H, W, V, seed=1 + rank * NSETS + s)
        host.append((imgs.pin_memory(), {k: v.pin_memory() for k, v in proj.items()}, dv.pin_memory()))
    with torch.no_grad():
        f = model.feature(host[0][0][:1, 0].to(dev))
    fstd = {k: float(f[k].std()) for k in f}
    sd = synth.calibrate_msred_state_dict(sd, fstd, 4.0) if msred else synth.calibrate_state_dict(sd, fstd, 60.0)
    model.load_state_dict(sd)
    resident = [(i.to(dev), {k: v.to(dev) for k, v in p.items()}, d.to(dev)) for i, p, d in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        imgs, proj, dv = resident[i % NSETS]
        return model(imgs, proj, dv)

    out_host = [torch.empty((B, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]

    from adamvs_b200.pipeline import InputPrefetcher
    prefetch = InputPrefetcher(dev)

    def run_e2e(n):
        """n steps through forward() from pinned host memory: every step's H2D copy and D2H read-back is inside; the
        copy of step i+1 runs on a side stream while step i computes (adamvs_b200/pipeline.py)."""
        prefetch.stage(host[0])
        for i in range(n):
            imgs, proj, dv = prefetch.take()
            if i + 1 < n:
                prefetch.stage(host[(i + 1) % NSETS])
            out = model(imgs, proj, dv)
            out_host[0].copy_(out["depth"], non_blocking=True)
            out_host[1].copy_(out["photometric_confidence"], non_blocking=True)

    def timed(fn, n, whole=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if whole:
            fn(n)
        else:
            for i in range(n):
                fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        for i in range(warm):
            step_resident(i)
        # ---- timed region 1: device-resident inputs, per-kernel CUDA-event timers on, clocks sampled
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        timing = {}
        ops.set_timing(timing)
        ops.LAUNCHES[0] = 0
        ms_total = timed(step_resident, steps)
        launches = ops.LAUNCHES[0]
        ops.set_timing(None)
        clocks = sampler.stop() if rank == 0 else None
        # ---- timed region 2: end to end through forward() with pinned-host inputs and D2H of the result
        run_e2e(2)
        ms_e2e = timed(run_e2e, steps, whole=True)

    maps = world * B * steps
    value = maps / (ms_total * 1e-3)
    e2e_value = maps / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel accounting (rank 0): mean device time per launch from the CUDA events
    hbm_peak, tensor_peak, peak_src = _peaks()
    ffma_peak = 56.3                                         # TFLOP/s, tools/ffma_probe.cu on this pool (profiles/r01_ffma_probe.txt)
    per_kernel = {}
    for name, evs in timing.items():
        ms = [a.elapsed_time(b) for a, b in evs]
        per_kernel[name] = {"launches": len(ms), "ms_mean": sum(ms) / len(ms), "ms_total": sum(ms)}
    step_ms = ms_total / steps
    nd = MS_NDEPTHS if msred else NDEPTHS
    shapes = {"stage1": (32, nd[0], H // 4, W // 4), "stage2": (16, nd[1], H // 2, W // 2), "stage3": (8, nd[2], H, W)}
    kernels = {}
    for name, st in per_kernel.items():
        entry = {"ms_per_step": st["ms_total"] / steps, "share_of_step": st["ms_total"] / steps / step_ms}
        stage = name.split("/")[-1]
        if name.startswith("fused_volume/") or name.startswith("variance_volume/"):
            C, D, h, w = shapes[stage]
            by = costvolume_algorithmic_bytes(B, C, D, h, w, view_weights=not msred, weight_px=(H // 4) * (W // 4))
            entry.update({"bound": "hbm", "algorithmic_bytes": by, "achieved_GBps": by / st["ms_mean"] * 1e-6,
                          "frac_of_hbm_peak": by / st["ms_mean"] * 1e-6 / hbm_peak})
        elif name.startswith("regnet_red/") or name.startswith("regnet_msred/"):
            C, D, h, w = shapes[stage]
            fl = msred_flops(B, C, D, h, w) if msred else regnet_flops(B, C, D, h, w)
            entry.update({"flops": fl, "achieved_TFLOPs": fl / st["ms_mean"] * 1e-9,
                          "frac_of_ffma_peak": fl / st["ms_mean"] * 1e-9 / ffma_peak,
                          "frac_of_bf16_tensor_peak": fl / st["ms_mean"] * 1e-9 / tensor_peak})
            if msred:
                entry["bound"] = "fp32 FFMA (GroupNorm conv-GRU on the FFMA kernels)"
            else:
                # fp32 accuracy on kind::tf32: (A_hi + A_lo)(W_hi + W_lo) = 4 tf32 products per fp32 product
                tcf = regnet_tc_flops(B, C, D, h, w)
                entry.update({"bound": "tensor (tcgen05 kind::tf32, exact hi/lo operand split; conv2 + tail on FFMA)",
                              "tensor_share_of_flops": tcf / fl,
                              "tf32_TFLOPs_issued": 4 * tcf / st["ms_mean"] * 1e-9,
                              "frac_of_tf32_tensor_peak": 4 * tcf / st["ms_mean"] * 1e-9 / (tensor_peak / 2),
                              "tf32_peak_note": "dense tf32 taken as half the measured bf16 rate"})
        kernels[name] = entry
    # headline roofline: the fused warp + cost-volume kernel with the largest launch (stage 2)
    dom = max((k for k in kernels if k.startswith(("fused_volume/", "variance_volume/"))), key=lambda k: kernels[k]["algorithmic_bytes"])
    roofline = {"kernel": ("warp_volume_tma_kernel (K5 variance) " if msred else "warp_volume_tma_kernel (K2 fused volume) ") + dom.split("/")[-1], "bound": "hbm",
                "achieved": kernels[dom]["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"]}
    traffic_file = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                t = json.load(fh)
            if int(t.get("batch", -1)) == B:
                roofline["traffic"] = t.get("dram_bytes_per_launch")
                roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    h2d = sum(t.numel() * t.element_size() for t in (host[0][0], host[0][2])) + \
        sum(v.numel() * v.element_size() for v in host[0][1].values())
    d2h = 2 * B * H * W * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": MS_WORKLOAD if msred else WORKLOAD,
                   "class": "Infer_CascadeREDNet" if msred else "Infer_AdaMVSNet", "ndepths": list(nd),
                   "num_depth": MS_NUM_DEPTH if msred else NUM_DEPTH, "views": V, "batch_per_gpu_per_step": B,
                   "weights": "seeded random init, calibrated (SURVEY A.6)",
                   "l2": f"{NSETS} distinct input batches cycled; >1 GB of cost volume written/read per step (>> 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
    }
    if world == 1 and not args.no_cpu_baseline:
        mps, cores, times = cpu_reference_maps_per_s(n_maps=3, warmup=1, model=args.model)
        line["cpu_baseline"] = {"value": mps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "3 depth maps (B=1, same 5-view 768x384 workload), median, after 1 warm-up; "
                                          "oracle torch-CPU port of the predict class' forward"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
