#!/usr/bin/env python
"""Stand-alone driver of the C-ABI kernels at the bench shapes (B reference views, 5-view 768x384 stage
shapes), for ncu captures and quick CUDA-event timings without the cuDNN parts of the forward.

    python tools/prof_kernels.py [--batch 8] [--planes 3] [--iters 5] [--only k1,k2,k3]

Inputs are seeded random tensors with the synthetic camera rig (geometry matters for K1/K2's gather
locality).  K3 runs only `--planes` planes per stage (every plane launches the same kernels).
Prints one JSON line per kernel group with the CUDA-event time per call (after warm-up).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from adamvs_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--planes", type=int, default=3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="k1,k2,k3")
    ap.add_argument("--stages", default="1,2,3")
    a = ap.parse_args()
    only = set(a.only.split(","))
    dev = torch.device("cuda:0")
    B, H, W, V = a.batch, 384, 768, 5
    g = torch.Generator().manual_seed(0)
    _, proj, dv = synth.make_sample(B, H, W, V, seed=3)
    relproj, half = ops.cascade_prepare([proj[k].to(dev) for k in ("stage1", "stage2", "stage3")], dv.to(dev),
                                        ops.INTERVAL_FROM_RANGE, 192, [48, 32, 8], [4.0, 2.0, 1.0])
    sd = synth.fill_state_dict(synth.state_dict_shapes(48), 0)
    stages = [(32, 48, H // 4, W // 4, True), (16, 32, H // 2, W // 2, True), (8, 8, H, W, False)]

    def time_it(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i, (C, D, h, w, up) in enumerate(stages):
        if str(i + 1) not in a.stages.split(","):
            continue
        feat = torch.randn(B, V, C, h, w, generator=g).to(dev)
        wts = (torch.rand(B, V - 1, h, w, generator=g) * 0.9 + 0.05).to(dev)
        cur = (600 + 5 * torch.randn(B, h, w, generator=g)).to(dev)
        hyp = ops.Hyp(ops.HYP_PLANES, dv.to(dev)) if i == 0 else ops.Hyp(ops.HYP_PER_PIXEL, cur, half[i:i + 1])
        if i == 0 and "k1" in only:
            ms = time_it(lambda: (flush.zero_(), ops.pair_score(feat, relproj[0], hyp, D))[1])
            ms0 = time_it(lambda: flush.zero_())
            by = 4 * B * (V * C * h * w + (V - 1) * D * h * w)
            print(json.dumps({"kernel": "pair_score", "stage": 1, "B": B, "ms": ms - ms0, "GBps": by / (ms - ms0) * 1e-6}), flush=True)
        if "k2" in only:
            out = torch.empty((B, C, D, h, w), device=dev)
            ms = time_it(lambda: (flush.zero_(), ops.fused_volume(feat, relproj[i], hyp, wts, ops.EPS_DENOMINATOR, D, out=out))[1])
            ms0 = time_it(lambda: flush.zero_())
            by = 4 * B * (V * C * h * w + h * w + (V - 1) * h * w + C * D * h * w)
            print(json.dumps({"kernel": "fused_volume", "stage": i + 1, "B": B, "ms": ms - ms0, "GBps": by / (ms - ms0) * 1e-6,
                              "frac_hbm_6537": by / (ms - ms0) * 1e-6 / 6537.3}), flush=True)
        if "k3" in only:
            Dk = a.planes
            vol = torch.randn(B, C, Dk, h, w, generator=g).to(dev)
            p = f"DepthNet.{i}.reg_fuse"
            names = {"conv1_w": ".conv1.conv.weight", "gates1_w": ".conv_gru1.conv_gates.0.weight",
                     "gates1_b": ".conv_gru1.conv_gates.0.bias", "cand1_w": ".conv_gru1.convc.0.weight",
                     "cand1_b": ".conv_gru1.convc.0.bias", "conv2_w": ".conv2.conv.weight",
                     "gates2_w": ".conv_gru2.conv_gates.0.weight", "gates2_b": ".conv_gru2.conv_gates.0.bias",
                     "cand2_w": ".conv_gru2.convc.0.weight", "cand2_b": ".conv_gru2.convc.0.bias",
                     "up1_w": ".upconv1.weight", "up1_b": ".upconv1.bias", "out_w": ".upconv2d.weight",
                     "out_b": ".upconv2d.bias"}
            wd = {k: sd[p + v].to(dev) for k, v in names.items()}
            ws = torch.empty(ops.regnet_workspace_floats(B, C, Dk, h, w, up), device=dev)
            ms = time_it(lambda: ops.regnet_red(vol, wd, hyp, up, ops.PROB_EXP_EPS, workspace=ws))
            full, hf = h * w, (h // 2) * (w // 2)
            fl = B * Dk * 18 * (C * 8 * full + 256 * full + 128 * full + 128 * hf + 1024 * hf + 512 * hf + 128 * hf + 8 * full)
            print(json.dumps({"kernel": "regnet_red", "stage": i + 1, "B": B, "planes": Dk, "ms": ms, "ms_per_plane": ms / Dk,
                              "TFLOPs": fl / ms * 1e-9}), flush=True)


if __name__ == "__main__":
    main()
