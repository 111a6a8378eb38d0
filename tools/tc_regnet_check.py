"""K3 on the tensor cores vs K3 on the FFMA kernels (the parity-tested path): differences and per-plane times.
usage: python tools/tc_regnet_check.py [--batch 8] [--planes 4] [--stages 1,2,3] [--small]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adamvs_b200 import ops, synth

NAMES = {"conv1_w": ".conv1.conv.weight", "gates1_w": ".conv_gru1.conv_gates.0.weight",
         "gates1_b": ".conv_gru1.conv_gates.0.bias", "cand1_w": ".conv_gru1.convc.0.weight",
         "cand1_b": ".conv_gru1.convc.0.bias", "conv2_w": ".conv2.conv.weight",
         "gates2_w": ".conv_gru2.conv_gates.0.weight", "gates2_b": ".conv_gru2.conv_gates.0.bias",
         "cand2_w": ".conv_gru2.convc.0.weight", "cand2_b": ".conv_gru2.convc.0.bias",
         "up1_w": ".upconv1.weight", "up1_b": ".upconv1.bias", "out_w": ".upconv2d.weight", "out_b": ".upconv2d.bias"}


def time_it(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--planes", type=int, default=4)
    ap.add_argument("--stages", default="1,2,3")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-time", action="store_true")
    a = ap.parse_args()
    sd = synth.fill_state_dict(synth.state_dict_shapes(8), 21)
    shapes = {1: (32, 96, 192, True), 2: (16, 192, 384, True), 3: (8, 384, 768, False)}
    if a.small:
        shapes = {1: (32, 16, 24, True), 2: (16, 32, 48, True), 3: (8, 64, 96, False)}
    for s in [int(x) for x in a.stages.split(",")]:
        C, h, w, up = shapes[s]
        B, D = a.batch, a.planes
        p = f"DepthNet.{s - 1}.reg_fuse"
        wd = {k: sd[p + v].cuda() for k, v in NAMES.items()}
        wd["out_w"] = wd["out_w"] * 20
        g = torch.Generator().manual_seed(s)
        vol = torch.randn(B, C, D, h, w, generator=g).cuda()
        cur = (600 + 10 * torch.randn(B, h, w, generator=g)).cuda()
        hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur, torch.tensor([3.3]).cuda())
        ref = ops.regnet_red(vol, wd, hyp, up, ops.PROB_SOFTMAX, want_logits=True, math=ops.MATH_FFMA)
        torch.cuda.synchronize()
        out = {"stage": s, "B": B, "D": D, "h": h, "w": w}
        for name, mode in (("tc_fp32", ops.MATH_TC_FP32), ("tc_tf32", ops.MATH_TC_TF32)):
            got = ops.regnet_red(vol, wd, hyp, up, ops.PROB_SOFTMAX, want_logits=True, math=mode)
            torch.cuda.synchronize()
            out[name] = {"logit_abs": float((got[2] - ref[2]).abs().max()), "logit_max": float(ref[2].abs().max()),
                         "depth_rel": float(((got[0] - ref[0]).abs() / ref[0].abs()).max()),
                         "conf_abs": float((got[1] - ref[1]).abs().max()),
                         "nan": int(torch.isnan(got[2]).sum())}
        if not a.no_time:
            ws = torch.empty(ops.regnet_workspace_floats(B, C, D, h, w, up), device="cuda")
            for name, mode in (("ffma", ops.MATH_FFMA), ("tc_fp32", ops.MATH_TC_FP32), ("tc_tf32", ops.MATH_TC_TF32)):
                ms = time_it(lambda: ops.regnet_red(vol, wd, hyp, up, ops.PROB_EXP_EPS, workspace=ws, math=mode))
                out["ms_per_plane_" + name] = ms / D
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
