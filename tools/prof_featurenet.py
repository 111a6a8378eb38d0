#!/usr/bin/env python
"""Time FeatureNet0 and the stage-1 pair U-Net (the 2-D networks around the hot path) with their 3x3 convolutions on
adamvs_b200's FFMA kernels vs all-cuDNN fp32 (and cuDNN TF32 for reference).   python tools/prof_featurenet.py [--batch 8]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import models.adamvs as M
from adamvs_b200 import synth

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=8); a = ap.parse_args()
dev = torch.device("cuda:0")
sd = synth.fill_state_dict(synth.state_dict_shapes(48), 0)
m = M.Infer_AdaMVSNet(num_depth=192); m.load_state_dict(sd); m = m.to(dev).eval()
B = a.batch
x = torch.randn(B * 5, 3, 384, 768, device=dev)
s = torch.randn(B * 4, 48, 96, 192, device=dev)

def time_it(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for native, tf32 in ((True, False), (False, False), (False, True)):
    M._NATIVE_CONV = native
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=tf32):
        tf = time_it(lambda: m.feature(x))
        tp = time_it(lambda: m.DepthNet[0].reg(s))
    print(json.dumps({"native_3x3": native, "cudnn_tf32": tf32, "B": B, "featurenet_ms": tf, "pair_unet_ms": tp}), flush=True)
