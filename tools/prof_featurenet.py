#!/usr/bin/env python
"""Time FeatureNet0 and the stage-1 pair U-Net (the two cuDNN parts around the hot path) under a few
cuDNN settings: BN folding on/off, cudnn.benchmark on/off, channels_last on/off, TF32 on/off.
    python tools/prof_featurenet.py [--batch 8]"""
import argparse, itertools, json, os, sys
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

for fold, bench, cl, tf32 in itertools.product((True, False), (False, True), (False, True), (False, True)):
    if tf32 and not (fold and bench): continue
    M._FOLD_BN = fold
    xx = x.contiguous(memory_format=torch.channels_last) if cl else x
    ss = s.contiguous(memory_format=torch.channels_last) if cl else s
    mm = m.to(memory_format=torch.channels_last) if cl else m.to(memory_format=torch.contiguous_format)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=bench, deterministic=False, allow_tf32=tf32):
        try:
            tf = time_it(lambda: mm.feature(xx))
            tp = time_it(lambda: mm.DepthNet[0].reg(ss))
        except Exception as e:
            print(json.dumps({"fold": fold, "benchmark": bench, "channels_last": cl, "tf32": tf32, "error": str(e)[:200]})); continue
    print(json.dumps({"fold": fold, "benchmark": bench, "channels_last": cl, "tf32": tf32, "B": B,
                      "featurenet_ms": tf, "pair_unet_ms": tp}), flush=True)
