#!/usr/bin/env python
"""Per-kernel breakdown of FeatureNet0 (5 x B images of 768x384) with torch.profiler: which torch / cuDNN ops remain
around the native convolutions.   python tools/prof_featurenet_ops.py [--batch 8]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import models.adamvs as M
from adamvs_b200 import synth

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=8); a = ap.parse_args()
dev = torch.device("cuda:0")
sd = synth.fill_state_dict(synth.state_dict_shapes(48), 0)
m = M.Infer_AdaMVSNet(num_depth=192); m.load_state_dict(sd); m = m.to(dev).eval()
x = torch.randn(a.batch * 5, 3, 384, 768, device=dev)
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
    for _ in range(3):
        m.feature(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            m.feature(x)
        torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t and e.device_type.name == "CUDA":
        rows.append((t / 3e3, e.count // 3, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per forward: {tot:.3f} ms")
for t, n, k in rows[:30]:
    print(f"{t:8.3f} ms  x{n:3d}  {k[:110]}")
