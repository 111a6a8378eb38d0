#!/usr/bin/env python
"""GPU debug: per-layer outputs of FeatureNet0 with native convs vs cuDNN (both fp32, same device)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import models.adamvs as M
from adamvs_b200 import synth
dev = torch.device("cuda:0")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 192)
sd = synth.fill_state_dict(synth.state_dict_shapes(48), 5)
imgs, _, _ = synth.make_sample(1, H, W, 5, seed=2)
m = M.Infer_AdaMVSNet(num_depth=192); m.load_state_dict(sd); m = m.to(dev).eval()
x = imgs[0].to(dev)
acts = {}
def hook(name):
    def f(mod, inp, out):
        acts.setdefault(name, []).append(out.detach().clone() if torch.is_tensor(out) else None)
    return f
for n, mod in m.feature.named_modules():
    if isinstance(mod, (M._ConvBN, M._UpFuse)) or n in ("out1", "out2", "out3"):
        mod.register_forward_hook(hook(n))
with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
    M._NATIVE_CONV = True; m.feature(x)
    M._NATIVE_CONV = False; m.feature(x)
    # isolate each layer: feed the cuDNN run's input into the native layer
for n, (a, b) in acts.items():
    if a is None: continue
    d = (a - b).abs()
    idx = torch.nonzero(d == d.max())[0].tolist()
    print(f"{n:22s} shape {tuple(a.shape)} max|b| {float(b.abs().max()):9.3e} err {float(d.max()):9.3e} rel {float(d.max()/b.abs().max()):8.1e} at {idx}")
