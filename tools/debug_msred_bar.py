"""Pairwise differences between our MS-REDNet forward, the unmodified reference on CUDA (true fp32) and the reference on
CPU for bench.py's B=1 sample: which two agree?  (library_bar's parity block showed 0.36 between ours and reference-CUDA.)"""
import os, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from adamvs_b200 import synth
from baseline import refload
from models.msrednet import Infer_CascadeREDNet

wl = bench.WORKLOADS["msrednet"]
dev = torch.device("cuda:0")
imgs, proj, dv = bench._build_case(wl)
with contextlib.redirect_stdout(io.StringIO()):
    m = Infer_CascadeREDNet(num_depth=wl.num_depth, ndepths=list(wl.ndepths), depth_interals_ratio=list(bench.RATIOS)).to(dev).eval()
def feat(sd, img):
    m.load_state_dict(sd)
    with torch.no_grad():
        return m.feature(img.to(dev))
sd = bench._calibrated_state_dict(wl, imgs, feat)
m.load_state_dict(sd)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    ours = m(imgs.to(dev), {k: v.to(dev) for k, v in proj.items()}, dv.to(dev))
    ref = bench._reference_model(wl, {k: v.cpu() for k, v in sd.items()})
    rc = ref.to(dev)(imgs.to(dev), {k: v.to(dev) for k, v in proj.items()}, dv.to(dev))
    refc = bench._reference_model(wl, {k: v.cpu() for k, v in sd.items()})
    with refload.cpu_cuda_shim():
        rcpu = refc(imgs, proj, dv)
def d(a, b, s, k):
    return float((a[s][k].detach().cpu().double() - b[s][k].detach().cpu().double()).abs().max())
for s in ("stage1", "stage2", "stage3"):
    print(s, "prob |ours-refCUDA| %.3e  |ours-refCPU| %.3e  |refCUDA-refCPU| %.3e" % (d(ours, rc, s, "photometric_confidence"), d(ours, rcpu, s, "photometric_confidence"), d(rc, rcpu, s, "photometric_confidence")),
          " depth |ours-refCUDA| %.3e |ours-refCPU| %.3e |refCUDA-refCPU| %.3e" % (d(ours, rc, s, "depth"), d(ours, rcpu, s, "depth"), d(rc, rcpu, s, "depth")))
