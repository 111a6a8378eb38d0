#!/usr/bin/env python
"""GPU debug: stage-by-stage comparison of the MS-REDNet path (K5, K6) against the oracle at full size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from adamvs_b200 import synth, ops, cascade_msred
from oracle import msrednet_oracle as MO
from models.msrednet import Infer_CascadeREDNet

dev = torch.device("cuda:0")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (384, 768)
nd = (8, 4, 2)
imgs, proj, dv2 = synth.make_sample(1, H, W, 5, seed=23)
sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 37)
f = MO.feature_net(sd, imgs[:, 0])
sd = synth.calibrate_msred_state_dict(sd, {k: float(f[k].std()) for k in f}, 4.0)
co = {}
want = MO.infer_cascade_rednet_forward(sd, imgs, proj, dv2, num_depth=32, ndepths=nd, capture=co)
m = Infer_CascadeREDNet(num_depth=32, ndepths=list(nd), depth_interals_ratio=[4.0, 2.0, 1.0])
m.load_state_dict(sd); m = m.to(dev).eval()
cg = {}
out = cascade_msred.forward(m, imgs.to(dev), {k: v.to(dev) for k, v in proj.items()}, dv2.to(dev), capture=cg)
for i, s in enumerate(("stage1", "stage2", "stage3")):
    fe = (cg[s]["features"].cpu() - co[s]["features"]).abs().max()
    ve = (cg[s]["variance"].cpu() - co[s]["variance"]).abs()
    le = (cg[s]["logits"].cpu() - co[s]["logits"]).abs()
    print(s, "feat err %.2e" % fe, "var err %.2e (max %.2e)" % (ve.max(), co[s]["variance"].abs().max()),
          "logit err %.2e (max %.2e)" % (le.max(), co[s]["logits"].abs().max()),
          "depth rel %.2e" % ((out[s]["depth"].cpu() - want[s]["depth"]).abs() / want[s]["depth"]).max(),
          "prob %.2e" % (out[s]["photometric_confidence"].cpu() - want[s]["photometric_confidence"]).abs().max())
    # K6 alone on the oracle's variance volume
    hyps = co[s]["hyps"]
    holder = m.cost_regularization[i]
    # hypotheses: use the oracle's per-pixel centre (mean of first and last plane) for the regression only
    cur = ((hyps[:, 0] + hyps[:, -1]) / 2).to(dev).contiguous()
    hr = ((hyps[:, -1] - hyps[:, 0]) / 2).mean().reshape(1).to(dev)
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur, hr)
    _, _, lg = ops.regnet_msred(co[s]["variance"].to(dev), holder.kernel_weights(), hyp, ops.PROB_EXP_EPS, want_logits=True)
    le2 = (lg.cpu() - co[s]["logits"]).abs()
    idx = torch.nonzero(le2 == le2.max())[0].tolist()
    print("   K6 on oracle variance: logit err %.2e at %s" % (le2.max(), idx), " per-plane max:", [float(le2[:, k].max()) for k in range(le2.shape[1])])
    idx = torch.nonzero(ve == ve.max())[0].tolist()
    print("   K5 worst at", idx)
