"""Which rows of the K3 logits differ from the oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adamvs_b200 import ops, synth
from oracle import adamvs_oracle as O
from tools.tc_regnet_check import NAMES

for (C, D, h, w, up) in ((8, 3, 64, 96, False), (8, 2, 34, 46, False), (16, 4, 32, 48, True)):
    B = 2
    g = torch.Generator().manual_seed(h + D)
    i = {32: 0, 16: 1, 8: 2}[C]
    sd = synth.fill_state_dict(synth.state_dict_shapes(8), 21)
    p = f"DepthNet.{i}.reg_fuse"
    sd[p + ".upconv2d.weight"] = sd[p + ".upconv2d.weight"] * 20
    vol = torch.randn(B, C, D, h, w, generator=g)
    cur = 600 + 10 * torch.randn(B, h, w, generator=g)
    want = O.red_regulariser(sd, p, vol, up)
    wd = {k: sd[p + v].cuda() for k, v in NAMES.items()}
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur.cuda(), torch.tensor([3.3]).cuda())
    outs = []
    for rep in range(2):
        d, c, lg = ops.regnet_red(vol.cuda(), wd, hyp, up, ops.PROB_SOFTMAX, want_logits=True)
        torch.cuda.synchronize()
        outs.append(lg.cpu())
    err = (outs[0] - want).abs()
    print((C, D, h, w, up), "max err", float(err.max()), "rep diff", float((outs[0] - outs[1]).abs().max()))
    rows = err.amax(dim=(0, 3))            # [D, H]
    for k in range(D):
        bad = [(y, round(float(rows[k, y]), 4)) for y in range(rows.shape[1]) if rows[k, y] > 1e-3]
        print("  plane", k, "bad rows:", bad[:12], "..." if len(bad) > 12 else "")
    cols = err.amax(dim=(0, 1, 2))
    print("  bad cols:", [x for x in range(cols.shape[0]) if cols[x] > 1e-3][:20])
