#!/usr/bin/env python
"""Generate tests/golden/grad_small_d8.npz: loss, per-stage outputs and ALL parameter gradients of the UNMODIFIED
reference (/root/reference models.adamvs.AdaMVSNet, CPU, model.train()) for one training step's forward/backward on
seeded synthetic inputs - what pins the training path.  Build container only; the fixture is committed.

    python tests/golden/make_grad_golden.py

The step mirrors train_whu.py:265-300: outputs = model(imgs, proj_matrices, depth_values);
loss, _ = cas_mvs_vis_loss(outputs, depth_gt_ms, mask_ms, dlossw=[0.5, 1.0, 2.0]); loss.backward().
Accommodation: Tensor.cuda is the identity for the duration (the reference hard-codes .cuda() in its forwards)."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASE = dict(B=1, H=64, W=96, ndepths=(8, 4, 2), num_depth=32, gain=20.0, wseed=11, iseed=3, gtseed=101)
DLOSSW = [0.5, 1.0, 2.0]


def ground_truth(B, H, W, seed):
    """Smooth synthetic depth inside the search range and an all-valid mask at the three ground-truth resolutions
    (H/4, H/2, H: datasets/cas_total_rscv.py builds the same pyramid)."""
    rng = np.random.default_rng(seed)
    low = torch.from_numpy(rng.uniform(560.0, 640.0, (B, 1, H // 16 + 2, W // 16 + 2)).astype(np.float32))
    full = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False)
    gt, mask = {}, {}
    for k, s in (("stage1", 4), ("stage2", 2), ("stage3", 1)):
        g = torch.nn.functional.interpolate(full, size=(H // s, W // s), mode="bilinear", align_corners=False)[:, 0]
        gt[k], mask[k] = g.contiguous(), torch.ones_like(g)
    return gt, mask


def main():
    warnings.filterwarnings("ignore")
    from adamvs_b200 import synth
    from baseline import refload
    ref = refload.load("adamvs")
    torch.set_num_threads(os.cpu_count())
    c = CASE
    imgs, proj, dv2 = synth.make_sample(c["B"], c["H"], c["W"], 5, seed=c["iseed"])
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / c["num_depth"]
    dv3 = torch.cat([dv2, torch.full((c["B"], 1), interval)], 1)
    sd = synth.fill_state_dict(synth.state_dict_shapes(c["ndepths"][0]), c["wseed"])
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.AdaMVSNet(ndepths=list(c["ndepths"]), depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        f = m.feature(imgs[:, 0])
    fstd = {k: float(f[k].std()) for k in ("stage1", "stage2", "stage3")}
    sd = synth.calibrate_state_dict(sd, fstd, c["gain"])
    m.load_state_dict(sd)
    m.train()
    gt, mask = ground_truth(c["B"], c["H"], c["W"], c["gtseed"])
    with refload.cpu_cuda_shim():
        out = m(imgs, proj, dv3)
        loss, depth_loss = ref.cas_mvs_vis_loss(out, gt, mask, dlossw=DLOSSW)
        loss.backward()
    blob = {"meta_" + k: np.array(v) for k, v in c.items()}
    blob["meta_fstd"] = np.array([fstd["stage1"], fstd["stage2"], fstd["stage3"]], dtype=np.float64)
    blob["loss"] = np.array(float(loss))
    for s in ("stage1", "stage2", "stage3"):
        blob[f"{s}_depth"] = out[s]["depth"].detach().numpy()
        blob[f"{s}_conf"] = out[s]["photometric_confidence"].detach().numpy()
    n = 0
    for name, p in m.named_parameters():
        if p.grad is not None:
            blob["grad_" + name] = p.grad.numpy()
            n += p.grad.numel()
    path = os.path.join(HERE, "grad_small_d8.npz")
    np.savez_compressed(path, **blob)
    print("loss", float(loss), "| parameters with a gradient:", sum(k.startswith("grad_") for k in blob), "tensors,", n, "values ->",
          path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
