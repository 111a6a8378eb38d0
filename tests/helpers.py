"""Shared fixtures-by-function for the test-suite: golden loading, weight/input reconstruction."""
import os

import numpy as np
import torch

from adamvs_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def rebuild_case(g):
    """Recreate (state_dict, imgs, proj, dv2, dv3, meta) from a golden file's seeds."""
    B, H, W = int(g["meta_B"]), int(g["meta_H"]), int(g["meta_W"])
    ndepths = tuple(int(x) for x in g["meta_ndepths"])
    num_depth = int(g["meta_num_depth"])
    imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=int(g["meta_iseed"]))
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
    dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
    sd = synth.fill_state_dict(synth.state_dict_shapes(ndepths[0]), int(g["meta_wseed"]))
    fstd = dict(zip(("stage1", "stage2", "stage3"), [float(x) for x in g["meta_fstd"]]))
    sd = synth.calibrate_state_dict(sd, fstd, float(g["meta_gain"]))
    return sd, imgs, proj, dv2, dv3, dict(B=B, H=H, W=W, ndepths=ndepths, num_depth=num_depth)


def rebuild_msred_case(g):
    """MS-REDNet counterpart of rebuild_case."""
    B, H, W = int(g["meta_B"]), int(g["meta_H"]), int(g["meta_W"])
    ndepths = tuple(int(x) for x in g["meta_ndepths"])
    num_depth = int(g["meta_num_depth"])
    imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=int(g["meta_iseed"]))
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
    dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), int(g["meta_wseed"]))
    fstd = dict(zip(("stage1", "stage2", "stage3"), [float(x) for x in g["meta_fstd"]]))
    sd = synth.calibrate_msred_state_dict(sd, fstd, float(g["meta_gain"]))
    return sd, imgs, proj, dv2, dv3, dict(B=B, H=H, W=W, ndepths=ndepths, num_depth=num_depth)


def make_case(B, H, W, ndepths, num_depth, gain, wseed, iseed, feature_fn):
    """Seeded calibrated case for sizes that have no golden file; feature_fn(sd, img)->dict gives
    the features used for calibration (oracle on CPU or product on GPU — both see the result)."""
    imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=iseed)
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
    dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
    sd = synth.fill_state_dict(synth.state_dict_shapes(ndepths[0]), wseed)
    f = feature_fn(sd, imgs[:, 0])
    fstd = {k: float(f[k].std()) for k in ("stage1", "stage2", "stage3")}
    sd = synth.calibrate_state_dict(sd, fstd, gain)
    return sd, imgs, proj, dv2, dv3


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def abs_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max())
