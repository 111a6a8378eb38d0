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


def make_blendedmvs_scene(root, height=64, width=96, views=5, num_depth=32, seed=4):
    """A tiny training scene in the BlendedMVS layout the reference's datasets/cas_total_rscv.py reads
    (BlendedMVS_list :164-207, tr_read_blendedmvs_cam :357-386): index.txt, <scene>/blended_images/%08d.jpg,
    <scene>/cams/%08d_cam.txt (Tcw, K, depth start / interval / count / end), <scene>/cams/pair.txt,
    <scene>/rendered_depth_maps/%08d.pfm.  Synthetic rig and texture of adamvs_b200.synth; two reference views."""
    from PIL import Image
    from adamvs_b200 import sceneio as S
    scene = os.path.join(root, "scene0")
    for d in ("blended_images", "cams", "rendered_depth_maps"):
        os.makedirs(os.path.join(scene, d), exist_ok=True)
    with open(os.path.join(root, "index.txt"), "w") as f:
        f.write("scene0\n")
    K, poses = synth.camera_rig(height, width, views - 1)
    imgs = synth.make_images(views, height, width, seed)                       # [V,3,H,W], zero mean / unit variance
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
    for v in range(views):
        a = imgs[v].permute(1, 2, 0).numpy()
        a = np.clip(a * 48.0 + 128.0, 0, 255).astype(np.uint8)
        Image.fromarray(a).save(os.path.join(scene, "blended_images", f"{v:08d}.jpg"), quality=95)
        R, t = poses[v]
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        with open(os.path.join(scene, "cams", f"{v:08d}_cam.txt"), "w") as f:
            f.write("extrinsic\n" + "\n".join(" ".join(f"{x:.10f}" for x in row) for row in T) + "\n\n")
            f.write("intrinsic\n" + "\n".join(" ".join(f"{x:.10f}" for x in row) for row in K) + "\n\n")
            f.write(f"{synth.DEPTH_MIN} {interval} {num_depth} {synth.DEPTH_MAX}\n")
    with open(os.path.join(scene, "cams", "pair.txt"), "w") as f:
        f.write("2\n")
        for ref in (0, 1):
            others = [v for v in range(views) if v != ref]
            f.write(f"{ref}\n{len(others)} " + " ".join(f"{v} 1.0" for v in others) + "\n")
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    for ref in (0, 1):
        depth = (600.0 + 25.0 * np.sin(xx / 17.0 + ref) * np.cos(yy / 13.0) + rng.uniform(-1, 1)).astype(np.float32)
        S.write_pfm(os.path.join(scene, "rendered_depth_maps", f"{ref:08d}.pfm"), depth)
    return root
