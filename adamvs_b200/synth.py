"""Synthetic WHU-OMVS-shaped inputs and seeded weights (no dataset or checkpoint ships with the
reference: /root/reference/.MISSING_LARGE_BLOBS).

Everything here is derived from numpy's PCG64 streams so that the same seed gives the same bytes on
every machine and torch version; golden fixtures (tests/golden) only need to store seeds + outputs.

Input contract being synthesised (reference datasets/predict_oblique.py:154-177,
datasets/cas_total_rscv.py:513-549): imgs [B,V,3,H,W] zero-mean/unit-var per image,
proj_matrices {"stage1","stage2","stage3"} -> [B,V,4,4] with K[R|t] in the top three rows and rows
0-1 divided by 4 / 2 / 1, depth_values [B,2]=[min,max] (predict) or [B,3]=[min,max,interval] (train/test).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch

DEPTH_MIN = 520.0
DEPTH_MAX = 680.0
LOOK_AT_Z = 600.0
_BASE_X = (60.0, -45.0, 20.0, -30.0, 50.0, -55.0, 35.0, -25.0)
_BASE_Y = (10.0, 25.0, -40.0, -35.0, -20.0, 15.0, 45.0, -10.0)


def _rot_x(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def camera_rig(height: int, width: int, num_src: int = 4, jitter_seed: int | None = None):
    """Look-at rig: reference camera at the origin looking down +z, source cameras on baselines of up to 60 m, all aimed
    at (0, 0, 600).  Returns (K [3,3], [(R [3,3], t [3]) per view]) in float64, world -> camera (X right, Y down).

    With jitter_seed the baselines are perturbed by a few metres so that a batch of reference views
    does not share one geometry."""
    f = 1.2 * width
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], dtype=np.float64)
    rng = np.random.default_rng(1000003 + jitter_seed) if jitter_seed is not None else None
    poses = []
    for v in range(num_src + 1):
        if v == 0:
            R = np.eye(3)
            c = np.zeros(3)
        else:
            bx, by = _BASE_X[(v - 1) % 8], _BASE_Y[(v - 1) % 8]
            if rng is not None:
                bx += rng.uniform(-4, 4)
                by += rng.uniform(-4, 4)
            c = np.array([bx, by, 3.0 * v])
            yaw = math.atan2(-bx, LOOK_AT_Z)
            pitch = math.atan2(by, LOOK_AT_Z)
            R = (_rot_y(yaw) @ _rot_x(pitch)).T
        poses.append((R, -R @ c))
    return K, poses


def make_cameras(height: int, width: int, num_src: int = 4, jitter_seed: int | None = None) -> Dict[str, torch.Tensor]:
    """Projection matrices of camera_rig(): {"stageN": [V,4,4] float32} with K[R|t] in the top three rows and rows 0-1
    divided by 4 / 2 / 1."""
    K, poses = camera_rig(height, width, num_src, jitter_seed)
    mats = []
    for R, t in poses:
        P = np.eye(4)
        P[:3, :3] = K @ R
        P[:3, 3] = K @ t
        mats.append(P)
    full = np.stack(mats)                       # [V,4,4], stage3 (full resolution)
    out = {}
    for name, div in (("stage1", 4.0), ("stage2", 2.0), ("stage3", 1.0)):
        m = full.copy()
        m[:, :2, :] /= div
        out[name] = torch.from_numpy(m.astype(np.float32))
    return out


def make_images(num_views: int, height: int, width: int, seed: int) -> torch.Tensor:
    """[V,3,H,W] float32: smooth texture (x4 bicubic of low-res noise) plus fine noise, then per-image
    zero-mean / unit-variance like the reference's center_image (datasets/predict_oblique.py:59-64).
    All views show the same texture shifted by a few pixels so that matching costs carry signal."""
    rng = np.random.default_rng(seed)
    pad = 16
    lh, lw = (height + 2 * pad) // 4 + 2, (width + 2 * pad) // 4 + 2
    low = torch.from_numpy(rng.standard_normal((1, 3, lh, lw)).astype(np.float32))
    base = torch.nn.functional.interpolate(low, scale_factor=4, mode="bicubic", align_corners=False)[0]
    imgs = []
    for v in range(num_views):
        dy, dx = (v * 3) % 7, (v * 5) % 11
        crop = base[:, pad + dy: pad + dy + height, pad + dx: pad + dx + width]
        fine = torch.from_numpy(rng.standard_normal((3, height, width)).astype(np.float32))
        im = crop + 0.3 * fine
        im = (im - im.mean()) / (im.std() + 1e-8)
        imgs.append(im)
    return torch.stack(imgs).contiguous()


def make_sample(batch: int, height: int, width: int, num_views: int = 5, seed: int = 0,
                with_interval: bool = False, num_depth: int = 192):
    """One forward()'s worth of inputs: imgs [B,V,3,H,W], proj dict of [B,V,4,4], depth_values."""
    imgs = torch.stack([make_images(num_views, height, width, seed * 4099 + b) for b in range(batch)])
    cams = [make_cameras(height, width, num_views - 1, jitter_seed=(seed * 4099 + b) if (seed or b) else None)
            for b in range(batch)]
    proj = {k: torch.stack([c[k] for c in cams]).contiguous() for k in ("stage1", "stage2", "stage3")}
    if with_interval:
        dv = torch.tensor([[DEPTH_MIN, DEPTH_MAX, (DEPTH_MAX - DEPTH_MIN) / num_depth]] * batch, dtype=torch.float32)
    else:
        dv = torch.tensor([[DEPTH_MIN, DEPTH_MAX]] * batch, dtype=torch.float32)
    return imgs, proj, dv


def fill_state_dict(shapes: Dict[str, Sequence[int]], seed: int) -> Dict[str, torch.Tensor]:
    """Seeded values for a state_dict with the given key -> shape map (keys follow the reference's
    module tree). Conv weights/biases are U(-1/sqrt(fan_in), +) like torch's
    default; BatchNorm affine and running stats are made non-trivial on purpose so that eval-mode BN
    folding is exercised."""
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    fan_in_of: Dict[str, int] = {}
    for key in shapes:                      # first pass: fan-in per conv (shared with its bias)
        shp = tuple(shapes[key])
        if key.endswith("weight") and len(shp) == 4:
            fan_in_of[key[: -len("weight")]] = shp[1] * shp[2] * shp[3]
    for key in shapes:
        shp = tuple(shapes[key])
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_mean"):
            out[key] = torch.from_numpy((0.1 * rng.standard_normal(shp)).astype(np.float32))
        elif key.endswith("running_var"):
            out[key] = torch.from_numpy(rng.uniform(0.5, 1.5, shp).astype(np.float32))
        elif len(shp) == 4:
            bound = 1.0 / math.sqrt(fan_in_of[key[: -len("weight")]])
            out[key] = torch.from_numpy(rng.uniform(-bound, bound, shp).astype(np.float32))
        elif key.endswith("bias") and key[: -len("bias")] in fan_in_of:
            bound = 1.0 / math.sqrt(fan_in_of[key[: -len("bias")]])
            out[key] = torch.from_numpy(rng.uniform(-bound, bound, shp).astype(np.float32))
        elif key.endswith("weight"):        # BatchNorm gamma
            out[key] = torch.from_numpy(rng.uniform(0.8, 1.2, shp).astype(np.float32))
        else:                               # BatchNorm beta
            out[key] = torch.from_numpy((0.1 * rng.standard_normal(shp)).astype(np.float32))
    return out


def calibrate_state_dict(sd: Dict[str, torch.Tensor], feature_std: Dict[str, float], gain: float) -> Dict[str, torch.Tensor]:
    """scale the three 1x1 output convs so features are O(1) and sharpen the two
    logit-producing convs so probabilities are far from uniform. Without this every parity check is
    vacuous (fused volume ~1e-4 constant, p ~ 1/D). Returns a new dict."""
    sd = {k: v.clone() for k, v in sd.items()}
    for stage, key in (("stage1", "feature.out1.weight"), ("stage2", "feature.out2.weight"),
                       ("stage3", "feature.out3.weight")):
        sd[key] = sd[key] / float(feature_std[stage])
    for i in range(3):
        sd[f"DepthNet.{i}.reg_fuse.upconv2d.weight"] = sd[f"DepthNet.{i}.reg_fuse.upconv2d.weight"] * gain
        sd[f"DepthNet.{i}.reg.prob.weight"] = sd[f"DepthNet.{i}.reg.prob.weight"] * gain
    return sd


def state_dict_shapes(ndepths0: int = 48) -> Dict[str, List[int]]:
    """key -> shape for AdaMVSNet / Infer_AdaMVSNet (339 entries incl. num_batches_tracked),
    generated from the layer table below rather than from a module instance so that tests can build
    weights without constructing either implementation."""
    shapes: Dict[str, List[int]] = {}

    def bn(prefix: str, c: int):
        shapes[prefix + ".weight"] = [c]
        shapes[prefix + ".bias"] = [c]
        shapes[prefix + ".running_mean"] = [c]
        shapes[prefix + ".running_var"] = [c]
        shapes[prefix + ".num_batches_tracked"] = []

    def cbr(prefix: str, cin: int, cout: int, k: int):
        shapes[prefix + ".conv.weight"] = [cout, cin, k, k]
        bn(prefix + ".bn", cout)

    b = 8
    cbr("feature.conv0.0", 3, b, 3); cbr("feature.conv0.1", b, b, 3)
    cbr("feature.conv1.0", b, 2 * b, 5); cbr("feature.conv1.1", 2 * b, 2 * b, 3); cbr("feature.conv1.2", 2 * b, 2 * b, 3)
    cbr("feature.conv2.0", 2 * b, 4 * b, 5); cbr("feature.conv2.1", 4 * b, 4 * b, 3); cbr("feature.conv2.2", 4 * b, 4 * b, 3)
    cbr("feature.branch1_1.1", 4 * b, 2 * b, 1); cbr("feature.branch1_2.1", 4 * b, 2 * b, 1)
    shapes["feature.out1.weight"] = [4 * b, 8 * b, 1, 1]
    for name, cin, cout in (("feature.deconv1", 4 * b, 2 * b), ("feature.deconv2", 2 * b, b)):
        shapes[name + ".deconv.conv.weight"] = [cin, cout, 3, 3]
        bn(name + ".deconv.bn", cout)
        cbr(name + ".conv", 2 * cout, cout, 3)
    cbr("feature.branch2_1.1", 2 * b, b, 1); cbr("feature.branch2_2.1", 2 * b, b, 1)
    cbr("feature.branch3_1.1", b, b // 2, 1); cbr("feature.branch3_2.1", b, b // 2, 1)
    shapes["feature.out2.weight"] = [2 * b, 4 * b, 1, 1]
    shapes["feature.out3.weight"] = [b, 2 * b, 1, 1]
    n = ndepths0
    for i, c in enumerate((4 * b, 2 * b, b)):
        p = f"DepthNet.{i}.reg"
        for j in range(7):
            cbr(f"{p}.conv{j}", n, n, 3)
        for j in (7, 9, 11):
            shapes[f"{p}.conv{j}.0.weight"] = [n, n, 3, 3]
            bn(f"{p}.conv{j}.1", n)
        shapes[f"{p}.prob.weight"] = [n, n, 3, 3]
        shapes[f"{p}.prob.bias"] = [n]
        q = f"DepthNet.{i}.reg_fuse"
        shapes[f"{q}.conv1.conv.weight"] = [8, c, 3, 3]
        shapes[f"{q}.conv_gru1.conv_gates.0.weight"] = [16, 16, 3, 3]
        shapes[f"{q}.conv_gru1.conv_gates.0.bias"] = [16]
        shapes[f"{q}.conv_gru1.convc.0.weight"] = [8, 16, 3, 3]
        shapes[f"{q}.conv_gru1.convc.0.bias"] = [8]
        shapes[f"{q}.conv2.conv.weight"] = [16, 8, 3, 3]
        shapes[f"{q}.conv_gru2.conv_gates.0.weight"] = [32, 32, 3, 3]
        shapes[f"{q}.conv_gru2.conv_gates.0.bias"] = [32]
        shapes[f"{q}.conv_gru2.convc.0.weight"] = [16, 32, 3, 3]
        shapes[f"{q}.conv_gru2.convc.0.bias"] = [16]
        shapes[f"{q}.upconv1.weight"] = [16, 8, 3, 3]
        shapes[f"{q}.upconv1.bias"] = [8]
        shapes[f"{q}.upconv2d.weight"] = [8, 1, 3, 3] if i < 2 else [1, 8, 3, 3]
        shapes[f"{q}.upconv2d.bias"] = [1]
    return shapes


def msred_state_dict_shapes() -> Dict[str, List[int]]:
    """key -> shape for CascadeREDNet / Infer_CascadeREDNet (219 tensors + BatchNorm counters;
    reference models/msrednet.py:29-88, 134-148)."""
    shapes: Dict[str, List[int]] = {}

    def bn(prefix: str, c: int):
        shapes[prefix + ".weight"] = [c]
        shapes[prefix + ".bias"] = [c]
        shapes[prefix + ".running_mean"] = [c]
        shapes[prefix + ".running_var"] = [c]
        shapes[prefix + ".num_batches_tracked"] = []

    def cbr(prefix: str, cin: int, cout: int, k: int):
        shapes[prefix + ".conv.weight"] = [cout, cin, k, k]
        bn(prefix + ".bn", cout)

    b = 8
    cbr("feature.conv0.0", 3, b, 3); cbr("feature.conv0.1", b, b, 3)
    cbr("feature.conv1.0", b, 2 * b, 5); cbr("feature.conv1.1", 2 * b, 2 * b, 3); cbr("feature.conv1.2", 2 * b, 2 * b, 3)
    cbr("feature.conv2.0", 2 * b, 4 * b, 5); cbr("feature.conv2.1", 4 * b, 4 * b, 3); cbr("feature.conv2.2", 4 * b, 4 * b, 3)
    shapes["feature.out1.weight"] = [4 * b, 4 * b, 1, 1]
    for name, cin, cout in (("feature.deconv1", 4 * b, 2 * b), ("feature.deconv2", 2 * b, b)):
        shapes[name + ".deconv.conv.weight"] = [cin, cout, 3, 3]
        bn(name + ".deconv.bn", cout)
        cbr(name + ".conv", 2 * cout, cout, 3)
    shapes["feature.out2.weight"] = [2 * b, 2 * b, 1, 1]
    shapes["feature.out3.weight"] = [b, b, 1, 1]
    for i, c in enumerate((4 * b, 2 * b, b)):
        p = f"cost_regularization.{i}"
        for l, (x, hc) in enumerate(((c, 8), (16, 16), (32, 32), (64, 64)), start=1):
            g = f"{p}.conv_gru{l}"
            shapes[g + ".gate_conv.weight"] = [2 * hc, x + hc, 3, 3]
            shapes[g + ".gate_conv.bias"] = [2 * hc]
            for n in ("reset_gate_norm", "update_gate_norm"):
                shapes[f"{g}.{n}.weight"] = [hc]
                shapes[f"{g}.{n}.bias"] = [hc]
            shapes[g + ".output_conv.weight"] = [hc, x + hc, 3, 3]
            shapes[g + ".output_conv.bias"] = [hc]
            shapes[g + ".output_norm.weight"] = [hc]
            shapes[g + ".output_norm.bias"] = [hc]
        shapes[p + ".conv1.conv.weight"] = [16, c, 3, 3]
        shapes[p + ".conv2.conv.weight"] = [32, 16, 3, 3]
        shapes[p + ".conv3.conv.weight"] = [64, 32, 3, 3]
        shapes[p + ".upconv3.conv.weight"] = [64, 32, 3, 3]
        shapes[p + ".upconv2.conv.weight"] = [32, 16, 3, 3]
        shapes[p + ".upconv1.conv.weight"] = [16, 8, 3, 3]
        shapes[p + ".upconv2d.weight"] = [8, 1, 3, 3]
        shapes[p + ".upconv2d.bias"] = [1]
    return shapes


def calibrate_msred_state_dict(sd: Dict[str, torch.Tensor], feature_std: Dict[str, float], gain: float) -> Dict[str, torch.Tensor]:
    """MS-REDNet counterpart of calibrate_state_dict: O(1) features, sharpened output layer."""
    sd = {k: v.clone() for k, v in sd.items()}
    for stage, key in (("stage1", "feature.out1.weight"), ("stage2", "feature.out2.weight"),
                       ("stage3", "feature.out3.weight")):
        sd[key] = sd[key] / float(feature_std[stage])
    for i in range(3):
        sd[f"cost_regularization.{i}.upconv2d.weight"] = sd[f"cost_regularization.{i}.upconv2d.weight"] * gain
    return sd
