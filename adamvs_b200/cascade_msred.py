"""Host-side cascade of MS-REDNet: the glue of CascadeREDNet.forward / Infer_CascadeREDNet.forward
(reference models/msrednet.py:279-334 / 473-525) and DepthNet.forward / InferDepthNet.forward
(:203-242 / :379-436) as a short sequence of C-ABI kernel calls per stage:

    stage s:  hypotheses (planes at stage 1; previous depth resized to the image and back to this stage's grid)
              -> K5 variance_volume -> K6 regnet_msred (regression fused) -> depth, conf at this stage's grid

The reference builds the hypothesis tensor at full image resolution and resamples it with a trilinear
``F.interpolate`` to [D, h, w] (:316-318, 510-512).  The depth axis keeps its size, so that resampling
is a bilinear (align_corners=False) down-scaling of every plane; hypotheses are affine in the per-pixel
centre depth, hence resampling the centre depth map instead and generating ``lo + k*step`` in registers is
the same function up to fp32 rounding (checked against the reference golden in the tests).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops
from .cascade import _STAGES, _true_fp32


def _features(net, imgs: torch.Tensor) -> Dict[str, torch.Tensor]:
    B, V = imgs.shape[:2]
    with _true_fp32():
        if net.training:
            per_view = [net.feature(imgs[:, v]) for v in range(V)]
            return {k: torch.stack([f[k] for f in per_view], 1).contiguous() for k in _STAGES}
        f = net.feature(imgs.reshape(B * V, *imgs.shape[2:]))
    return {k: f[k].reshape(B, V, *f[k].shape[1:]) for k in _STAGES}


def forward(net, imgs, proj_matrices, depth_values, capture: dict | None = None):
    if not imgs.is_cuda:
        raise ops.AdamvsError("adamvs_b200 runs on CUDA tensors only (no CPU fallback)")
    if torch.is_grad_enabled() and net.training:
        raise NotImplementedError("adamvs_b200 has forward kernels only so far: call under torch.no_grad() or model.eval()")
    with torch.no_grad():
        return _forward(net, imgs, proj_matrices, depth_values, capture)


def _forward(net, imgs, proj_matrices, depth_values, capture):
    stream_conv = bool(net.stream_convention)
    ndepths = [int(d) for d in net.ndepths]
    ratios = [float(r) for r in net.depth_interals_ratio]
    assert len(ndepths) == 3, "the cascade kernels are built for the reference's three stages"
    B, V = imgs.shape[:2]
    H, W = int(imgs.shape[3]), int(imgs.shape[4])
    for k in _STAGES:
        assert proj_matrices[k].shape[1] == V, "Different number of images and projection matrices"
    imgs = imgs.float()
    depth_values = depth_values.float().contiguous()
    ops.set_tag("all")
    with ops.timed("featurenet", 0):
        feats = _features(net, imgs)
    relproj, half = ops.cascade_prepare(
        [proj_matrices[k] for k in _STAGES], depth_values,
        ops.INTERVAL_FROM_RANGE if stream_conv else ops.INTERVAL_LAST_COLUMN,
        getattr(net, "num_depth", 0), ndepths, ratios)
    # stage-1 planes run from column 0 to the LAST column of what the reference passes as cur_depth: [min,max] for
    # the predict class, all of [min,max,interval] for the train/test class (msrednet.py:308; module.py:651-653)
    plane_range = torch.stack((depth_values[:, 0], depth_values[:, -1]), 1).contiguous()
    prob_mode = ops.PROB_EXP_EPS if stream_conv else ops.PROB_SOFTMAX
    outputs: dict = {}
    depth = None
    for i, key in enumerate(_STAGES):
        ops.set_tag(key)
        feat = feats[key]
        _, _, C, h, w = feat.shape
        D = ndepths[i]
        if depth is None:
            hyp = ops.Hyp(ops.HYP_PLANES, plane_range)
        else:
            cur = depth if tuple(depth.shape[1:]) == (H, W) else ops.resize_bilinear(depth, H, W)
            cur = cur if (h, w) == (H, W) else ops.resize_bilinear(cur, h, w)
            hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur, half[i:i + 1])
        volume = ops.variance_volume(feat, relproj[i], hyp, D)
        res = ops.regnet_msred(volume, net.cost_regularization[i].kernel_weights(), hyp, prob_mode,
                               want_logits=capture is not None)
        depth, conf = res[0], res[1]
        if capture is not None:
            capture[key] = {"features": feat, "variance": volume, "logits": res[2]}
        out = {"depth": depth, "photometric_confidence": conf}
        outputs[key] = out
        outputs.update(out)
    return outputs
