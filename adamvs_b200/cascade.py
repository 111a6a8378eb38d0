"""Host-side cascade: the glue of AdaMVSNet.forward / Infer_AdaMVSNet.forward
(reference models/adamvs.py:342-396 / 567-620) and of DepthNet0.forward / InferDepthNet0.forward
(:247-312 / :433-533), re-expressed as a short sequence of C-ABI kernel calls per stage:

    stage 1:  K1 pair_score -> pair U-Net (native tcgen05 / FFMA convs) -> K4 softmax_regress -> view weights
    stage s:  resize(view weights) -> K2 fused_volume -> K3 regnet_red (regression fused) -> depth, conf

No host synchronisation happens inside: range scalars and relative projections are produced on the
device by ``cascade_prepare`` (the reference syncs ~550 times per depth map through torch.inverse and
``float(depth_values[0,0].cpu())``).
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import ops

_STAGES = ("stage1", "stage2", "stage3")


def _true_fp32():
    # Layers that fall back to torch (channel counts the native kernels do not cover, e.g. the pair U-Net's stride-2 and
    # transposed convolutions when ndepths[0] != 48; everything under train()) must not run in cuDNN's default TF32, which
    # alone breaks the 1e-4 probability tolerance - and must be deterministic: cuDNN's default choice for
    # the transposed convolutions uses atomics, which made the stage-1 pair logits (and now and then a depth value in the
    # last bit) differ from run to run at ndepths[0] = 8 (tools/gpu_r2z2.sh; the all-native 48-plane path never did).
    return torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark,
                                      deterministic=True, allow_tf32=False)


def extract_features(net, imgs: torch.Tensor) -> Dict[str, torch.Tensor]:
    """imgs [B,V,3,H,W] -> {stage: [B,V,C,h,w]}.  Views are batched through FeatureNet0 in eval mode;
    in train mode BatchNorm statistics are per call in the reference, so views go one by one."""
    B, V = imgs.shape[:2]
    with _true_fp32():
        if net.training:
            per_view = [net.feature(imgs[:, v]) for v in range(V)]
            return {k: torch.stack([f[k] for f in per_view], 1).contiguous() for k in _STAGES}
        f = net.feature(imgs.reshape(B * V, *imgs.shape[2:]))
    return {k: f[k].reshape(B, V, *f[k].shape[1:]) for k in _STAGES}


def forward(net, imgs: torch.Tensor, proj_matrices: Dict[str, torch.Tensor], depth_values: torch.Tensor,
            capture: dict | None = None):
    if not imgs.is_cuda:
        raise ops.AdamvsError("adamvs_b200 runs on CUDA tensors only (no CPU fallback)")
    if torch.is_grad_enabled() and net.training:
        # train_whu.py --mode train: the differentiable forward over the forward / backward kernels (autograd.py)
        from . import autograd
        return autograd.forward_train(net, imgs, proj_matrices, depth_values)
    with torch.no_grad():
        return _forward(net, imgs, proj_matrices, depth_values, capture)


def _forward(net, imgs, proj_matrices, depth_values, capture):
    stream_conv = bool(net.stream_convention)
    ndepths = [int(d) for d in net.ndepths]
    ratios = [float(r) for r in net.depth_intervals_ratio]
    assert len(ndepths) == 3, "the cascade kernels are built for the reference's three stages"
    B, V = imgs.shape[:2]
    for k in _STAGES:
        assert proj_matrices[k].shape[1] == V, "Different number of images and projection matrices"

    imgs = imgs.float()
    depth_values = depth_values.float().contiguous()
    ops.set_tag("all")
    with ops.timed("featurenet", 0):
        feats = extract_features(net, imgs)
    relproj, half = ops.cascade_prepare(
        [proj_matrices[k] for k in _STAGES], depth_values,
        ops.INTERVAL_FROM_RANGE if stream_conv else ops.INTERVAL_LAST_COLUMN,
        getattr(net, "num_depth", 0), ndepths, ratios)

    eps_mode = ops.EPS_DENOMINATOR if stream_conv else ops.EPS_NUMERATOR
    prob_mode = ops.PROB_EXP_EPS if stream_conv else ops.PROB_SOFTMAX
    outputs: dict = {}
    depth = None
    stage1_w = None          # [B,Vs,h1,w1]
    prev_w = None            # weights as resized by the previous stage (predict class chains them)
    for i, key in enumerate(_STAGES):
        ops.set_tag(key)
        feat = feats[key]
        _, _, C, h, w = feat.shape
        D = ndepths[i]
        if depth is None:
            # first-stage planes run from depth_values[:,0] to [:,-1] (predict class, adamvs.py:569-570) or [:,-2]
            # (train/test class, :344-345: the last column is the interval) whatever the number of columns
            plane_range = torch.stack((depth_values[:, 0], depth_values[:, -1 if stream_conv else -2]), 1).contiguous()
            hyp = ops.Hyp(ops.HYP_PLANES, plane_range)
        else:
            assert tuple(depth.shape) == (B, h, w), (depth.shape, (B, h, w))
            hyp = ops.Hyp(ops.HYP_PER_PIXEL, depth, half[i:i + 1])
        pair_depths: List[torch.Tensor] = []
        if stage1_w is None:
            score = ops.pair_score(feat, relproj[i], hyp, D)                     # [B,Vs,D,h,w]
            with ops.timed("pair_unet", 0), _true_fp32():
                pair_logits = net.DepthNet[i].reg(score.reshape(B * (V - 1), D, h, w))
            pd, pc = ops.softmax_regress(pair_logits, hyp, ops.PROB_SOFTMAX, n_per_batch=V - 1)
            stage1_w = pc.reshape(B, V - 1, h, w)
            pair_depths = list(pd.reshape(B, V - 1, h, w).unbind(1))
            weights = stage1_w
            if capture is not None:
                capture.setdefault(key, {})["pair_score"] = score
                capture[key]["pair_logits"] = pair_logits.reshape(B, V - 1, D, h, w)
        else:
            src = prev_w if stream_conv else stage1_w
            weights = ops.resize_bilinear(src, h, w)
        prev_w = weights
        volume = ops.fused_volume(feat, relproj[i], hyp, weights, eps_mode, D)
        res = ops.regnet_red(volume, net.DepthNet[i].reg_fuse.kernel_weights(), hyp, out_up=(i < 2),
                             prob_mode=prob_mode, want_logits=capture is not None)
        depth, conf = res[0], res[1]
        if capture is not None:
            cap = capture.setdefault(key, {})
            cap["features"], cap["fused"], cap["logits"], cap["weights"] = feat, volume, res[2], weights
        views = [weights[:, v:v + 1] for v in range(V - 1)]
        if stream_conv:
            # list shape of the reference (adamvs.py:489-490, 506): stage 1 = the 4 maps + one resized
            # copy per (plane, view); later stages = one resized copy per (plane, view).
            pair_conf = (views if i == 0 else []) + views * D
        else:
            pair_conf = [stage1_w[:, v:v + 1] for v in range(V - 1)]
        out = {"depth": depth, "photometric_confidence": conf, "pair_confidence": pair_conf,
               "pair_result": pair_depths}
        outputs[key] = out
        outputs.update(out)
    return outputs
