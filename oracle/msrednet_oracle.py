"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (torch fp32) restatement of MS-REDNet's cascade (BASELINE config 5) as pure functions over a flat
``state_dict``; same rules as ``oracle/adamvs_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may import it.

Parity status: PINNED by ``tests/golden/msred_*.npz`` — outputs of the reference's own ``CascadeREDNet`` /
``Infer_CascadeREDNet`` executed from /root/reference by ``tests/golden/make_golden.py``.
Each function names the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from .adamvs_oracle import SD, _conv_bn_relu, _deconv_bn_relu, depth_hypotheses, expected_depth, homography_warp


def feature_net(sd: SD, img: torch.Tensor, prefix: str = "feature") -> Dict[str, torch.Tensor]:
    """FeatureNet(arch_mode='unet') — models/msrednet.py:29-130."""
    f = prefix
    c0 = _conv_bn_relu(_conv_bn_relu(img, sd, f + ".conv0.0"), sd, f + ".conv0.1")
    c1 = _conv_bn_relu(c0, sd, f + ".conv1.0", stride=2, pad=2)
    c1 = _conv_bn_relu(_conv_bn_relu(c1, sd, f + ".conv1.1"), sd, f + ".conv1.2")
    c2 = _conv_bn_relu(c1, sd, f + ".conv2.0", stride=2, pad=2)
    c2 = _conv_bn_relu(_conv_bn_relu(c2, sd, f + ".conv2.1"), sd, f + ".conv2.2")

    def up_fuse(skip, x, name):                                    # DeConv2dFuse, models/module.py:506-524
        y = _deconv_bn_relu(x, sd, name + ".deconv.conv.weight", name + ".deconv.bn")
        return _conv_bn_relu(torch.cat((y, skip), 1), sd, name + ".conv")

    out = {"stage1": F.conv2d(c2, sd[f + ".out1.weight"])}
    y = up_fuse(c1, c2, f + ".deconv1")
    out["stage2"] = F.conv2d(y, sd[f + ".out2.weight"])
    y = up_fuse(c0, y, f + ".deconv2")
    out["stage3"] = F.conv2d(y, sd[f + ".out3.weight"])
    return out


def variance_volume(feats: Sequence[torch.Tensor], projs: torch.Tensor, hyps: torch.Tensor) -> torch.Tensor:
    """E[x^2] - E[x]^2 over the reference view and the warped source views — models/msrednet.py:214-231 / 402-420."""
    V = len(feats)
    ref = feats[0]
    D = hyps.shape[1]
    vsum = ref.unsqueeze(2).repeat(1, 1, D, 1, 1)
    vsq = vsum ** 2
    for v in range(1, V):
        w = homography_warp(feats[v], projs[:, v], projs[:, 0], hyps)
        vsum = vsum + w
        vsq = vsq + w ** 2
    return vsq / V - (vsum / V) ** 2


def gn_gru_cell(sd: SD, p: str, x, h):
    """ConvGRUCell2 — models/module.py:54-106 (GroupNorm(1, HC) on both gates and on the candidate)."""
    hc = h.shape[1]
    f = F.conv2d(torch.cat((x, h), 1), sd[p + ".gate_conv.weight"], sd[p + ".gate_conv.bias"], padding=1)
    r, u = torch.split(f, hc, 1)
    r = torch.sigmoid(F.group_norm(r, 1, sd[p + ".reset_gate_norm.weight"], sd[p + ".reset_gate_norm.bias"], 1e-5))
    u = torch.sigmoid(F.group_norm(u, 1, sd[p + ".update_gate_norm.weight"], sd[p + ".update_gate_norm.bias"], 1e-5))
    o = F.conv2d(torch.cat((x, r * h), 1), sd[p + ".output_conv.weight"], sd[p + ".output_conv.bias"], padding=1)
    y = torch.tanh(F.group_norm(o, 1, sd[p + ".output_norm.weight"], sd[p + ".output_norm.bias"], 1e-5))
    return u * h + (1 - u) * y


def red_step(sd: SD, p: str, cost, states: List[torch.Tensor]):
    """One plane of slice_RED_Regularization.forward — models/msrednet.py:355-372.  states = [s1,s2,s3,s4]."""
    def down(x, name):
        return F.relu(F.conv2d(x, sd[f"{p}.{name}.conv.weight"], None, 2, 1))

    def up(x, name):
        return F.relu(F.conv_transpose2d(x, sd[f"{p}.{name}.conv.weight"], None, stride=2, padding=1, output_padding=1))

    x = -cost
    c1 = down(x, "conv1")
    c2 = down(c1, "conv2")
    c3 = down(c2, "conv3")
    s4 = gn_gru_cell(sd, p + ".conv_gru4", c3, states[3])
    u3 = up(s4, "upconv3")
    s3 = gn_gru_cell(sd, p + ".conv_gru3", c2, states[2])
    u2 = up(u3 + s3, "upconv2")
    s2 = gn_gru_cell(sd, p + ".conv_gru2", c1, states[1])
    u1 = up(u2 + s2, "upconv1")
    s1 = gn_gru_cell(sd, p + ".conv_gru1", x, states[0])
    logit = F.conv_transpose2d(u1 + s1, sd[p + ".upconv2d.weight"], sd[p + ".upconv2d.bias"], stride=1, padding=1)
    return logit, [s1, s2, s3, s4]


def red_regulariser(sd: SD, p: str, volume: torch.Tensor) -> torch.Tensor:
    """volume [B,C,D,h,w] -> logits [B,D,h,w] (RED_Regularization.forward, msrednet.py:150-181)."""
    B, C, D, h, w = volume.shape
    states = [volume.new_zeros(B, 8 << l, h >> l, w >> l) for l in range(4)]
    planes = []
    for k in range(D):
        logit, states = red_step(sd, p, volume[:, :, k], states)
        planes.append(logit)
    return torch.cat(planes, 1)


def _stage(sd, i, feats, projs, hyps, stream: bool, cap):
    p = f"cost_regularization.{i}"
    var = variance_volume(feats, projs, hyps)
    logits = red_regulariser(sd, p, var)
    if stream:                                                   # InferDepthNet.forward, msrednet.py:422-436
        e = logits.exp()
        den = e.sum(1) + 1e-10
        depth = (e * hyps).sum(1) / den
        conf = e.max(1)[0] / den
    else:                                                        # DepthNet.forward, msrednet.py:233-240
        prob = F.softmax(logits, 1)
        depth = expected_depth(prob, hyps)
        conf = prob.max(1)[0]
    if cap is not None:
        cap.update(variance=var, logits=logits, hyps=hyps)
    return {"depth": depth, "photometric_confidence": conf}


def _cascade(sd, imgs, proj_matrices, first_range, interval, ndepths, ratios, stream, capture):
    """CascadeREDNet.forward / Infer_CascadeREDNet.forward — models/msrednet.py:279-334 / 473-525."""
    B, V, _, H, W = imgs.shape
    feats = [feature_net(sd, imgs[:, v]) for v in range(V)]
    outputs: dict = {}
    depth = None
    for i in range(3):
        key = f"stage{i + 1}"
        fs = [f[key] for f in feats]
        h, w = fs[0].shape[2:]
        if depth is not None:
            cur = F.interpolate(depth.unsqueeze(1), [H, W], mode="bilinear", align_corners=False).squeeze(1)
        else:
            cur = first_range
        samples = depth_hypotheses(cur, ndepths[i], ratios[i] * interval, [B, H, W])
        hyps = F.interpolate(samples.unsqueeze(1), [ndepths[i], h, w], mode="trilinear", align_corners=False).squeeze(1)
        cap = None if capture is None else capture.setdefault(key, {})
        out = _stage(sd, i, fs, proj_matrices[key], hyps, stream, cap)
        if cap is not None:
            cap["features"] = torch.stack(fs, 1)
        depth = out["depth"]
        outputs[key] = out
        outputs.update(out)
    return outputs


def cascade_rednet_forward(sd: SD, imgs, proj_matrices, depth_values, ndepths=(48, 32, 8), ratios=(4, 2, 1), capture=None):
    """Train/test class: interval = depth_values[0,-1]; ALL columns of depth_values are handed to the first
    stage's get_depth_range_samples (msrednet.py:282-283, 308), which therefore runs from column 0 to the last."""
    interval = float(depth_values[0, -1])
    return _cascade(sd, imgs, proj_matrices, depth_values, interval, ndepths, ratios, False, capture)


def infer_cascade_rednet_forward(sd: SD, imgs, proj_matrices, depth_values, num_depth=192, ndepths=(48, 32, 8),
                                 ratios=(4, 2, 1), capture=None):
    """Predict class: interval = (max - min) / num_depth (msrednet.py:475-477)."""
    interval = (float(depth_values[0, -1]) - float(depth_values[0, 0])) / num_depth
    return _cascade(sd, imgs, proj_matrices, depth_values, interval, ndepths, ratios, True, capture)
