"""Training path of the cascade hot path: ``torch.autograd.Function``s over the C-ABI forward and
backward kernels, and the train-mode forward of ``AdaMVSNet`` built from them (reference models/adamvs.py:247-312,
342-396 under ``model.train()``; the loop that calls it is train_whu.py:265-300).

What is native (csrc/train.cu, csrc/costvolume.cu): K1 / K2 forward and backward (bilinear gather / scatter-add, view
weights), every convolution of the recurrent regulariser - forward, data gradient and weight gradient -, the softmax
expectation with its backward.  What is torch: the pointwise GRU algebra between the convolutions and the back-propagation
*through time* over the depth planes (autograd walks the per-plane Functions in reverse), FeatureNet0 and the pair U-Net,
whose train-mode BatchNorm needs batch statistics (they are outside the hot path), and the
hypothesis / weight resampling glue.  Gradients follow the reference exactly: none through the sampling grid
(models/module.py:538), but through the hypotheses' VALUES into the previous stage's depth map (not detached,
adamvs.py:365) and through the view weights into the pair U-Net.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import ops
from .ops import _check, _f32c, _guard, _p, _stream, lib


# --------------------------------------------------------------------------------------------------------------------
# 3x3 convolution (stride 1 | 2) and stride-2 transposed convolution with run-time channel counts
# --------------------------------------------------------------------------------------------------------------------

def conv2d_raw(x, w, bias, stride: int, transposed: bool, relu: bool):
    x, w = _f32c(x, "x"), _f32c(w, "w")
    N, Cin, h, wd = x.shape
    Cout = w.shape[1] if transposed else w.shape[0]
    assert (w.shape[0] if transposed else w.shape[1]) == Cin and tuple(w.shape[2:]) == (3, 3), (x.shape, w.shape, transposed)
    ho, wo = (2 * h, 2 * wd) if transposed else (h // stride, wd // stride)
    y = torch.empty((N, Cout, ho, wo), device=x.device, dtype=torch.float32)
    b = None if bias is None else _f32c(bias, "bias")
    with _guard(x):
        _check(lib().adamvs_conv2d_f32(_p(x), _p(w), _p(b), _p(y), N, Cin, Cout, h, wd, int(stride), int(transposed), int(relu),
                                       _stream()), "conv2d")
    return y


def conv2d_wgrad_raw(x, gy, stride: int):
    """-> [Cout,Cin,3,3] with Cin = x's channels, Cout = gy's channels."""
    x, gy = _f32c(x, "x"), _f32c(gy, "gy")
    N, Cin, h, wd = x.shape
    Cout = gy.shape[1]
    assert tuple(gy.shape) == (N, Cout, h // stride, wd // stride), (x.shape, gy.shape, stride)
    gw = torch.zeros((Cout, Cin, 3, 3), device=x.device, dtype=torch.float32)
    with _guard(x):
        _check(lib().adamvs_conv2d_wgrad_f32(_p(x), _p(gy), _p(gw), N, Cin, Cout, h, wd, int(stride), _stream()), "conv2d_wgrad")
    return gw


class Conv3x3Fn(torch.autograd.Function):
    """y = act(conv(x, w) + b): Conv2d [Cout,Cin,3,3] stride 1|2 padding 1, or (transposed) ConvTranspose2d [Cin,Cout,3,3]
    stride 2 padding 1 output_padding 1."""

    @staticmethod
    def forward(ctx, x, w, b, stride, transposed, relu):
        y = conv2d_raw(x, w, b, stride, transposed, relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (int(stride), bool(transposed), bool(relu), b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, transposed, relu, has_b = ctx.cfg
        gy = gy.contiguous()
        if relu:
            gy = gy * (y > 0)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if transposed:                       # conv stride 2 with the same tensor read as [out = Cin, in = Cout]
                gx = conv2d_raw(gy, w, None, 2, False, False)
            elif stride == 2:                    # transposed conv with the same tensor read as [Cin_t = Cout, Cout_t = Cin]
                gx = conv2d_raw(gy, w, None, 2, True, False)
            else:
                gx = conv2d_raw(gy, w.flip(2, 3).transpose(0, 1).contiguous(), None, 1, False, False)
        if ctx.needs_input_grad[1]:
            gw = conv2d_wgrad_raw(gy, x, 2) if transposed else conv2d_wgrad_raw(x, gy, stride)
        if has_b and ctx.needs_input_grad[2]:
            gb = gy.sum((0, 2, 3))
        return gx, gw, gb, None, None, None


def conv3x3(x, w, b=None, stride=1, transposed=False, relu=False):
    return Conv3x3Fn.apply(x, w, b, stride, transposed, relu)


# --------------------------------------------------------------------------------------------------------------------
# K1 / K2
# --------------------------------------------------------------------------------------------------------------------

class PairScoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, relproj, hyp_mode, hyp_src, half_range, D):
        hyp = ops.Hyp(hyp_mode, hyp_src, half_range)
        ctx.save_for_backward(feat, relproj, hyp.src, half_range)
        ctx.cfg = (hyp_mode, hyp.ncol, int(D))
        return ops.pair_score(feat, relproj, hyp, D)

    @staticmethod
    def backward(ctx, g):
        feat, relproj, hyp_src, half = ctx.saved_tensors
        mode, ncol, D = ctx.cfg
        feat = _f32c(feat, "feat")
        B, V, C, h, w = feat.shape
        g_feat = torch.zeros_like(feat)
        with _guard(feat):
            _check(lib().adamvs_pair_score_bwd_f32(_p(feat), _p(_f32c(relproj, "relproj")), mode, _p(hyp_src), ncol, _p(half),
                                                   _p(_f32c(g, "g_score")), _p(g_feat), B, V, C, D, h, w, _stream()), "pair_score_bwd")
        return g_feat, None, None, None, None, None


class FusedVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weights, relproj, hyp_mode, hyp_src, half_range, eps_mode, D):
        hyp = ops.Hyp(hyp_mode, hyp_src, half_range)
        ctx.save_for_backward(feat, weights, relproj, hyp.src, half_range)
        ctx.cfg = (hyp_mode, hyp.ncol, int(eps_mode), int(D))
        return ops.fused_volume(feat, relproj, hyp, weights, eps_mode, D)

    @staticmethod
    def backward(ctx, g):
        feat, weights, relproj, hyp_src, half = ctx.saved_tensors
        mode, ncol, eps_mode, D = ctx.cfg
        feat, weights = _f32c(feat, "feat"), _f32c(weights, "weights")
        B, V, C, h, w = feat.shape
        g_feat, g_w = torch.zeros_like(feat), torch.zeros_like(weights)
        with _guard(feat):
            _check(lib().adamvs_fused_volume_bwd_f32(_p(feat), _p(_f32c(relproj, "relproj")), mode, _p(hyp_src), ncol, _p(half),
                                                     _p(weights), eps_mode, _p(_f32c(g, "g_volume")), _p(g_feat), _p(g_w),
                                                     B, V, C, D, h, w, _stream()), "fused_volume_bwd")
        return g_feat, g_w, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------------------------------
# K4
# --------------------------------------------------------------------------------------------------------------------

class SoftmaxExpectFn(torch.autograd.Function):
    """logits, hyp [N,D,h,w] -> depth = sum_k softmax(logits)_k * hyp_k, conf = max_k softmax(logits)_k   [N,h,w]."""

    @staticmethod
    def forward(ctx, logits, hyp):
        logits, hyp = _f32c(logits, "logits"), _f32c(hyp, "hyp")
        N, D, h, w = logits.shape
        assert tuple(hyp.shape) == (N, D, h, w), (logits.shape, hyp.shape)
        depth = torch.empty((N, h, w), device=logits.device, dtype=torch.float32)
        conf = torch.empty_like(depth)
        with _guard(logits):
            _check(lib().adamvs_softmax_expect_f32(_p(logits), _p(hyp), _p(depth), _p(conf), N, D, h, w, _stream()), "softmax_expect")
        ctx.save_for_backward(logits, hyp, depth)
        return depth, conf

    @staticmethod
    def backward(ctx, g_depth, g_conf):
        logits, hyp, depth = ctx.saved_tensors
        N, D, h, w = logits.shape
        g_logits = torch.empty_like(logits)
        g_hyp = torch.empty_like(hyp) if ctx.needs_input_grad[1] else None
        gd = None if g_depth is None else _f32c(g_depth, "g_depth")
        gc = None if g_conf is None else _f32c(g_conf, "g_conf")
        with _guard(logits):
            _check(lib().adamvs_softmax_expect_bwd_f32(_p(logits), _p(hyp), _p(depth), _p(gd), _p(gc), _p(g_logits), _p(g_hyp),
                                                       N, D, h, w, _stream()), "softmax_expect_bwd")
        return g_logits, g_hyp


# --------------------------------------------------------------------------------------------------------------------
# recurrent regulariser, training form (reference CostRegNetRED.forward, adamvs.py:172-195; ConvGRUCell, module.py:24-52)
# --------------------------------------------------------------------------------------------------------------------

def _gru(x, h, wg, bg, wc, bc):
    g = conv3x3(torch.cat((x, h), 1), wg, bg)
    r, u = torch.chunk(g, 2, dim=1)
    r, u = torch.sigmoid(r), torch.sigmoid(u)
    c = torch.tanh(conv3x3(torch.cat((x, r * h), 1), wc, bc))
    return u * h + (1 - u) * c


def regulariser_train(volume: torch.Tensor, p: dict, out_up: bool) -> torch.Tensor:
    """volume [B,C,D,h,w] -> logits [B,D,Ho,Wo]; `p` = _RecurrentRegulariserParams.kernel_weights()."""
    B, C, D, h, w = volume.shape
    h1 = volume.new_zeros((B, 8, h, w))
    h2 = volume.new_zeros((B, 16, h // 2, w // 2))
    logits: List[torch.Tensor] = []
    for k in range(D):
        x1 = conv3x3(volume[:, :, k].contiguous(), p["conv1_w"], None, relu=True)
        h1 = _gru(x1, h1, p["gates1_w"], p["gates1_b"], p["cand1_w"], p["cand1_b"])
        x2 = conv3x3(h1, p["conv2_w"], None, stride=2, relu=True)
        h2 = _gru(x2, h2, p["gates2_w"], p["gates2_b"], p["cand2_w"], p["cand2_b"])
        y = torch.relu(conv3x3(h2, p["up1_w"], p["up1_b"], stride=2, transposed=True) + h1)
        if out_up:
            logits.append(conv3x3(y, p["out_w"], p["out_b"], stride=2, transposed=True))
        else:
            logits.append(conv3x3(y, p["out_w"], p["out_b"]))
    return torch.cat(logits, 1)


# --------------------------------------------------------------------------------------------------------------------
# AdaMVSNet.forward under model.train()
# --------------------------------------------------------------------------------------------------------------------

_STAGES = ("stage1", "stage2", "stage3")


def _hypotheses(cur_depth, D: int, interval_pixel: float, h: int, w: int):
    """get_depth_range_samples (module.py:646-663): cur_depth [B,2] = (min, max) planes, or [B,h,w] -> [B,D,h,w]."""
    k = torch.arange(D, dtype=cur_depth.dtype, device=cur_depth.device)
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (D - 1)
        planes = lo[:, None] + k[None, :] * step[:, None]
        return planes[:, :, None, None].repeat(1, 1, h, w)
    lo = cur_depth - D / 2 * interval_pixel
    hi = cur_depth + D / 2 * interval_pixel
    step = (hi - lo) / (D - 1)
    return lo.unsqueeze(1) + k.reshape(1, -1, 1, 1) * step.unsqueeze(1)


def forward_train(net, imgs, proj_matrices: Dict[str, torch.Tensor], depth_values):
    if net.stream_convention:
        raise NotImplementedError("Infer_AdaMVSNet is the predict class (predict_whu.py): train AdaMVSNet")
    from .cascade import _true_fp32
    ndepths = [int(d) for d in net.ndepths]
    ratios = [float(r) for r in net.depth_intervals_ratio]
    B, V = imgs.shape[:2]
    Vs = V - 1
    imgs = imgs.float()
    depth_values = depth_values.float().contiguous()
    depth_interval = float(depth_values[0, -1])                     # adamvs.py:346 (a host read, as in the reference)
    depth_range = depth_values[:, 0:-1]
    with _true_fp32():
        per_view = [net.feature(imgs[:, v]) for v in range(V)]       # train-mode BatchNorm: statistics per call, like the reference
    feats = {k: torch.stack([f[k] for f in per_view], 1).contiguous() for k in _STAGES}
    with torch.no_grad():
        relproj, half = ops.cascade_prepare([proj_matrices[k] for k in _STAGES], depth_values, ops.INTERVAL_LAST_COLUMN, 0,
                                            ndepths, ratios)
    outputs: dict = {}
    depth = None
    stage1_conf = None                                              # list of [B,1,h1,w1], carries gradient into the pair U-Net
    for i, key in enumerate(_STAGES):
        feat = feats[key]
        _, _, C, h, w = feat.shape
        D = ndepths[i]
        if depth is None:
            plane_range = torch.stack((depth_range[:, 0], depth_range[:, -1]), 1).contiguous()
            hyp_args = (ops.HYP_PLANES, plane_range, None)
            hyp_vals = _hypotheses(depth_range, D, ratios[i] * depth_interval, h, w)
        else:
            hyp_args = (ops.HYP_PER_PIXEL, depth.detach().contiguous(), half[i:i + 1])   # positions carry no gradient
            hyp_vals = _hypotheses(depth, D, ratios[i] * depth_interval, h, w)            # values do (adamvs.py:365)
        pair_depths: List[torch.Tensor] = []
        if stage1_conf is None:
            score = PairScoreFn.apply(feat, relproj[i], hyp_args[0], hyp_args[1], hyp_args[2], D)      # [B,Vs,D,h,w]
            stage1_conf = []
            for v in range(Vs):                                      # one U-Net call per view: BatchNorm statistics per call
                with _true_fp32():
                    pl = net.DepthNet[i].reg(score[:, v].contiguous())
                pd, pc = SoftmaxExpectFn.apply(pl, hyp_vals)
                pair_depths.append(pd)
                stage1_conf.append(pc.unsqueeze(1))
            weights = torch.cat(stage1_conf, 1)
        else:
            weights = torch.cat([F.interpolate(c, [h, w], mode="bilinear", align_corners=False) for c in stage1_conf], 1)
        volume = FusedVolumeFn.apply(feat, weights.contiguous(), relproj[i], hyp_args[0], hyp_args[1], hyp_args[2],
                                     ops.EPS_NUMERATOR, D)
        out_up = i < 2
        logits = regulariser_train(volume, net.DepthNet[i].reg_fuse.kernel_weights(), out_up)
        if out_up:
            hyp_out = F.interpolate(hyp_vals, [2 * h, 2 * w], mode="bilinear", align_corners=False)   # module.py:622
        else:
            hyp_out = hyp_vals
        depth, conf = SoftmaxExpectFn.apply(logits, hyp_out.contiguous())
        out = {"depth": depth, "photometric_confidence": conf, "pair_confidence": list(stage1_conf), "pair_result": pair_depths}
        outputs[key] = out
        outputs.update(out)
    return outputs
