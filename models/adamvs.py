"""Drop-in replacement for the reference's ``models/adamvs.py`` — same importable names, constructor
signatures, ``state_dict`` keys/shapes and forward contract
(``imgs, proj_matrices, depth_values -> {"stage1..3": {...}, "depth", "photometric_confidence", ...}``),
so the reference's ``train_whu.py`` (train / test / profile) and ``predict_whu.py`` run against it unchanged
(reference: models/adamvs.py:316-396 AdaMVSNet, :537-620 Infer_AdaMVSNet, :8-46 cas_mvs_vis_loss; proven by
tests/test_reference_scripts.py, which executes the unmodified scripts).

The modules below only *hold parameters* in the reference's tree; the cascade cost-volume hot path
(warp + cost volume, recurrent regulariser, regression, hypothesis narrowing) runs in the sm_100a
kernels of ``adamvs_b200`` via ``adamvs_b200.cascade`` (inference) and ``adamvs_b200.autograd`` (model.train():
the same ops with their backward kernels).  In eval mode FeatureNet0 and the stage-1 pair U-Net (CostRegNet2D,
SURVEY.md §8f-1) run on the same native convolution kernels with BatchNorm folded; under model.train() they are
torch modules (batch-statistics BatchNorm, cuDNN in true fp32).  There is no CPU fallback: forward() on CPU tensors raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from adamvs_b200 import cascade as _cascade
from adamvs_b200 import ops as _ops

__all__ = ["AdaMVSNet", "Infer_AdaMVSNet", "cas_mvs_vis_loss", "FeatureNet0", "CostRegNet2D", "invalidate_folded"]


# --------------------------------------------------------------------------------------------------
# parameter containers (attribute names are the checkpoint contract)
# --------------------------------------------------------------------------------------------------

class _ConvBN(nn.Module):
    """bias-free conv -> BatchNorm -> ReLU; children are named ``conv`` and ``bn``."""

    def __init__(self, cin, cout, k=3, stride=1, pad=1, transposed=False):
        super().__init__()
        if transposed:
            self.conv = nn.ConvTranspose2d(cin, cout, k, stride=stride, padding=pad, output_padding=stride - 1, bias=False)
        else:
            self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x, x2=None):
        """x2: optional second tensor, concatenated behind x along the channels (read in place by the native conv)."""
        c = self.conv
        if self.training or not _FOLD_BN:
            if x2 is not None:
                x = torch.cat((x, x2), 1)
            return F.relu(self.bn(c(x)), inplace=True)
        w, b, wpk, mode = _folded(c, self.bn)
        if mode is not None and _NATIVE_CONV and x.is_cuda and x.dtype == torch.float32:
            y = _native_conv(x, x2, wpk, b, mode, c, True)
            if y is not None:
                return y
        if x2 is not None:
            x = torch.cat((x, x2), 1)
        if isinstance(c, nn.ConvTranspose2d):
            y = F.conv_transpose2d(x, w, b, c.stride, c.padding, c.output_padding, c.groups, c.dilation)
        else:
            y = F.conv2d(x, w, b, c.stride, c.padding, c.dilation, c.groups)
        return F.relu_(y)


_FOLD_BN = True      # eval-mode BatchNorm is an affine map: fold it into the preceding bias-free conv
_NATIVE_CONV = True  # convolutions of FeatureNet0 / CostRegNet2D on adamvs_b200's kernels instead of cuDNN where supported


def _conv_mode(conv):
    """Which native kernel can run this conv: '3x3' (stride 1/2), 'poly' (5x5 stride 2 in polyphase form), 'deconv'
    (3x3 stride-2 transposed), or None."""
    if conv.groups != 1 or conv.dilation != (1, 1):
        return None
    if isinstance(conv, nn.ConvTranspose2d):
        ok = conv.kernel_size == (3, 3) and conv.stride == (2, 2) and conv.padding == (1, 1) and conv.output_padding == (1, 1)
        return "deconv" if ok else None
    if conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2)):
        return "3x3"
    if conv.kernel_size == (5, 5) and conv.padding == (2, 2) and conv.stride == (2, 2):
        return "poly"
    return None


def _native_conv(x, x2, wpk, b, mode, c, relu, residual=None):
    """Run one folded conv on the C-ABI kernels; None when this channel combination has no kernel.
    `residual` (transposed convs only) is added after the activation."""
    if mode == "deconv":
        if x2 is None and _ops.deconv3x3_supported(x.shape[1], wpk.shape[2]):
            return _ops.deconv3x3(x, wpk, b, relu, residual)
        return None
    if mode == "poly":
        if x2 is None and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and _ops.conv3x3_supported(4 * x.shape[1], 0, wpk.shape[2], 1):
            return _ops.conv3x3(F.pixel_unshuffle(x, 2), None, wpk, b, relu, 1)
        return None
    stride = c.stride[0]
    if x2 is None and x.shape[1] == 3 and wpk.shape[0] == 8 and stride == 1 and x.shape[3] % 4 == 0 and x.is_contiguous() \
            and x.data_ptr() % 16 == 0 and _ops.conv3x3_supported(3, 0, wpk.shape[2], 1):
        return _ops.conv3x3(x, None, wpk, b, relu, 1)                 # the image is read in place: rows 3..7 of wpk are zero
    if x2 is None and x.shape[1] < 8 and wpk.shape[0] == 8:           # other narrow inputs: zero-padded to one 8-channel chunk
        x = F.pad(x, (0, 0, 0, 0, 0, 8 - x.shape[1]))
    if _ops.conv3x3_supported(x.shape[1], 0 if x2 is None else x2.shape[1], wpk.shape[2], stride):
        return _ops.conv3x3(x, x2, wpk, b, relu, stride)
    return None


def invalidate_folded(module: nn.Module) -> None:
    """Drop every cached BatchNorm fold below `module`.  The cache key is (data_ptr, _version) of the tensors involved,
    which in-place tensor methods, load_state_dict, optimiser steps and .to() all change; writes through `.data`
    (`p.data.copy_(...)`) bypass the version counter - call this after such writes."""
    for m in module.modules():
        m.__dict__.pop("_adamvs_folded", None)


def _folded(conv, bn):
    """(weight, bias, packed weight | None, native mode | None) of conv followed by eval-mode BatchNorm (bn may be None:
    plain conv with bias), cached on the conv module and rebuilt whenever a tensor involved is modified or moved
    (see invalidate_folded for the one exception)."""
    src = (conv.weight,) + ((bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else (conv.bias,))
    key = tuple((t.data_ptr(), t._version) for t in src)
    cache = conv.__dict__.get("_adamvs_folded")
    if cache is None or cache[0] != key:
        with torch.no_grad():
            if bn is not None:
                scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
                shape = (1, -1, 1, 1) if isinstance(conv, nn.ConvTranspose2d) else (-1, 1, 1, 1)
                w = (conv.weight * scale.reshape(shape)).contiguous()
                b = (bn.bias - bn.running_mean * scale).contiguous()
            else:
                w, b = conv.weight.detach().contiguous(), conv.bias.detach().contiguous()
            mode = _conv_mode(conv)
            wpk = None
            if mode == "3x3":
                wp = F.pad(w, (0, 0, 0, 0, 0, 8 - w.shape[1])) if w.shape[1] < 8 else w
                wpk = _ops.pack_conv3x3_weight(wp)
            elif mode == "poly":
                wpk = _ops.pack_conv3x3_weight(_ops.polyphase_5x5_s2_weight(w))
            elif mode == "deconv":
                wpk = _ops.pack_deconv3x3_weight(w)
        cache = (key, w, b, wpk, mode)
        conv.__dict__["_adamvs_folded"] = cache
    return cache[1], cache[2], cache[3], cache[4]


class _UpFuse(nn.Module):
    """x2 transposed-conv block, concatenation with the skip tensor, 3x3 fusion conv
    (children ``deconv`` and ``conv``; reference module.py:506-524)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.deconv = _ConvBN(cin, cout, 3, stride=2, pad=1, transposed=True)
        self.conv = _ConvBN(2 * cout, cout, 3, 1, 1)

    def forward(self, skip, x):
        return self.conv(self.deconv(x), skip)


def _context_branch(pool, cin, cout):
    return nn.Sequential(nn.AvgPool2d((pool, pool), stride=(pool, pool)), _ConvBN(cin, cout, 1, 1, 0))


class FeatureNet0(nn.Module):
    """Feature pyramid: [N,3,H,W] -> stage1 [N,32,H/4,W/4], stage2 [N,16,H/2,W/2], stage3 [N,8,H,W]
    (reference models/adamvs.py:49-152)."""

    def __init__(self, base_channels=8, num_stage=3, stride=4):
        super().__init__()
        b = base_channels
        self.base_channels, self.num_stage, self.stride = b, num_stage, stride
        self.conv0 = nn.Sequential(_ConvBN(3, b), _ConvBN(b, b))
        self.conv1 = nn.Sequential(_ConvBN(b, 2 * b, 5, 2, 2), _ConvBN(2 * b, 2 * b), _ConvBN(2 * b, 2 * b))
        self.conv2 = nn.Sequential(_ConvBN(2 * b, 4 * b, 5, 2, 2), _ConvBN(4 * b, 4 * b), _ConvBN(4 * b, 4 * b))
        self.branch1_1 = _context_branch(4, 4 * b, 2 * b)
        self.branch1_2 = _context_branch(8, 4 * b, 2 * b)
        self.out1 = nn.Conv2d(8 * b, 4 * b, 1, bias=False)
        self.deconv1 = _UpFuse(4 * b, 2 * b)
        self.deconv2 = _UpFuse(2 * b, b)
        self.branch2_1 = _context_branch(4, 2 * b, b)
        self.branch2_2 = _context_branch(8, 2 * b, b)
        self.branch3_1 = _context_branch(4, b, b // 2)
        self.branch3_2 = _context_branch(8, b, b // 2)
        self.out2 = nn.Conv2d(4 * b, 2 * b, 1, bias=False)
        self.out3 = nn.Conv2d(2 * b, b, 1, bias=False)
        self.out_channels = [4 * b, 2 * b, b]

    @staticmethod
    def _head(x, br_a, br_b, proj):
        size = x.shape[2:]
        if (_NATIVE_CONV and not proj.training and x.is_cuda and x.dtype == torch.float32 and x.shape[3] % 4 == 0
                and _ops.context_head_supported(x.shape[1], br_a[1].conv.out_channels, proj.out_channels)):
            # both upsamplings, the concatenation and the 1x1 projection in one pass over x (adamvs_context_head_f32);
            # the 8x8 average is taken from the 4x4 averages (one read of x instead of two; a mean of equal-sized means)
            if ((br_a[0].kernel_size, br_b[0].kernel_size) == ((4, 4), (8, 8)) and x.shape[2] % 8 == 0 and x.shape[3] % 8 == 0
                    and _ops.context_pool_supported(x.shape[1], br_a[1].conv.out_channels)):
                # pooling + 1x1 conv + folded BatchNorm + ReLU of both branches in one pass too (adamvs_context_pool_f32)
                wa, ba, _, _ = _folded(br_a[1].conv, br_a[1].bn)
                wc, bc, _, _ = _folded(br_b[1].conv, br_b[1].bn)
                ca, cc = _ops.context_pool(x, wa, ba, wc, bc)
                return _ops.context_head(x, ca, cc, proj.weight)
            pa = br_a[0](x)
            pb = F.avg_pool2d(pa, 2) if (br_a[0].kernel_size, br_b[0].kernel_size) == ((4, 4), (8, 8)) else br_b[0](x)
            return _ops.context_head(x, br_a[1](pa), br_b[1](pb), proj.weight)
        a = F.interpolate(br_a(x), size=size, mode="bilinear", align_corners=False)
        c = F.interpolate(br_b(x), size=size, mode="bilinear", align_corners=False)
        return proj(torch.cat((a, c, x), 1))

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        out = {"stage1": self._head(c2, self.branch1_1, self.branch1_2, self.out1)}
        y = self.deconv1(c1, c2)
        out["stage2"] = self._head(y, self.branch2_1, self.branch2_2, self.out2)
        y = self.deconv2(c0, y)
        out["stage3"] = self._head(y, self.branch3_1, self.branch3_2, self.out3)
        return out


class CostRegNet2D(nn.Module):
    """Stage-1 pairwise U-Net with the depth hypotheses as channels (reference adamvs.py:198-238)."""

    def __init__(self, in_channels, base_channels=8):
        super().__init__()
        n = in_channels
        for i, s in enumerate((1, 2, 1, 2, 1, 2, 1)):
            setattr(self, f"conv{i}", _ConvBN(n, n, 3, s, 1))
        for i in (7, 9, 11):
            setattr(self, f"conv{i}", nn.Sequential(
                nn.ConvTranspose2d(n, n, 3, padding=1, output_padding=1, stride=2, bias=False),
                nn.BatchNorm2d(n), nn.ReLU(inplace=True)))
        self.prob = nn.Conv2d(n, n, 3, stride=1, padding=1)

    def _up(self, seq, x, skip):
        """skip + relu(bn(convT(x))): the skip addition rides in the native kernel's epilogue."""
        if self.training or not _FOLD_BN:
            return skip + seq(x)
        c = seq[0]
        w, b, wpk, mode = _folded(c, seq[1])
        if mode is not None and _NATIVE_CONV and x.is_cuda and x.dtype == torch.float32:
            y = _native_conv(x, None, wpk, b, mode, c, True, residual=skip)
            if y is not None:
                return y
        return skip + F.relu_(F.conv_transpose2d(x, w, b, c.stride, c.padding, c.output_padding, c.groups, c.dilation))

    def forward(self, x):
        e0 = self.conv0(x)
        e2 = self.conv2(self.conv1(e0))
        e4 = self.conv4(self.conv3(e2))
        y = self.conv6(self.conv5(e4))
        y = self._up(self.conv7, y, e4)
        y = self._up(self.conv9, y, e2)
        y = self._up(self.conv11, y, e0)
        if not self.training and _NATIVE_CONV and y.is_cuda and y.dtype == torch.float32:
            w, b, wpk, mode = _folded(self.prob, None)
            out = _native_conv(y, None, wpk, b, mode, self.prob, False) if mode is not None else None
            if out is not None:
                return out
        return self.prob(y)


class _BiasFreeConv(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False)


class _GRUParams(nn.Module):
    def __init__(self, cin, hidden):
        super().__init__()
        self.conv_gates = nn.Sequential(nn.Conv2d(cin + hidden, 2 * hidden, 3, padding=1))
        self.convc = nn.Sequential(nn.Conv2d(cin + hidden, hidden, 3, padding=1))


class _RecurrentRegulariserParams(nn.Module):
    """Weights of the recurrent encoder-decoder (reference adamvs.py:157-170 / 400-413); executed by
    adamvs_b200's K3 kernels, never by torch."""

    def __init__(self, in_channels, up=True, base_channels=8):
        super().__init__()
        b = base_channels
        self.base_channels = b
        self.conv1 = _BiasFreeConv(in_channels, b, 1)
        self.conv_gru1 = _GRUParams(b, b)
        self.conv2 = _BiasFreeConv(b, 2 * b, 2)
        self.conv_gru2 = _GRUParams(2 * b, 2 * b)
        self.upconv1 = nn.ConvTranspose2d(2 * b, b, 3, stride=2, padding=1, output_padding=1)
        if up:
            self.upconv2d = nn.ConvTranspose2d(b, 1, 3, stride=2, padding=1, output_padding=1)
        else:
            self.upconv2d = nn.Conv2d(b, 1, 3, stride=1, padding=1)

    def kernel_weights(self):
        g1, g2 = self.conv_gru1, self.conv_gru2
        return {
            "conv1_w": self.conv1.conv.weight,
            "gates1_w": g1.conv_gates[0].weight, "gates1_b": g1.conv_gates[0].bias,
            "cand1_w": g1.convc[0].weight, "cand1_b": g1.convc[0].bias,
            "conv2_w": self.conv2.conv.weight,
            "gates2_w": g2.conv_gates[0].weight, "gates2_b": g2.conv_gates[0].bias,
            "cand2_w": g2.convc[0].weight, "cand2_b": g2.convc[0].bias,
            "up1_w": self.upconv1.weight, "up1_b": self.upconv1.bias,
            "out_w": self.upconv2d.weight, "out_b": self.upconv2d.bias,
        }


class _StageNet(nn.Module):
    """Children ``reg`` (pair U-Net) and ``reg_fuse`` (recurrent regulariser), as in the reference's
    DepthNet0 / InferDepthNet0 (adamvs.py:241-245, 426-431)."""

    def __init__(self, in_depths, in_channels, in_up=True, base_channels=8):
        super().__init__()
        self.in_up = in_up
        self.reg = CostRegNet2D(in_depths, base_channels)
        self.reg_fuse = _RecurrentRegulariserParams(in_channels, in_up, base_channels)


class _CascadeNet(nn.Module):
    stream_convention = False

    def _build(self, ndepths, depth_intervals_ratio, share_cr, cr_base_chs):
        assert len(ndepths) == len(depth_intervals_ratio)
        self.share_cr = share_cr
        self.ndepths = ndepths
        self.depth_intervals_ratio = depth_intervals_ratio
        self.cr_base_chs = cr_base_chs
        self.num_stage = len(ndepths)
        self.stage_infos = {"stage1": {"scale": 4.0}, "stage2": {"scale": 2.0}, "stage3": {"scale": 1.0}}
        self.feature = FeatureNet0(base_channels=8, stride=4, num_stage=self.num_stage)
        ch = self.feature.out_channels
        # every stage's pair U-Net is sized by ndepths[0] (reference adamvs.py:340, 563-565)
        self.DepthNet = nn.ModuleList([
            _StageNet(self.ndepths[0], ch[0]), _StageNet(self.ndepths[0], ch[1]),
            _StageNet(self.ndepths[0], ch[2], in_up=False)])

    def forward(self, imgs, proj_matrices, depth_values):
        return _cascade.forward(self, imgs, proj_matrices, depth_values)


class AdaMVSNet(_CascadeNet):
    """Train/test class: whole-volume conventions (epsilon in the numerator, softmax, stage-1 view
    weights resized once per stage, interval = depth_values[0,-1])."""
    stream_convention = False

    def __init__(self, ndepths=[48, 32, 8], depth_intervals_ratio=[4, 2, 1], share_cr=False, cr_base_chs=[8, 8, 8]):
        super().__init__()
        self._build(ndepths, depth_intervals_ratio, share_cr, cr_base_chs)


class Infer_AdaMVSNet(_CascadeNet):
    """Predict class: plane-streaming conventions (epsilon in the denominator, un-shifted exp with
    +1e-10, view weights re-resized from the previous stage, interval = (max-min)/num_depth)."""
    stream_convention = True

    def __init__(self, num_depth=384, ndepths=[48, 32, 8], depth_intervals_ratio=[4, 2, 1], share_cr=False,
                 cr_base_chs=[8, 8, 8]):
        super().__init__()
        self.num_depth = num_depth
        self._build(ndepths, depth_intervals_ratio, share_cr, cr_base_chs)


# --------------------------------------------------------------------------------------------------
# loss (training only; pure torch) — reference models/adamvs.py:8-46
# --------------------------------------------------------------------------------------------------

def cas_mvs_vis_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """Sum over stages of smooth-L1(final depth) + mean smooth-L1(pairwise depths), each resized to the
    ground truth's resolution and masked; returns (total, last stage's depth loss)."""
    stage_weights = kwargs.get("dlossw", None)
    total = torch.tensor(0.0, dtype=torch.float32, device=mask_ms["stage1"].device, requires_grad=False)
    depth_loss = None
    for key in [k for k in inputs.keys() if "stage" in k]:
        out = inputs[key]
        gt = depth_gt_ms[key]
        valid = mask_ms[key] > 0.5
        size = [gt.shape[1], gt.shape[2]]

        def masked_l1(est):
            est = F.interpolate(est.unsqueeze(1), size, mode="bilinear", align_corners=False).squeeze(1)
            return F.smooth_l1_loss(est[valid], gt[valid], reduction="mean")

        depth_loss = masked_l1(out["depth"][0:1, :, :])
        pairs = out["pair_result"]
        pair_loss = sum(masked_l1(p) for p in pairs) / len(pairs) if len(pairs) > 0 else 0
        wgt = 1.0 if stage_weights is None else stage_weights[int(key.replace("stage", "")) - 1]
        total += wgt * pair_loss
        total += wgt * depth_loss
    return total, depth_loss
