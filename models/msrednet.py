"""Drop-in replacement for the reference's ``models/msrednet.py`` (MS-REDNet, BASELINE config 5) — same
importable names (``CascadeREDNet``, ``Infer_CascadeREDNet``, ``cas_rednet_loss``), constructor signatures
(including the reference's ``depth_interals_ratio`` spelling), 219 ``state_dict`` keys and forward contract
(reference: models/msrednet.py:246-334, 440-525, 8-27).

The modules below only hold parameters; the variance cost volume, the four-level GroupNorm conv-GRU
regulariser and the regression run in the sm_100a kernels of ``adamvs_b200`` (K5/K6) via
``adamvs_b200.cascade_msred``.  ``FeatureNet`` (outside the path) runs on the native
convolution kernels where its layer shapes have one (``adamvs_conv3x3_f32``), else as true-fp32 cuDNN convolutions.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from adamvs_b200 import cascade_msred as _cascade
from models.adamvs import _ConvBN, _UpFuse

__all__ = ["CascadeREDNet", "Infer_CascadeREDNet", "cas_rednet_loss", "FeatureNet"]


class FeatureNet(nn.Module):
    """U-Net feature pyramid (arch_mode='unet'): [N,3,H,W] -> stage1 [N,32,H/4,W/4], stage2 [N,16,H/2,W/2],
    stage3 [N,8,H,W] (reference models/msrednet.py:29-130; the 'fpn' mode is never constructed by the
    reference's models)."""

    def __init__(self, base_channels, num_stage=3, stride=4, arch_mode="unet"):
        super().__init__()
        assert arch_mode == "unet" and num_stage == 3, "only the configuration the reference constructs is provided"
        b = base_channels
        self.arch_mode, self.stride, self.base_channels, self.num_stage = arch_mode, stride, b, num_stage
        self.conv0 = nn.Sequential(_ConvBN(3, b), _ConvBN(b, b))
        self.conv1 = nn.Sequential(_ConvBN(b, 2 * b, 5, 2, 2), _ConvBN(2 * b, 2 * b), _ConvBN(2 * b, 2 * b))
        self.conv2 = nn.Sequential(_ConvBN(2 * b, 4 * b, 5, 2, 2), _ConvBN(4 * b, 4 * b), _ConvBN(4 * b, 4 * b))
        self.out1 = nn.Conv2d(4 * b, 4 * b, 1, bias=False)
        self.deconv1 = _UpFuse(4 * b, 2 * b)
        self.deconv2 = _UpFuse(2 * b, b)
        self.out2 = nn.Conv2d(2 * b, 2 * b, 1, bias=False)
        self.out3 = nn.Conv2d(b, b, 1, bias=False)
        self.out_channels = [4 * b, 2 * b, b]

    def forward(self, x):
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        out = {"stage1": self.out1(c2)}
        y = self.deconv1(c1, c2)
        out["stage2"] = self.out2(y)
        y = self.deconv2(c0, y)
        out["stage3"] = self.out3(y)
        return out


class _BiasFree(nn.Module):
    def __init__(self, conv):
        super().__init__()
        self.conv = conv


class _GNGRUParams(nn.Module):
    """Parameters of ConvGRUCell2 (reference models/module.py:54-72)."""

    def __init__(self, cin, hidden):
        super().__init__()
        self.output_channel = hidden
        self.gate_conv = nn.Conv2d(cin + hidden, 2 * hidden, 3, padding=1)
        self.reset_gate_norm = nn.GroupNorm(1, hidden, 1e-5, True)
        self.update_gate_norm = nn.GroupNorm(1, hidden, 1e-5, True)
        self.output_conv = nn.Conv2d(cin + hidden, hidden, 3, padding=1)
        self.output_norm = nn.GroupNorm(1, hidden, 1e-5, True)


class _REDRegularisationParams(nn.Module):
    """Weights of RED_Regularization / slice_RED_Regularization (reference msrednet.py:134-148, 339-353); executed
    by adamvs_b200's K6 kernels.  Hidden sizes are the reference's hard-coded 8/16/32/64 state channels."""

    def __init__(self, in_channels, base_channels=8):
        super().__init__()
        b = base_channels
        assert b == 8, "the reference allocates 8/16/32/64-channel states regardless of base_channels"
        self.base_channels = b
        self.conv_gru1 = _GNGRUParams(in_channels, b)
        self.conv_gru2 = _GNGRUParams(2 * b, 2 * b)
        self.conv_gru3 = _GNGRUParams(4 * b, 4 * b)
        self.conv_gru4 = _GNGRUParams(8 * b, 8 * b)
        self.conv1 = _BiasFree(nn.Conv2d(in_channels, 2 * b, 3, stride=2, padding=1, bias=False))
        self.conv2 = _BiasFree(nn.Conv2d(2 * b, 4 * b, 3, stride=2, padding=1, bias=False))
        self.conv3 = _BiasFree(nn.Conv2d(4 * b, 8 * b, 3, stride=2, padding=1, bias=False))
        self.upconv3 = _BiasFree(nn.ConvTranspose2d(8 * b, 4 * b, 3, stride=2, padding=1, output_padding=1, bias=False))
        self.upconv2 = _BiasFree(nn.ConvTranspose2d(4 * b, 2 * b, 3, stride=2, padding=1, output_padding=1, bias=False))
        self.upconv1 = _BiasFree(nn.ConvTranspose2d(2 * b, b, 3, stride=2, padding=1, output_padding=1, bias=False))
        self.upconv2d = nn.ConvTranspose2d(b, 1, kernel_size=3, stride=1, padding=1, output_padding=0)

    def kernel_weights(self):
        g = [self.conv_gru1, self.conv_gru2, self.conv_gru3, self.conv_gru4]
        return {
            "conv1_w": self.conv1.conv.weight, "conv2_w": self.conv2.conv.weight, "conv3_w": self.conv3.conv.weight,
            "gate_w": [c.gate_conv.weight for c in g], "gate_b": [c.gate_conv.bias for c in g],
            "rnorm_w": [c.reset_gate_norm.weight for c in g], "rnorm_b": [c.reset_gate_norm.bias for c in g],
            "unorm_w": [c.update_gate_norm.weight for c in g], "unorm_b": [c.update_gate_norm.bias for c in g],
            "out_w": [c.output_conv.weight for c in g], "out_b": [c.output_conv.bias for c in g],
            "onorm_w": [c.output_norm.weight for c in g], "onorm_b": [c.output_norm.bias for c in g],
            "up3_w": self.upconv3.conv.weight, "up2_w": self.upconv2.conv.weight, "up1_w": self.upconv1.conv.weight,
            "prob_w": self.upconv2d.weight, "prob_b": self.upconv2d.bias,
        }


class _CascadeRED(nn.Module):
    stream_convention = False

    def _build(self, ndepths, depth_interals_ratio, share_cr, cr_base_chs):
        assert len(ndepths) == len(depth_interals_ratio)
        self.share_cr = share_cr
        self.ndepths = ndepths
        self.depth_interals_ratio = depth_interals_ratio
        self.cr_base_chs = cr_base_chs
        self.num_stage = len(ndepths)
        self.stage_infos = {"stage1": {"scale": 4.0}, "stage2": {"scale": 2.0}, "stage3": {"scale": 1.0}}
        self.feature = FeatureNet(base_channels=8, stride=4, num_stage=self.num_stage, arch_mode="unet")
        if share_cr:
            # the reference passes the channel *list* here and cannot construct this configuration
            # (msrednet.py:272, 469); refuse it the same way instead of guessing
            raise TypeError("share_cr=True is not constructible in the reference (in_channels would be a list)")
        self.cost_regularization = nn.ModuleList([
            _REDRegularisationParams(self.feature.out_channels[i], self.cr_base_chs[i]) for i in range(self.num_stage)])

    def forward(self, imgs, proj_matrices, depth_values):
        return _cascade.forward(self, imgs, proj_matrices, depth_values)


class CascadeREDNet(_CascadeRED):
    """Train/test class: softmax over the whole logit volume; interval = depth_values[0,-1]; stage-1
    hypotheses run from depth_values[:,0] to depth_values[:,-1] (the reference passes all three columns,
    msrednet.py:308 — preserved, not fixed)."""
    stream_convention = False

    def __init__(self, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], share_cr=False, cr_base_chs=[8, 8, 8]):
        super().__init__()
        self._build(ndepths, depth_interals_ratio, share_cr, cr_base_chs)


class Infer_CascadeREDNet(_CascadeRED):
    """Predict class: plane streaming with un-shifted exp and +1e-10; interval = (max-min)/num_depth."""
    stream_convention = True

    def __init__(self, num_depth=384, ndepths=[48, 32, 8], depth_interals_ratio=[4, 2, 1], share_cr=False,
                 cr_base_chs=[8, 8, 8]):
        super().__init__()
        self.num_depth = num_depth
        self._build(ndepths, depth_interals_ratio, share_cr, cr_base_chs)


def cas_rednet_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """Sum over stages of the masked smooth-L1 depth loss (reference models/msrednet.py:8-27); returns
    (total, last stage's loss)."""
    stage_weights = kwargs.get("dlossw", None)
    total = torch.tensor(0.0, dtype=torch.float32, device=mask_ms["stage1"].device, requires_grad=False)
    depth_loss = None
    for key in [k for k in inputs.keys() if "stage" in k]:
        valid = mask_ms[key] > 0.5
        depth_loss = F.smooth_l1_loss(inputs[key]["depth"][valid], depth_gt_ms[key][valid], reduction="mean")
        wgt = 1.0 if stage_weights is None else stage_weights[int(key.replace("stage", "")) - 1]
        total += wgt * depth_loss
    return total, depth_loss
