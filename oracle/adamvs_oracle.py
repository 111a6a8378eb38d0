"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (torch fp32) restatement of the Ada-MVS cascade cost-volume hot path, written as pure
functions over a flat ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
(``adamvs_b200``, ``models``) never does.

Parity status: PINNED.  The reference ships no golden vectors or tests, so the pin is
``tests/golden/*.npz`` — outputs of the reference's own ``AdaMVSNet`` / ``Infer_AdaMVSNet`` executed
in the build container from /root/reference by ``tests/golden/make_golden.py`` (committed), on the
seeded weights and inputs of ``adamvs_b200.synth``.  ``tests/test_oracle_golden.py`` checks every
function chain below against those files.

The arithmetic lives in PyTorch (ATen), which is the reference's own third-party dependency
(README.md:15-18 "pytorch >= 1.3.1", unpinned; this image: torch 2.11.0).  Each function names the
reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# small building blocks
# ------------------------------------------------------------------------------------------------

def _bn_eval(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """Eval-mode BatchNorm2d (eps 1e-5, running statistics)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=1e-5)


def _conv_bn_relu(x, sd, p, stride=1, pad=1):
    """models/module.py:164-199 (Conv2d wrapper) and :254-261 (ConvBnReLU): bias-free conv, BN, ReLU."""
    return F.relu(_bn_eval(F.conv2d(x, sd[p + ".conv.weight"], None, stride, pad), sd, p + ".bn"))


def _deconv_bn_relu(x, sd, wkey, bnprefix):
    """k=3, s=2, p=1, output_padding=1 transposed conv, BN, ReLU (module.py:202-245, adamvs.py:212-225)."""
    y = F.conv_transpose2d(x, sd[wkey], None, stride=2, padding=1, output_padding=1)
    return F.relu(_bn_eval(y, sd, bnprefix))


def _resize(x, size):
    """Bilinear, align_corners=False (what F.upsample(mode='bilinear') and the explicit
    F.interpolate calls in adamvs.py:116,296,505,522 and module.py:622 resolve to)."""
    return F.interpolate(x, size=list(size), mode="bilinear", align_corners=False)


# ------------------------------------------------------------------------------------------------
# FeatureNet0 — models/adamvs.py:49-152
# ------------------------------------------------------------------------------------------------

def feature_net(sd: SD, img: torch.Tensor, prefix: str = "feature") -> Dict[str, torch.Tensor]:
    f = prefix
    c0 = _conv_bn_relu(_conv_bn_relu(img, sd, f + ".conv0.0"), sd, f + ".conv0.1")
    c1 = _conv_bn_relu(c0, sd, f + ".conv1.0", stride=2, pad=2)
    c1 = _conv_bn_relu(_conv_bn_relu(c1, sd, f + ".conv1.1"), sd, f + ".conv1.2")
    c2 = _conv_bn_relu(c1, sd, f + ".conv2.0", stride=2, pad=2)
    c2 = _conv_bn_relu(_conv_bn_relu(c2, sd, f + ".conv2.1"), sd, f + ".conv2.2")

    def head(x, level: int, out_key: str):
        # two pooled-context branches (4x4 and 8x8 average pooling -> 1x1 conv -> resize back),
        # concatenated in front of x, then a bias-free 1x1 projection (adamvs.py:115-124 etc.)
        size = x.shape[2:]
        b1 = _resize(_conv_bn_relu(F.avg_pool2d(x, 4, 4), sd, f"{f}.branch{level}_1.1", pad=0), size)
        b2 = _resize(_conv_bn_relu(F.avg_pool2d(x, 8, 8), sd, f"{f}.branch{level}_2.1", pad=0), size)
        return F.conv2d(torch.cat((b1, b2, x), 1), sd[f"{f}.{out_key}.weight"])

    def up_fuse(skip, x, name):
        # DeConv2dFuse, module.py:506-524
        y = _deconv_bn_relu(x, sd, name + ".deconv.conv.weight", name + ".deconv.bn")
        return _conv_bn_relu(torch.cat((y, skip), 1), sd, name + ".conv")

    out = {"stage1": head(c2, 1, "out1")}
    x = up_fuse(c1, c2, f + ".deconv1")
    out["stage2"] = head(x, 2, "out2")
    x = up_fuse(c0, x, f + ".deconv2")
    out["stage3"] = head(x, 3, "out3")
    return out


# ------------------------------------------------------------------------------------------------
# homography warp — models/module.py:527-568
# ------------------------------------------------------------------------------------------------

# Test knob: evaluate the relative projection in fp64 (rounded once to fp32) instead of the reference's fp32
# torch.inverse + matmul.  Mathematically the same function; the outputs move by the reference's own fp32
# arithmetic noise, which tests use to state how closely ANY independent implementation can agree with it.
RELPROJ_FP64 = False


def homography_warp(src_fea, src_proj, ref_proj, depth_values):
    """src_fea [B,C,h,w]; projections [B,4,4]; depth_values [B,D] or [B,D,h,w] -> [B,C,D,h,w].
    Same op order as the reference (relative projection by inverse+matmul, rotate the pixel grid,
    scale by depth, translate, perspective divide, normalise, grid_sample with zero padding and
    align_corners=True) so that CPU results agree with it to the last bits."""
    B, C, h, w = src_fea.shape
    D = depth_values.shape[1]
    with torch.no_grad():
        if RELPROJ_FP64:
            rel = torch.matmul(src_proj.double(), torch.linalg.inv(ref_proj.double())).to(src_proj.dtype)
        else:
            rel = torch.matmul(src_proj, torch.inverse(ref_proj))
        R, t = rel[:, :3, :3], rel[:, :3, 3:4]
        ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=src_fea.device),
                                torch.arange(w, dtype=torch.float32, device=src_fea.device), indexing="ij")
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=src_fea.device)))
        rays = torch.matmul(R, pix.unsqueeze(0).expand(B, 3, h * w))                # [B,3,hw]
        pts = rays.unsqueeze(2) * depth_values.reshape(B, 1, D, -1) + t.reshape(B, 3, 1, 1)
        uv = pts[:, :2] / pts[:, 2:3]
        gx = uv[:, 0] / ((w - 1) / 2) - 1
        gy = uv[:, 1] / ((h - 1) / 2) - 1
        grid = torch.stack((gx, gy), dim=3).reshape(B, D * h, w, 2)
    out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return out.reshape(B, C, D, h, w)


# ------------------------------------------------------------------------------------------------
# cascade hypotheses — models/module.py:628-663
# ------------------------------------------------------------------------------------------------

def depth_hypotheses(cur_depth, ndepth: int, interval_pixel: float, shape: Sequence[int]):
    """cur_depth [B,2+] (first stage: min in column 0, max in the last column) or [B,h,w]
    (later stages: previous depth map) -> [B,ndepth,h,w]."""
    k = torch.arange(ndepth, dtype=cur_depth.dtype, device=cur_depth.device)
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        planes = lo[:, None] + k[None, :] * step[:, None]
        return planes[:, :, None, None].repeat(1, 1, shape[1], shape[2])
    lo = cur_depth - ndepth / 2 * interval_pixel
    hi = cur_depth + ndepth / 2 * interval_pixel
    step = (hi - lo) / (ndepth - 1)
    return lo.unsqueeze(1) + k.reshape(1, -1, 1, 1) * step.unsqueeze(1)


# ------------------------------------------------------------------------------------------------
# regression — models/module.py:617-625
# ------------------------------------------------------------------------------------------------

def expected_depth(prob, hyps):
    if hyps.dim() <= 2:
        hyps = hyps.reshape(*hyps.shape, 1, 1)
    else:
        hyps = _resize(hyps, prob.shape[2:])
    return torch.sum(prob * hyps, 1)


# ------------------------------------------------------------------------------------------------
# CostRegNet2D (stage-1 pair U-Net, depth as channels) — models/adamvs.py:198-238
# ------------------------------------------------------------------------------------------------

def pair_unet(sd: SD, p: str, x):
    e0 = _conv_bn_relu(x, sd, p + ".conv0")
    e2 = _conv_bn_relu(_conv_bn_relu(e0, sd, p + ".conv1", stride=2), sd, p + ".conv2")
    e4 = _conv_bn_relu(_conv_bn_relu(e2, sd, p + ".conv3", stride=2), sd, p + ".conv4")
    y = _conv_bn_relu(_conv_bn_relu(e4, sd, p + ".conv5", stride=2), sd, p + ".conv6")
    y = e4 + _deconv_bn_relu(y, sd, p + ".conv7.0.weight", p + ".conv7.1")
    y = e2 + _deconv_bn_relu(y, sd, p + ".conv9.0.weight", p + ".conv9.1")
    y = e0 + _deconv_bn_relu(y, sd, p + ".conv11.0.weight", p + ".conv11.1")
    return F.conv2d(y, sd[p + ".prob.weight"], sd[p + ".prob.bias"], 1, 1)


# ------------------------------------------------------------------------------------------------
# recurrent regulariser — models/module.py:5-52 (ConvGRUCell), models/adamvs.py:172-195 / 415-424
# ------------------------------------------------------------------------------------------------

def gru_cell(sd: SD, p: str, x, h):
    """gates = conv(cat(x,h)); first half reset, second half update; candidate from cat(x, r*h);
    new state u*h + (1-u)*c."""
    g = F.conv2d(torch.cat((x, h), 1), sd[p + ".conv_gates.0.weight"], sd[p + ".conv_gates.0.bias"], 1, 1)
    r, u = torch.chunk(g, 2, dim=1)
    r, u = torch.sigmoid(r), torch.sigmoid(u)
    c = torch.tanh(F.conv2d(torch.cat((x, r * h), 1), sd[p + ".convc.0.weight"], sd[p + ".convc.0.bias"], 1, 1))
    return u * h + (1 - u) * c


def red_step(sd: SD, p: str, plane, h1, h2, upsample_out: bool):
    """One depth plane through the recurrent encoder-decoder. plane [B,C,h,w]; h1 [B,8,h,w];
    h2 [B,16,h/2,w/2]. Returns (logit [B,1,Ho,Wo], h1', h2')."""
    x1 = F.relu(F.conv2d(plane, sd[p + ".conv1.conv.weight"], None, 1, 1))
    h1 = gru_cell(sd, p + ".conv_gru1", x1, h1)
    x2 = F.relu(F.conv2d(h1, sd[p + ".conv2.conv.weight"], None, 2, 1))
    h2 = gru_cell(sd, p + ".conv_gru2", x2, h2)
    up = F.conv_transpose2d(h2, sd[p + ".upconv1.weight"], sd[p + ".upconv1.bias"], 2, 1, 1)
    y = F.relu(up + h1)
    if upsample_out:
        logit = F.conv_transpose2d(y, sd[p + ".upconv2d.weight"], sd[p + ".upconv2d.bias"], 2, 1, 1)
    else:
        logit = F.conv2d(y, sd[p + ".upconv2d.weight"], sd[p + ".upconv2d.bias"], 1, 1)
    return logit, h1, h2


def red_regulariser(sd: SD, p: str, volume, upsample_out: bool):
    """volume [B,C,D,h,w] -> logits [B,D,Ho,Wo]; states start at zero with 8 / 16 channels
    (adamvs.py:175-176)."""
    B, C, D, h, w = volume.shape
    h1 = volume.new_zeros((B, 8, h, w))
    h2 = volume.new_zeros((B, 16, h // 2, w // 2))
    logits = []
    for d in range(D):
        l, h1, h2 = red_step(sd, p, volume[:, :, d], h1, h2, upsample_out)
        logits.append(l)
    return torch.cat(logits, 1)


# ------------------------------------------------------------------------------------------------
# stage bodies
# ------------------------------------------------------------------------------------------------

def _pair_branch(sd, i, ref, srcs, ref_proj, src_projs, hyps, cap):
    """Stage-1 view weighting, shared by both classes (adamvs.py:268-283 / 464-490):
    per source view mean_c(ref*warp) -> pair U-Net -> softmax_D -> (max, expectation)."""
    weights, pair_depths, products = [], [], []
    for v, (s, sp) in enumerate(zip(srcs, src_projs)):
        prod = ref.unsqueeze(2) * homography_warp(s, sp, ref_proj, hyps)          # [B,C,D,h,w]
        score = prod.mean(dim=1)
        if cap is not None:
            cap.setdefault("pair_score", []).append(score)
        p = F.softmax(pair_unet(sd, f"DepthNet.{i}.reg", score), dim=1)
        weights.append(p.max(1)[0].unsqueeze(1))
        pair_depths.append(expected_depth(p, hyps))
        products.append(prod)
    return weights, pair_depths, products


def stage_whole_volume(sd: SD, i: int, feats: List[torch.Tensor], projs: torch.Tensor, hyps, ndepth: int,
                       conf_maps: Optional[List[torch.Tensor]], upsample_out: bool, cap: Optional[dict] = None):
    """DepthNet0.forward — models/adamvs.py:247-312 (train/test class).
    fused = (1e-5 + sum_v w_v * ref * warp_v) / sum_v w_v ; softmax ; expectation ; max."""
    assert hyps.shape[1] == ndepth
    ref, srcs = feats[0], feats[1:]
    pm = torch.unbind(projs, 1)
    assert len(pm) == len(feats)
    h, w = ref.shape[2:]
    acc = 1e-5
    wsum = 0
    pair_depths: List[torch.Tensor] = []
    if conf_maps is None:
        conf_maps, pair_depths, products = _pair_branch(sd, i, ref, srcs, pm[0], pm[1:], hyps, cap)
        for wv, prod in zip(conf_maps, products):
            wsum = wsum + wv.unsqueeze(1)
            acc = acc + prod * wv.unsqueeze(1)
        pair_conf = conf_maps
    else:
        for s, sp, cm in zip(srcs, pm[1:], conf_maps):
            prod = ref.unsqueeze(2) * homography_warp(s, sp, pm[0], hyps)
            wv = _resize(cm, (h, w))
            wsum = wsum + wv.unsqueeze(1)
            acc = acc + prod * wv.unsqueeze(1)
        pair_conf = conf_maps            # untouched stage-1 maps are handed on (adamvs.py:302)
    fused = acc / wsum
    logits = red_regulariser(sd, f"DepthNet.{i}.reg_fuse", fused, upsample_out)
    prob = F.softmax(logits, dim=1)
    if cap is not None:
        cap["fused"] = fused
        cap["logits"] = logits
    return {"depth": expected_depth(prob, hyps), "photometric_confidence": prob.max(1)[0],
            "pair_confidence": pair_conf, "pair_result": pair_depths}


def stage_plane_stream(sd: SD, i: int, feats: List[torch.Tensor], projs: torch.Tensor, hyps, ndepth: int,
                       conf_maps: Optional[List[torch.Tensor]], upsample_out: bool, cap: Optional[dict] = None):
    """InferDepthNet0.forward — models/adamvs.py:433-533 (predict class).
    plane = sum_v w_v*ref*warp_v / (1e-5 + sum_v w_v); e = exp(logit) without max shift;
    depth = sum d*e/(sum e + 1e-10); confidence = max e/(sum e + 1e-10).  The returned
    pair_confidence list carries this stage's *resized* weights (first four entries), which is what
    makes the next stage interpolate twice."""
    assert hyps.shape[1] == ndepth
    ref, srcs = feats[0], feats[1:]
    pm = torch.unbind(projs, 1)
    assert len(pm) == len(feats)
    B, C, h, w = ref.shape
    Ho, Wo = (2 * h, 2 * w) if upsample_out else (h, w)
    pair_conf: List[torch.Tensor] = []
    pair_depths: List[torch.Tensor] = []
    if conf_maps is None:
        conf_maps, pair_depths, _ = _pair_branch(sd, i, ref, srcs, pm[0], pm[1:], hyps, cap)
        pair_conf = conf_maps            # same list object: the plane loop below appends to it
    # weights at this stage's resolution (identity resize at stage 1)
    wmaps = [_resize(conf_maps[v], (h, w)) for v in range(len(srcs))]
    h1 = ref.new_zeros((B, 8, h, w))
    h2 = ref.new_zeros((B, 16, h // 2, w // 2))
    esum = ref.new_zeros((B, 1, Ho, Wo))
    dsum = ref.new_zeros((B, 1, Ho, Wo))
    emax = ref.new_zeros((B, 1, Ho, Wo))
    fused_planes, logit_planes = [], []
    for d in range(ndepth):
        hyp = hyps[:, d:d + 1]
        num = 0
        den = 1e-5
        for v, (s, sp) in enumerate(zip(srcs, pm[1:])):
            prod = homography_warp(s, sp, pm[0], hyp) * ref.unsqueeze(2)
            pair_conf.append(wmaps[v])
            num = num + prod * wmaps[v].unsqueeze(1)
            den = den + wmaps[v].unsqueeze(1)
        plane = (num / den).squeeze(2)
        logit, h1, h2 = red_step(sd, f"DepthNet.{i}.reg_fuse", plane, h1, h2, upsample_out)
        e = logit.exp()
        take = (emax < e).float()
        emax = take * e + (1 - take) * emax
        if upsample_out:
            hyp = _resize(hyp, (Ho, Wo))
        dsum = hyp * e + dsum
        esum = esum + e
        if cap is not None:
            fused_planes.append(plane)
            logit_planes.append(logit)
    if cap is not None:
        cap["fused"] = torch.stack(fused_planes, 2)
        cap["logits"] = torch.cat(logit_planes, 1)
    den = esum + 1e-10
    return {"depth": (dsum / den).squeeze(1), "photometric_confidence": (emax / den).squeeze(1),
            "pair_confidence": pair_conf, "pair_result": pair_depths}


# ------------------------------------------------------------------------------------------------
# top-level forwards
# ------------------------------------------------------------------------------------------------

def _cascade(sd, imgs, proj_matrices, first_range, interval, ndepths, ratios, stage_fn, capture):
    feats = [feature_net(sd, imgs[:, v]) for v in range(imgs.shape[1])]
    B, _, _, H, W = imgs.shape
    outputs: dict = {}
    depth = None
    conf = None
    for i, (nd, ratio) in enumerate(zip(ndepths, ratios)):
        key = f"stage{i + 1}"
        scale = (4, 2, 1)[i]
        if depth is None:
            cur, shape = first_range, [B, H // scale, W // scale]
        else:
            cur, shape = depth, [B, depth.shape[1], depth.shape[2]]
        hyps = depth_hypotheses(cur, nd, ratio * interval, shape)
        cap = None
        if capture is not None:
            cap = capture.setdefault(key, {})
            cap["hyps"] = hyps
            cap["features"] = [f[key] for f in feats]
        out = stage_fn(sd, i, [f[key] for f in feats], proj_matrices[key], hyps, nd, conf, i < 2, cap)
        depth, conf = out["depth"], out["pair_confidence"]
        outputs[key] = out
        outputs.update(out)
    return outputs


@torch.no_grad()
def adamvs_forward(sd: SD, imgs, proj_matrices, depth_values, ndepths=(48, 32, 8), ratios=(4, 2, 1),
                   capture: Optional[dict] = None):
    """AdaMVSNet.forward — models/adamvs.py:342-396. depth_values [B,3] = [min, max, interval];
    the interval scalar is taken from batch item 0."""
    interval = float(depth_values[0, -1])
    return _cascade(sd, imgs, proj_matrices, depth_values[:, 0:-1], interval, ndepths, ratios,
                    stage_whole_volume, capture)


@torch.no_grad()
def infer_adamvs_forward(sd: SD, imgs, proj_matrices, depth_values, num_depth=192, ndepths=(48, 32, 8),
                         ratios=(4, 2, 1), capture: Optional[dict] = None):
    """Infer_AdaMVSNet.forward — models/adamvs.py:567-620. depth_values [B,2] = [min, max];
    interval = (max - min) / num_depth from batch item 0."""
    lo, hi = float(depth_values[0, 0]), float(depth_values[0, -1])
    interval = (hi - lo) / num_depth
    return _cascade(sd, imgs, proj_matrices, depth_values, interval, ndepths, ratios,
                    stage_plane_stream, capture)
