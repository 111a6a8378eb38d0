"""GPU parity tests proper: every C-ABI kernel and the whole drop-in forward against the oracle
(oracle/adamvs_oracle.py, CPU fp32) on the same seeded inputs, and against the committed reference
outputs in tests/golden.  Tolerances are the north_star's: depth within 1e-4 relative, probability
within 1e-4 absolute per pixel (fp32 path); intermediates within 2e-4 * max|.|."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import adamvs_oracle as O
from tests.helpers import abs_err, load_golden, make_case, rebuild_case, rel_err

pytestmark = pytest.mark.gpu

DEPTH_RTOL = 1e-4
PROB_ATOL = 1e-4


def _dev():
    return torch.device("cuda:0")


def _ops():
    from adamvs_b200 import ops
    return ops


def _model(cls, sd, ndepths, num_depth=192):
    from models.adamvs import AdaMVSNet, Infer_AdaMVSNet
    if cls == "whole":
        m = AdaMVSNet(ndepths=list(ndepths), depth_intervals_ratio=[4.0, 2.0, 1.0])
    else:
        m = Infer_AdaMVSNet(num_depth=num_depth, ndepths=list(ndepths), depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    return m.to(_dev()).eval()


def _to_dev(proj):
    return {k: v.to(_dev()) for k, v in proj.items()}


# ------------------------------------------------------------------------------------------------
# kernels one by one
# ------------------------------------------------------------------------------------------------

def test_library_loaded_is_in_tree():
    ops = _ops()
    assert ops.lib().adamvs_abi_version() == 1
    assert ops.LIB_PATH.endswith("adamvs_b200/libadamvs_b200.so")


def test_cascade_prepare_matches_float64_inverse():
    ops = _ops()
    _, proj, dv2 = __import__("adamvs_b200.synth", fromlist=["x"]).make_sample(2, 64, 96, 5, seed=9)
    relproj, half = ops.cascade_prepare([proj[k].to(_dev()) for k in ("stage1", "stage2", "stage3")],
                                        dv2.to(_dev()), ops.INTERVAL_FROM_RANGE, 192, [48, 32, 8], [4.0, 2.0, 1.0])
    for s, k in enumerate(("stage1", "stage2", "stage3")):
        p = proj[k].double()
        rel = p[:, 1:] @ torch.linalg.inv(p[:, :1])
        want = torch.cat([rel[:, :, :3, :3].reshape(2, 4, 9), rel[:, :, :3, 3]], -1).float()
        assert torch.equal(relproj[s].cpu(), want) or abs_err(relproj[s].cpu(), want) <= 1e-6 * float(want.abs().max())
    interval = (float(dv2[0, 1]) - float(dv2[0, 0])) / 192
    want_half = torch.tensor([48 / 2 * (4.0 * interval), 32 / 2 * (2.0 * interval), 8 / 2 * (1.0 * interval)],
                             dtype=torch.float64).float()
    assert torch.equal(half.cpu(), want_half)


def test_resize_matches_interpolate():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 4, 24, 36, generator=g)
    for ho, wo in ((48, 72), (96, 144), (24, 36), (31, 50)):
        got = ops.resize_bilinear(x.to(_dev()), ho, wo).cpu()
        want = F.interpolate(x, size=[ho, wo], mode="bilinear", align_corners=False)
        assert abs_err(got, want) < 2e-7


@pytest.mark.parametrize("C,D,h,w", [(32, 8, 16, 24), (16, 5, 32, 48), (8, 3, 30, 50)])
def test_pair_score_and_fused_volume_vs_oracle(C, D, h, w):
    """K1 and K2 (both epsilon conventions, plane and per-pixel hypotheses) against the oracle's
    warp -> product -> weighted mean."""
    ops = _ops()
    from adamvs_b200 import synth
    B, V = 2, 5
    g = torch.Generator().manual_seed(C + D)
    feat = torch.randn(B, V, C, h, w, generator=g)
    cams = synth.make_cameras(h, w, V - 1)["stage3"]                      # full-res cameras of an h x w image
    proj = torch.stack([cams, synth.make_cameras(h, w, V - 1, jitter_seed=4)["stage3"]])
    dv = torch.tensor([[520.0, 680.0], [540.0, 660.0]])
    wts = torch.rand(B, V - 1, h, w, generator=g) * 0.9 + 0.05
    relproj, half = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 192,
                                        [D, D, D], [4.0, 2.0, 1.0])
    cur = 600 + 20 * torch.randn(B, h, w, generator=g)
    interval = (float(dv[0, 1]) - float(dv[0, 0])) / 192
    for mode in ("planes", "pixel"):
        if mode == "planes":
            hyp = ops.Hyp(ops.HYP_PLANES, dv.to(_dev()))
            hyps = O.depth_hypotheses(dv, D, 0.0, [B, h, w])
        else:
            hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur.to(_dev()), half[1:2])
            hyps = O.depth_hypotheses(cur, D, 2.0 * interval, [B, h, w])
        prods = [feat[:, 0].unsqueeze(2) * O.homography_warp(feat[:, v], proj[:, v], proj[:, 0], hyps)
                 for v in range(1, V)]
        score_want = torch.stack([p.mean(1) for p in prods], 1)
        score = ops.pair_score(feat.to(_dev()), relproj[0], hyp, D).cpu()
        assert abs_err(score, score_want) < 2e-4 * float(score_want.abs().max()), mode
        num = sum(p * wts[:, v].unsqueeze(1).unsqueeze(1) for v, p in enumerate(prods))
        wsum = wts.sum(1).unsqueeze(1).unsqueeze(1)
        for eps_mode, want in ((ops.EPS_NUMERATOR, (1e-5 + num) / wsum), (ops.EPS_DENOMINATOR, num / (1e-5 + wsum))):
            got = ops.fused_volume(feat.to(_dev()), relproj[0], hyp, wts.to(_dev()), eps_mode, D).cpu()
            assert abs_err(got, want) < 2e-4 * float(want.abs().max()), (mode, eps_mode)


@pytest.mark.parametrize("case", ["zoom", "rot90", "w_not_mult4"])
def test_cost_volume_fallback_paths_vs_oracle(case):
    """Geometries whose source footprint does not fit the shared-memory box (x2.2 zoom, 90 degree roll)
    take the in-kernel global-gather path; widths that are not a multiple of 4 take the non-TMA kernels.
    All must agree with the oracle like the staged path does."""
    ops = _ops()
    from adamvs_b200 import synth
    B, V, C, D = 1, 5, 8, 6
    h, w = (40, 64) if case != "w_not_mult4" else (30, 50)
    g = torch.Generator().manual_seed(17)
    feat = torch.randn(B, V, C, h, w, generator=g)
    proj = synth.make_cameras(h, w, V - 1)["stage3"].unsqueeze(0).clone()
    if case == "zoom":
        proj[0, 1:, :2, :] *= 2.2                                      # source focal length and principal point x2.2
    elif case == "rot90":
        rz = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        K = torch.tensor([[1.2 * w, 0, w / 2.0], [0, 1.2 * w, h / 2.0], [0, 0, 1.0]])
        for v in (1, 3):
            proj[0, v, :3, :] = K @ rz @ torch.linalg.inv(K) @ proj[0, v, :3, :]
    dv = torch.tensor([[520.0, 680.0]])
    wts = torch.rand(B, V - 1, h, w, generator=g) * 0.9 + 0.05
    relproj, _ = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 192, [D] * 3, [1.0] * 3)
    hyp = ops.Hyp(ops.HYP_PLANES, dv.to(_dev()))
    hyps = O.depth_hypotheses(dv, D, 0.0, [B, h, w])
    prods = [feat[:, 0].unsqueeze(2) * O.homography_warp(feat[:, v], proj[:, v], proj[:, 0], hyps) for v in range(1, V)]
    score_want = torch.stack([p.mean(1) for p in prods], 1)
    score = ops.pair_score(feat.to(_dev()), relproj[0], hyp, D).cpu()
    assert abs_err(score, score_want) < 2e-4 * float(score_want.abs().max())
    num = sum(p * wts[:, v].unsqueeze(1).unsqueeze(1) for v, p in enumerate(prods))
    want = num / (1e-5 + wts.sum(1).unsqueeze(1).unsqueeze(1))
    got = ops.fused_volume(feat.to(_dev()), relproj[0], hyp, wts.to(_dev()), ops.EPS_DENOMINATOR, D).cpu()
    assert float(want.abs().max()) > 0.1
    assert abs_err(got, want) < 2e-4 * float(want.abs().max())


def test_fused_volume_zero_padding_and_behind_camera():
    """Out-of-image taps contribute zero per corner (grid_sample zeros padding); a source camera that looks the other
    way (finite Z < 0 everywhere) is sampled at the mirrored coordinates X/Z, Y/Z exactly as the reference does
    (models/module.py:553-556 has no sign test) - compared against the oracle, not against zeros."""
    ops = _ops()
    B, V, C, D, h, w = 1, 2, 8, 2, 8, 12
    feat = torch.ones(B, V, C, h, w)
    proj = torch.eye(4).repeat(B, V, 1, 1)
    proj[0, 1, 0, 3] = 5.5 * 600.0          # shifts u by +5.5 px at depth 600 -> right part leaves the image
    dv = torch.tensor([[600.0, 600.0001]])
    relproj, _ = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 1, [D] * 3, [1.0] * 3)
    hyp = ops.Hyp(ops.HYP_PLANES, dv.to(_dev()))
    hyps = O.depth_hypotheses(dv, D, 0.0, [B, h, w])
    want = O.homography_warp(feat[:, 1], proj[:, 1], proj[:, 0], hyps)
    got = ops.fused_volume(feat.to(_dev()), relproj[0], hyp, torch.ones(B, 1, h, w, device=_dev()),
                           ops.EPS_DENOMINATOR, D).cpu() * (1 + 1e-5)
    assert abs_err(got, want) < 1e-5
    assert abs(float(want[0, 0, 0, 0, w - 6]) - 0.5) < 1e-4 and float(want[0, 0, 0, 0, w - 5]) == 0.0
    proj[0, 1, 2, 2] = -1.0                  # source camera looks the other way: Z = -600 everywhere
    proj[0, 1, 0, 3] = -9.0 * 600.0          # u = (x*d - 5400) / -d = 9 - x: columns 0..9 land inside the 12-wide image, mirrored
    proj[0, 1, 1, 1] = -1.0                  # v = (-y*d) / -d = y
    ramp = torch.arange(w, dtype=torch.float32).repeat(B, V, C, h, 1) + 1.0
    relproj, _ = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 1, [D] * 3, [1.0] * 3)
    want = O.homography_warp(ramp[:, 1], proj[:, 1], proj[:, 0], hyps) * ramp[:, 0].unsqueeze(2)
    got = ops.fused_volume(ramp.to(_dev()), relproj[0], hyp, torch.ones(B, 1, h, w, device=_dev()), ops.EPS_DENOMINATOR, D).cpu() * (1 + 1e-5)
    assert float(want.abs().max()) > 20.0 and float(want[0, 0, 0, 0, 11]) == 0.0       # mirrored ramp inside, zero where u < 0
    assert abs_err(got, want) < 2e-4 * float(want.abs().max())


@pytest.mark.parametrize("C,D,h,w,up,prob", [(32, 6, 16, 24, True, "softmax"), (16, 4, 32, 48, True, "exp"),
                                             (8, 3, 64, 96, False, "softmax"), (8, 2, 34, 46, False, "exp"),
                                             # ragged tiles of the TMA-fed tail (w % 8 == 0, neither a multiple of 32 nor of the tile height)
                                             (16, 3, 38, 72, True, "softmax"), (8, 2, 50, 104, False, "exp")])
def test_regnet_red_vs_oracle(C, D, h, w, up, prob, math=None):
    """K3 with the regression folded in: logits, depth and confidence against the oracle's plane loop."""
    ops = _ops()
    from adamvs_b200 import synth
    B = 2
    g = torch.Generator().manual_seed(h + D)
    i = {32: 0, 16: 1, 8: 2}[C]
    assert up == (i < 2)
    sd = synth.fill_state_dict(synth.state_dict_shapes(8), 21)
    p = f"DepthNet.{i}.reg_fuse"
    sd[p + ".upconv2d.weight"] = sd[p + ".upconv2d.weight"] * 20      # sharpen the logits
    vol = torch.randn(B, C, D, h, w, generator=g)
    cur = 600 + 10 * torch.randn(B, h, w, generator=g)
    half_range = torch.tensor([3.3], dtype=torch.float32)
    hyps = O.depth_hypotheses(cur, D, float(half_range) / (D / 2), [B, h, w])
    logits_want = O.red_regulariser(sd, p, vol, up)
    Ho, Wo = logits_want.shape[2:]
    hy = O._resize(hyps, (Ho, Wo)) if up else hyps
    if prob == "softmax":
        pr = F.softmax(logits_want, 1)
        depth_want, conf_want = (pr * hy).sum(1), pr.max(1)[0]
        mode = ops.PROB_SOFTMAX
    else:
        e = logits_want.exp()
        den = e.sum(1) + 1e-10
        depth_want, conf_want = (e * hy).sum(1) / den, e.max(1)[0] / den
        mode = ops.PROB_EXP_EPS
    names = {"conv1_w": ".conv1.conv.weight", "gates1_w": ".conv_gru1.conv_gates.0.weight",
             "gates1_b": ".conv_gru1.conv_gates.0.bias", "cand1_w": ".conv_gru1.convc.0.weight",
             "cand1_b": ".conv_gru1.convc.0.bias", "conv2_w": ".conv2.conv.weight",
             "gates2_w": ".conv_gru2.conv_gates.0.weight", "gates2_b": ".conv_gru2.conv_gates.0.bias",
             "cand2_w": ".conv_gru2.convc.0.weight", "cand2_b": ".conv_gru2.convc.0.bias",
             "up1_w": ".upconv1.weight", "up1_b": ".upconv1.bias", "out_w": ".upconv2d.weight", "out_b": ".upconv2d.bias"}
    wd = {k: sd[p + v].to(_dev()) for k, v in names.items()}
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur.to(_dev()), half_range.to(_dev()))
    kw = {} if math is None else {"math": {"tc": ops.MATH_TC_FP32, "ffma": ops.MATH_FFMA}[math]}
    depth, conf, logits = ops.regnet_red(vol.to(_dev()), wd, hyp, up, mode, want_logits=True, **kw)
    assert abs_err(logits.cpu(), logits_want) < 1e-4 * max(1.0, float(logits_want.abs().max()))
    assert rel_err(depth.cpu(), depth_want) < DEPTH_RTOL
    assert abs_err(conf.cpu(), conf_want) < PROB_ATOL


def test_softmax_regress_vs_torch():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    N, D, h, w = 8, 48, 12, 20
    logits = 6 * torch.randn(N, D, h, w, generator=g)
    dv = torch.tensor([[520.0, 680.0], [500.0, 700.0]])
    hyps = O.depth_hypotheses(dv, D, 0.0, [2, h, w]).repeat_interleave(4, 0)
    pr = F.softmax(logits, 1)
    depth, conf = ops.softmax_regress(logits.to(_dev()), ops.Hyp(ops.HYP_PLANES, dv.to(_dev())), ops.PROB_SOFTMAX, 4)
    assert rel_err(depth.cpu(), (pr * hyps).sum(1)) < 2e-6
    assert abs_err(conf.cpu(), pr.max(1)[0]) < 2e-6


# ------------------------------------------------------------------------------------------------
# whole forward
# ------------------------------------------------------------------------------------------------

def _compare_outputs(out, want_depth, want_conf, tag):
    d_err = rel_err(out["depth"].cpu(), want_depth)
    p_err = abs_err(out["photometric_confidence"].cpu(), want_conf)
    assert d_err < DEPTH_RTOL, f"{tag}: depth rel err {d_err:.3e}"
    assert p_err < PROB_ATOL, f"{tag}: prob abs err {p_err:.3e}"


@pytest.mark.parametrize("name", ["small_d8", "batch2_d8", "full_d48"])
@pytest.mark.parametrize("cls", ["whole", "stream"])
def test_forward_matches_reference_golden(name, cls):
    """The CUDA path against the reference's own outputs (fixtures generated from /root/reference)."""
    g = load_golden(name)
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    m = _model(cls, sd, meta["ndepths"], meta["num_depth"])
    dv = dv3 if cls == "whole" else dv2
    out = m(imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
    for s in ("stage1", "stage2", "stage3"):
        _compare_outputs(out[s], g[f"{cls}_{s}_depth"], g[f"{cls}_{s}_conf"], f"{name}/{cls}/{s}")
        assert abs_err(torch.stack(out[s]["pair_confidence"][:4], 1).cpu(), g[f"{cls}_{s}_pair_conf4"]) < PROB_ATOL
        assert len(out[s]["pair_confidence"]) == int(g[f"{cls}_{s}_pair_conf_len"])
        if s == "stage1":
            assert rel_err(torch.stack(out[s]["pair_result"], 1).cpu(), g[f"{cls}_{s}_pair_result"]) < DEPTH_RTOL
        else:
            assert out[s]["pair_result"] == []
    assert out["depth"] is out["stage3"]["depth"]
    for leaf in (out["depth"], out["photometric_confidence"], *out["pair_confidence"][:4]):
        assert isinstance(leaf, torch.Tensor)


@pytest.mark.parametrize("name,cls", [("full_d48", "stream"), ("batch2_d8", "whole")])
def test_forward_is_bit_reproducible(name, cls):
    """Six forwards of the same inputs give bit-identical outputs.  The kernels of the recurrent sweep overlap through
    programmatic dependent launch; a load of the GRU state through the non-coherent L1 path (stale line from the previous
    plane) once made planes >= 1 differ from run to run while staying inside the parity tolerance most of the time."""
    g = load_golden(name)
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    m = _model(cls, sd, meta["ndepths"], meta["num_depth"])
    dv = dv3 if cls == "whole" else dv2
    args = (imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
    outs = []
    for _ in range(6):
        out = m(*args)
        outs.append({s: (out[s]["depth"].clone(), out[s]["photometric_confidence"].clone(),
                         torch.stack(out["stage1"]["pair_result"], 1).clone(), torch.stack(out[s]["pair_confidence"][:4], 1).clone())
                     for s in ("stage1", "stage2", "stage3")})
    for o in outs[1:]:
        for s in ("stage1", "stage2", "stage3"):
            # the pair maps too: cuDNN's default (atomics) transposed convolutions in the torch-fallback layers of the
            # 8-plane pair U-Net once differed from run to run in them (cascade._true_fp32 now asks for deterministic ones)
            assert all(torch.equal(a, b) for a, b in zip(o[s], outs[0][s])), s


def test_intermediates_match_reference_golden():
    from adamvs_b200 import cascade
    g = load_golden("small_d8")
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    m = _model("whole", sd, meta["ndepths"])
    cap = {}
    cascade.forward(m, imgs.to(_dev()), _to_dev(proj), dv3.to(_dev()), capture=cap)
    for i, s in enumerate(("stage1", "stage2", "stage3")):
        want = g[f"features_{s}"]
        assert abs_err(cap[s]["features"].cpu(), want) < 2e-5 * float(np.abs(want).max())
        want = g[f"whole_s{i + 1}_fused"]
        assert abs_err(cap[s]["fused"].cpu(), want) < 2e-4 * float(np.abs(want).max()), s
        want = g[f"whole_s{i + 1}_logits"]
        assert abs_err(cap[s]["logits"].cpu(), want) < 2e-4 * max(1.0, float(np.abs(want).max())), s
    want = g["whole_s1_pair_score"]
    assert abs_err(cap["stage1"]["pair_score"].cpu(), want) < 2e-4 * float(np.abs(want).max())


@pytest.mark.parametrize("cls", ["whole", "stream"])
def test_forward_full_size_vs_oracle(cls):
    """BASELINE config 2: 1 ref + 4 src views, 768x384, ndepths 48/32/8, calibrated seeded weights."""
    sd, imgs, proj, dv2, dv3 = make_case(1, 384, 768, (48, 32, 8), 192, 60.0, 31, 17, O.feature_net)
    if cls == "whole":
        want = O.adamvs_forward(sd, imgs, proj, dv3)
    else:
        want = O.infer_adamvs_forward(sd, imgs, proj, dv2, num_depth=192)
    m = _model(cls, sd, (48, 32, 8), 192)
    out = m(imgs.to(_dev()), _to_dev(proj), (dv3 if cls == "whole" else dv2).to(_dev()))
    assert tuple(out["stage1"]["depth"].shape) == (1, 192, 384)
    assert tuple(out["depth"].shape) == (1, 384, 768)
    for s in ("stage1", "stage2", "stage3"):
        _compare_outputs(out[s], want[s]["depth"], want[s]["photometric_confidence"], f"full/{cls}/{s}")
    conf = want["stage1"]["photometric_confidence"]                  # not the vacuous 1/D regime:
    assert float(conf.max()) > 5.0 / 48 and float(conf.std()) > 2e-3
    assert float(want["stage3"]["photometric_confidence"].max()) > 0.5


@pytest.mark.parametrize("V", [2, 3, 4, 7])
@pytest.mark.parametrize("cls", ["whole", "stream"])
def test_forward_other_view_counts_vs_oracle(V, cls):
    """`--view_num` is a run-time choice in the reference's scripts (predict_whu.py:33): 1, 2, 3 and 6 source views
    (the kernels are instantiated for 1..6) through the whole forward against the oracle, both classes."""
    from adamvs_b200 import synth
    nd = (8, 4, 2)
    imgs, proj, dv2 = synth.make_sample(1, 64, 96, V, seed=50 + V)
    dv3 = torch.cat([dv2, torch.full((1, 1), (synth.DEPTH_MAX - synth.DEPTH_MIN) / 32)], 1)
    sd = synth.fill_state_dict(synth.state_dict_shapes(nd[0]), 60 + V)
    f = O.feature_net(sd, imgs[:, 0])
    sd = synth.calibrate_state_dict(sd, {k: float(f[k].std()) for k in f}, 20.0)
    if cls == "whole":
        want = O.adamvs_forward(sd, imgs, proj, dv3, ndepths=nd)
    else:
        want = O.infer_adamvs_forward(sd, imgs, proj, dv2, num_depth=32, ndepths=nd)
    m = _model(cls, sd, nd, 32)
    out = m(imgs.to(_dev()), _to_dev(proj), (dv3 if cls == "whole" else dv2).to(_dev()))
    for s in ("stage1", "stage2", "stage3"):
        _compare_outputs(out[s], want[s]["depth"], want[s]["photometric_confidence"], f"V={V}/{cls}/{s}")
        assert len(out[s]["pair_confidence"]) >= V - 1
    assert len(out["stage1"]["pair_result"]) == V - 1


def test_forward_tf32_math_within_its_own_stated_tolerance():
    """The reduced-precision variant (bench.py --math tf32: K3's activations rounded to tf32 in one pass, weights still
    split exactly) is reported separately with ITS OWN tolerance: depth 5e-3 relative, probability 3e-2
    absolute - it does NOT meet the fp32 bar, which is why it is not the default."""
    ops = _ops()
    sd, imgs, proj, dv2, dv3 = make_case(1, 384, 768, (48, 32, 8), 192, 60.0, 31, 17, O.feature_net)
    want = O.infer_adamvs_forward(sd, imgs, proj, dv2, num_depth=192)
    m = _model("stream", sd, (48, 32, 8), 192)
    ops.set_default_math(ops.MATH_TC_TF32)
    try:
        out = m(imgs.to(_dev()), _to_dev(proj), dv2.to(_dev()))
    finally:
        ops.set_default_math(None)
    worst_d = max(rel_err(out[s]["depth"].cpu(), want[s]["depth"]) for s in ("stage1", "stage2", "stage3"))
    worst_p = max(abs_err(out[s]["photometric_confidence"].cpu(), want[s]["photometric_confidence"]) for s in ("stage1", "stage2", "stage3"))
    print(f"tf32 math: worst depth rel err {worst_d:.3e}, worst prob abs err {worst_p:.3e}")
    assert worst_d < 5e-3 and worst_p < 3e-2
    assert worst_p > PROB_ATOL                                       # and it really is another arithmetic than the default


def test_dataparallel_wrapper_and_module_prefix():
    """predict_whu.py wraps the model in nn.DataParallel and loads 'module.'-prefixed keys."""
    g = load_golden("small_d8")
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    from models.adamvs import Infer_AdaMVSNet
    m = torch.nn.DataParallel(Infer_AdaMVSNet(num_depth=meta["num_depth"], ndepths=list(meta["ndepths"]),
                                              depth_intervals_ratio=[4.0, 2.0, 1.0]), device_ids=[0])
    m.cuda()
    m.load_state_dict({"module." + k: v for k, v in sd.items()})
    m.eval()
    with torch.no_grad():
        out = m(imgs.cuda(), {k: v.cuda() for k, v in proj.items()}, dv2.cuda())
    _compare_outputs(out, g["stream_stage3_depth"], g["stream_stage3_conf"], "dp")


def test_cpu_tensors_are_refused():
    from adamvs_b200 import ops
    g = load_golden("small_d8")
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    m = _model("whole", sd, meta["ndepths"])
    with pytest.raises(ops.AdamvsError):
        m(imgs, proj, dv3)


# ------------------------------------------------------------------------------------------------
# MS-REDNet (BASELINE config 5)
# ------------------------------------------------------------------------------------------------

def _msred_model(cls, sd, ndepths, num_depth):
    from models.msrednet import CascadeREDNet, Infer_CascadeREDNet
    if cls == "whole":
        m = CascadeREDNet(ndepths=list(ndepths), depth_interals_ratio=[4.0, 2.0, 1.0])
    else:
        m = Infer_CascadeREDNet(num_depth=num_depth, ndepths=list(ndepths), depth_interals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    return m.to(_dev()).eval()


@pytest.mark.parametrize("C,D,h,w", [(32, 5, 16, 24), (16, 4, 32, 64), (8, 6, 40, 56)])
def test_variance_volume_vs_oracle(C, D, h, w):
    """K5 (staged and global-gather paths: w = 24/56 are not TMA-eligible... 24 % 4 == 0 is) against the oracle."""
    ops = _ops()
    from adamvs_b200 import synth
    from oracle import msrednet_oracle as MO
    B, V = 2, 5
    g = torch.Generator().manual_seed(C + D)
    feat = torch.randn(B, V, C, h, w, generator=g)
    proj = torch.stack([synth.make_cameras(h, w, V - 1)["stage3"], synth.make_cameras(h, w, V - 1, jitter_seed=4)["stage3"]])
    dv = torch.tensor([[520.0, 680.0], [540.0, 660.0]])
    relproj, half = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 192, [D, D, D], [4.0, 2.0, 1.0])
    cur = 600 + 20 * torch.randn(B, h, w, generator=g)
    interval = (float(dv[0, 1]) - float(dv[0, 0])) / 192
    for mode in ("planes", "pixel"):
        if mode == "planes":
            hyp, hyps = ops.Hyp(ops.HYP_PLANES, dv.to(_dev())), O.depth_hypotheses(dv, D, 0.0, [B, h, w])
        else:
            hyp, hyps = ops.Hyp(ops.HYP_PER_PIXEL, cur.to(_dev()), half[1:2]), O.depth_hypotheses(cur, D, 2.0 * interval, [B, h, w])
        want = MO.variance_volume([feat[:, v] for v in range(V)], proj, hyps)
        got = ops.variance_volume(feat.to(_dev()), relproj[0], hyp, D).cpu()
        assert abs_err(got, want) < 2e-4 * float(want.abs().max()), mode


@pytest.mark.parametrize("C,D,h,w,prob", [(32, 4, 16, 24, "exp"), (16, 3, 32, 64, "softmax"), (8, 3, 64, 96, "exp")])
def test_regnet_msred_vs_oracle(C, D, h, w, prob):
    """K6: logits, depth and confidence against the oracle's plane loop (TMA and generic-tile layer paths)."""
    ops = _ops()
    from adamvs_b200 import synth
    from oracle import msrednet_oracle as MO
    B = 2
    g = torch.Generator().manual_seed(h + D)
    i = {32: 0, 16: 1, 8: 2}[C]
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 23)
    p = f"cost_regularization.{i}"
    sd[p + ".upconv2d.weight"] = sd[p + ".upconv2d.weight"] * 4
    vol = torch.rand(B, C, D, h, w, generator=g) * 2.0
    cur = 600 + 10 * torch.randn(B, h, w, generator=g)
    half_range = torch.tensor([3.3], dtype=torch.float32)
    hyps = O.depth_hypotheses(cur, D, float(half_range) / (D / 2), [B, h, w])
    logits_want = MO.red_regulariser(sd, p, vol)
    if prob == "softmax":
        pr = F.softmax(logits_want, 1)
        depth_want, conf_want, mode = (pr * hyps).sum(1), pr.max(1)[0], ops.PROB_SOFTMAX
    else:
        e = logits_want.exp()
        den = e.sum(1) + 1e-10
        depth_want, conf_want, mode = (e * hyps).sum(1) / den, e.max(1)[0] / den, ops.PROB_EXP_EPS
    import models.msrednet as M
    holder = M._REDRegularisationParams(C)
    holder.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")})
    holder = holder.to(_dev())
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur.to(_dev()), half_range.to(_dev()))
    depth, conf, logits = ops.regnet_msred(vol.to(_dev()), holder.kernel_weights(), hyp, mode, want_logits=True)
    assert abs_err(logits.cpu(), logits_want) < 1e-4 * max(1.0, float(logits_want.abs().max()))
    assert rel_err(depth.cpu(), depth_want) < DEPTH_RTOL
    assert abs_err(conf.cpu(), conf_want) < PROB_ATOL


def _arithmetic_noise(run):
    """max |prob| and relative depth deviation per stage between the oracle as the reference computes (fp32
    torch.inverse for the relative projection) and the same function with that 4x4 product evaluated in fp64:
    the reference's own fp32 arithmetic noise on these inputs."""
    a = run()
    O.RELPROJ_FP64 = True
    try:
        b = run()
    finally:
        O.RELPROJ_FP64 = False
    noise = {s: (rel_err(b[s]["depth"], a[s]["depth"]), abs_err(b[s]["photometric_confidence"], a[s]["photometric_confidence"]))
             for s in ("stage1", "stage2", "stage3")}
    return a, noise


def _compare_msred(out, want, noise, tag):
    """MS-REDNet's probabilities are far more sensitive to the sample positions of the warp than Ada-MVS's (the
    variance is quadratic in the features and GroupNorm rescales every plane): the reference moves by `noise`
    when a single 4x4 product is evaluated more accurately.  Bar: the north_star's 1e-4 / 1e-4 plus twice that
    noise, capped at 2e-3 absolute probability."""
    for s in ("stage1", "stage2", "stage3"):
        d_err = rel_err(out[s]["depth"].cpu(), want[s]["depth"])
        p_err = abs_err(out[s]["photometric_confidence"].cpu(), want[s]["photometric_confidence"])
        d_tol, p_tol = DEPTH_RTOL + 2 * noise[s][0], min(PROB_ATOL + 2 * noise[s][1], 2e-3)
        assert d_err < d_tol, f"{tag}/{s}: depth rel err {d_err:.3e} (tol {d_tol:.3e})"
        assert p_err < p_tol, f"{tag}/{s}: prob abs err {p_err:.3e} (tol {p_tol:.3e}, reference noise {noise[s][1]:.3e})"


@pytest.mark.parametrize("name", ["msred_small_d8", "msred_batch2_d6"])
@pytest.mark.parametrize("cls", ["whole", "stream"])
def test_msred_forward_matches_reference_golden(name, cls):
    from oracle import msrednet_oracle as MO
    from tests.helpers import rebuild_msred_case
    g = load_golden(name)
    sd, imgs, proj, dv2, dv3, meta = rebuild_msred_case(g)
    if cls == "whole":
        run = lambda: MO.cascade_rednet_forward(sd, imgs, proj, dv3, ndepths=meta["ndepths"], ratios=(4.0, 2.0, 1.0))
    else:
        run = lambda: MO.infer_cascade_rednet_forward(sd, imgs, proj, dv2, num_depth=meta["num_depth"],
                                                      ndepths=meta["ndepths"], ratios=(4.0, 2.0, 1.0))
    _, noise = _arithmetic_noise(run)
    m = _msred_model(cls, sd, meta["ndepths"], meta["num_depth"])
    out = m(imgs.to(_dev()), _to_dev(proj), (dv3 if cls == "whole" else dv2).to(_dev()))
    want = {s: {"depth": torch.from_numpy(g[f"{cls}_{s}_depth"]), "photometric_confidence": torch.from_numpy(g[f"{cls}_{s}_conf"])}
            for s in ("stage1", "stage2", "stage3")}
    _compare_msred(out, want, noise, f"{name}/{cls}")
    assert out["depth"] is out["stage3"]["depth"]


def test_msred_forward_full_size_vs_oracle():
    """BASELINE config 5 shape at reduced depth counts (the CPU oracle needs ~1 s per plane at 768x384):
    5-view 768x384, ndepths 8/4/2, Infer_CascadeREDNet."""
    from adamvs_b200 import synth
    from oracle import msrednet_oracle as MO
    imgs, proj, dv2 = synth.make_sample(1, 384, 768, 5, seed=23)
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 37)
    f = MO.feature_net(sd, imgs[:, 0])
    sd = synth.calibrate_msred_state_dict(sd, {k: float(f[k].std()) for k in f}, 4.0)
    want, noise = _arithmetic_noise(lambda: MO.infer_cascade_rednet_forward(sd, imgs, proj, dv2, num_depth=32, ndepths=(8, 4, 2)))
    m = _msred_model("stream", sd, (8, 4, 2), 32)
    out = m(imgs.to(_dev()), _to_dev(proj), dv2.to(_dev()))
    assert tuple(out["stage1"]["depth"].shape) == (1, 96, 192) and tuple(out["depth"].shape) == (1, 384, 768)
    _compare_msred(out, want, noise, "msred-full")
    print("msred full-size: reference arithmetic noise (depth rel, prob abs) per stage:", noise)


def test_msred_forward_config5_d128_vs_oracle():
    """BASELINE configs[4] as written: Infer_CascadeREDNet, 5-view 768x384, ndepths 128/32/8, num_depth 512.

    MS-REDNet's probabilities are an ill-conditioned function of the warp's sample positions (variance by
    E[x^2] - E[x]^2, then a one-group GroupNorm per plane): the REFERENCE moves by `noise` (2e-3 at stage 2 here) when one
    4x4 product of its own forward is evaluated in fp64 instead of fp32.  Ours is a third valid rounding of the same
    function, so the bar is stated against that ball, in bulk and at the maximum, against BOTH reference roundings:
      depth: 1e-4 relative + 2 x noise (the north_star's figure holds: noise is ~1e-5)
      prob:  99th percentile <= 1e-4 + 2 x p99(noise);  maximum <= 1e-4 + 3 x max(noise)
    The kernel-level tests (variance volume and regulariser on identical inputs) hold 1e-4 unconditionally."""
    from adamvs_b200 import synth
    from oracle import msrednet_oracle as MO
    nd = (128, 32, 8)
    imgs, proj, dv2 = synth.make_sample(1, 384, 768, 5, seed=29)
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 43)
    f = MO.feature_net(sd, imgs[:, 0])
    sd = synth.calibrate_msred_state_dict(sd, {k: float(f[k].std()) for k in f}, 4.0)
    run = lambda: MO.infer_cascade_rednet_forward(sd, imgs, proj, dv2, num_depth=512, ndepths=nd)
    ref32 = run()
    O.RELPROJ_FP64 = True                                    # msrednet_oracle warps through adamvs_oracle
    try:
        ref64 = run()
    finally:
        O.RELPROJ_FP64 = False
    m = _msred_model("stream", sd, nd, 512)
    out = m(imgs.to(_dev()), _to_dev(proj), dv2.to(_dev()))
    assert tuple(out["stage1"]["depth"].shape) == (1, 96, 192) and tuple(out["depth"].shape) == (1, 384, 768)

    def stats(a, b):
        e = (a.double() - b.double()).abs().flatten()
        return float(e.mean()), float(torch.quantile(e[:: max(1, e.numel() // 200000)], 0.99)), float(e.max())
    report = {}
    for s in ("stage1", "stage2", "stage3"):
        mine_p, mine_d = out[s]["photometric_confidence"].cpu(), out[s]["depth"].cpu()
        n_mean, n_p99, n_max = stats(ref64[s]["photometric_confidence"], ref32[s]["photometric_confidence"])
        d_noise = rel_err(ref64[s]["depth"], ref32[s]["depth"])
        report[s] = {"noise(mean,p99,max)": (n_mean, n_p99, n_max)}
        for tag, ref in (("fp32", ref32), ("fp64", ref64)):
            e_mean, e_p99, e_max = stats(mine_p, ref[s]["photometric_confidence"])
            d_err = rel_err(mine_d, ref[s]["depth"])
            report[s][f"ours_vs_ref_{tag}(mean,p99,max,depth_rel)"] = (e_mean, e_p99, e_max, d_err)
            assert d_err < DEPTH_RTOL + 2 * d_noise, f"msred-d128/{s}/{tag}: depth rel err {d_err:.3e} (noise {d_noise:.3e})"
            assert e_p99 < PROB_ATOL + 2 * n_p99, f"msred-d128/{s}/{tag}: prob p99 {e_p99:.3e} (noise p99 {n_p99:.3e})"
            assert e_max < PROB_ATOL + 3 * n_max, f"msred-d128/{s}/{tag}: prob max {e_max:.3e} (noise max {n_max:.3e})"
    print("msred D=128/32/8 prob abs err:", report)


@pytest.mark.parametrize("cfg", ["0", "1", "2"])
def test_conv_tile_configurations_forced(cfg):
    """The persistent conv kernels pick one of three tile configurations from the plane size (32x16 tiles with 4x2
    patches, 32x8 with 4x1, the latter with split-K).  The large one is only chosen for batched full-size planes,
    which the CPU oracle cannot check in seconds, so each configuration is forced (ADAMVS_CONV_CFG) in a fresh
    process and the K3 and K6 kernel tests are re-run under it."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ADAMVS_CONV_CFG=cfg)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_regnet_red_vs_oracle or test_regnet_msred_vs_oracle or test_forward_matches_reference_golden"],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.parametrize("math", ["tc", "ffma"])
def test_msred_regulariser_math_paths_forced(math):
    """MS-REDNet's GRU convolutions of levels 1-3 run on the tensor-core kernel (EPI_RAW_STATS epilogue: raw output +
    GroupNorm moments) when a plane has >= 30 k pixels, which the CPU oracle cannot check in seconds: both paths are
    forced (ADAMVS_K3_MATH) in a fresh process and the K6 kernel test and the whole MS-REDNet forwards re-run."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ADAMVS_K3_MATH=math)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_regnet_msred_vs_oracle or test_msred_forward_matches_reference_golden or test_msred_forward_full_size_vs_oracle"],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("var,cfg", [("ADAMVS_TAIL_CFG", "0"), ("ADAMVS_TAIL_CFG", "16"), ("ADAMVS_TAIL_CFG", "24"), ("ADAMVS_DECONV_CFG", "0")])
def test_tail_and_transposed_conv_variants_forced(var, cfg):
    """The regulariser's tail picks the plain kernel (rows TMA cannot address), or the TMA-fed one with 32x16 or 32x24
    tiles (the latter only for thousands of tiles); the transposed convolutions pick the plain or the TMA-fed kernel
    from the row width.  Each variant is forced in a fresh process and the kernel tests and a whole forward re-run."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, **{var: cfg})
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_regnet_red_vs_oracle or test_native_deconv3x3_vs_torch or test_forward_matches_reference_golden or "
                        "test_feature_net_and_pair_unet_native_vs_oracle"],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.parametrize("cfg", ["1", "2", "3"])
def test_warp_box_cuts_forced(cfg):
    """The TMA cost-volume kernel re-cuts its stage buffer per block into 4x12-, 2x24- or 1x48-row boxes from the
    footprint of the tile, and takes the global-gather path beyond that.  Smooth synthetic depth only ever needs the
    first cut, so the other cuts (ADAMVS_WARP_CFG=1|2) and the gather path (3) are forced in a fresh process and the
    K1/K2/K5 kernel tests and a whole forward are re-run under them."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, ADAMVS_WARP_CFG=cfg)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_pair_score_and_fused_volume_vs_oracle or test_cost_volume_fallback_paths_vs_oracle or "
                        "test_fused_volume_zero_padding or test_variance_volume or test_forward_matches_reference_golden or "
                        "test_cost_volume_rough_depth"],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]


@pytest.mark.parametrize("C,D,h,w,sigma", [(16, 8, 64, 96, 25.0), (8, 4, 96, 128, 8.0), (16, 8, 64, 96, 120.0)])
def test_cost_volume_rough_depth(C, D, h, w, sigma):
    """Per-pixel hypotheses around a white-noise depth map (what an untrained stage 1 hands to stage 2): tile
    footprints spread over tens of source rows, so blocks pick the taller box cuts on their own (and the gather path
    for the largest sigma).  Same oracle, same tolerance as the smooth case."""
    ops = _ops()
    from adamvs_b200 import synth
    B, V = 2, 5
    g = torch.Generator().manual_seed(int(sigma) + C)
    feat = torch.randn(B, V, C, h, w, generator=g)
    proj = torch.stack([synth.make_cameras(h, w, V - 1)["stage3"], synth.make_cameras(h, w, V - 1, jitter_seed=4)["stage3"]])
    dv = torch.tensor([[520.0, 680.0], [540.0, 660.0]])
    wts = torch.rand(B, V - 1, h, w, generator=g) * 0.9 + 0.05
    relproj, half = ops.cascade_prepare([proj.to(_dev())] * 3, dv.to(_dev()), ops.INTERVAL_FROM_RANGE, 192,
                                        [D, D, D], [4.0, 2.0, 1.0])
    cur = 600 + sigma * torch.randn(B, h, w, generator=g)
    interval = (float(dv[0, 1]) - float(dv[0, 0])) / 192
    hyp = ops.Hyp(ops.HYP_PER_PIXEL, cur.to(_dev()), half[1:2])
    hyps = O.depth_hypotheses(cur, D, 2.0 * interval, [B, h, w])
    prods = [feat[:, 0].unsqueeze(2) * O.homography_warp(feat[:, v], proj[:, v], proj[:, 0], hyps) for v in range(1, V)]
    num = sum(p * wts[:, v].unsqueeze(1).unsqueeze(1) for v, p in enumerate(prods))
    want = num / (1e-5 + wts.sum(1).unsqueeze(1).unsqueeze(1))
    got = ops.fused_volume(feat.to(_dev()), relproj[0], hyp, wts.to(_dev()), ops.EPS_DENOMINATOR, D).cpu()
    assert float(want.abs().max()) > 0.1
    assert abs_err(got, want) < 2e-4 * float(want.abs().max())
    score_want = torch.stack([p.mean(1) for p in prods], 1)
    score = ops.pair_score(feat.to(_dev()), relproj[0], hyp, D).cpu()
    assert abs_err(score, score_want) < 2e-4 * float(score_want.abs().max())


@pytest.mark.parametrize("C,D,h,w,up,prob", [(32, 6, 16, 24, True, "softmax"), (16, 4, 32, 48, True, "exp"),
                                             (8, 3, 64, 96, False, "softmax"), (16, 3, 96, 160, True, "softmax")])
def test_regnet_red_tensor_core_vs_oracle(C, D, h, w, up, prob):
    """K3 with its stride-1 convolutions on tcgen05 (kind::tf32 with the exact hi/lo operand split, fp32 accumulation
    in TMEM): the same oracle and the same fp32 tolerances as the FFMA path."""
    test_regnet_red_vs_oracle(C, D, h, w, up, prob, math="tc")


@pytest.mark.parametrize("ca,cb,cout,stride,relu,h,w", [
    (8, 0, 8, 1, True, 40, 64), (16, 0, 16, 1, True, 36, 52), (32, 0, 32, 1, True, 24, 32), (16, 16, 16, 1, True, 32, 64),
    (8, 8, 8, 1, True, 64, 96), (48, 0, 48, 1, False, 24, 48), (48, 0, 48, 2, True, 24, 48), (48, 0, 48, 1, True, 18, 30),
    (32, 0, 16, 1, True, 64, 96), (64, 0, 32, 1, True, 32, 48), (64, 0, 32, 1, True, 30, 44),
    # >= 30000 pixels: the stride-1 layers switch to the tcgen05 kernel (ragged tiles: 100 = 3 x 30 + 10 columns, 106 rows)
    (8, 0, 8, 1, True, 106, 100), (16, 0, 16, 1, True, 106, 100), (32, 0, 32, 1, False, 106, 100), (16, 16, 16, 1, True, 106, 100),
    (8, 8, 8, 1, True, 106, 100), (32, 0, 16, 1, True, 106, 100), (64, 0, 32, 1, True, 106, 100),
    (48, 0, 48, 1, True, 106, 100)])                        # two 24-channel slices on the tensor cores
def test_native_conv3x3_vs_torch(ca, cb, cout, stride, relu, h, w):
    """The 3x3 convolutions of FeatureNet0 / CostRegNet2D on the native kernels (FFMA: TMA and generic-tile paths;
    tcgen05 with the exact hi/lo tf32 split for large stride-1 layers) against F.conv2d in fp32 on the CPU: same math,
    different summation order."""
    ops = _ops()
    g = torch.Generator().manual_seed(ca + cout + h)
    N = 3
    xa = torch.randn(N, ca, h, w, generator=g)
    xb = torch.randn(N, cb, h, w, generator=g) if cb else None
    wt = torch.randn(cout, ca + cb, 3, 3, generator=g) / (3.0 * (ca + cb) ** 0.5)
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(xa if xb is None else torch.cat((xa, xb), 1), wt, bias, stride, 1)
    want = F.relu(want) if relu else want
    got = ops.conv3x3(xa.to(_dev()), None if xb is None else xb.to(_dev()), ops.pack_conv3x3_weight(wt).to(_dev()),
                      bias.to(_dev()), relu, stride).cpu()
    assert tuple(got.shape) == tuple(want.shape)
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("cx,cctx,cout,h,w", [(32, 16, 32, 24, 48), (16, 8, 16, 48, 96), (8, 4, 8, 96, 192), (8, 4, 8, 40, 72)])
def test_context_head_vs_torch(cx, cctx, cout, h, w):
    """FeatureNet0's output heads fused (both bilinear upsamplings + cat + 1x1 conv) against the torch ops the reference
    uses (F.interpolate align_corners=False, torch.cat, F.conv2d) in fp32 on the CPU."""
    ops = _ops()
    g = torch.Generator().manual_seed(cx + h)
    N = 3
    x = torch.randn(N, cx, h, w, generator=g)
    a = torch.randn(N, cctx, h // 4, w // 4, generator=g)
    c = torch.randn(N, cctx, h // 8, w // 8, generator=g)
    wt = torch.randn(cout, 2 * cctx + cx, 1, 1, generator=g) / (2 * cctx + cx) ** 0.5
    up = lambda t: F.interpolate(t, size=(h, w), mode="bilinear", align_corners=False)
    want = F.conv2d(torch.cat((up(a), up(c), x), 1), wt)
    got = ops.context_head(x.to(_dev()), a.to(_dev()), c.to(_dev()), wt.to(_dev())).cpu()
    assert tuple(got.shape) == tuple(want.shape)
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("cin,cout,h,w", [(32, 16, 12, 20), (16, 8, 24, 36), (48, 48, 6, 10),       # w % 4 != 0: the plain kernel
                                          (16, 8, 21, 72), (32, 16, 9, 100), (48, 48, 18, 40)])     # several tiles, ragged edges
def test_native_deconv3x3_vs_torch(cin, cout, h, w):
    """Transposed 3x3 stride-2 convolution (+ bias, ReLU, optional skip addition after the activation): the TMA-fed
    persistent kernel (w % 4 == 0) and the plain one against torch."""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(3, cin, h, w, generator=g)
    wt = torch.randn(cin, cout, 3, 3, generator=g) / (3.0 * cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    skip = torch.randn(3, cout, 2 * h, 2 * w, generator=g)
    want = F.relu(F.conv_transpose2d(x, wt, bias, stride=2, padding=1, output_padding=1))
    wpk = ops.pack_deconv3x3_weight(wt).to(_dev())
    got = ops.deconv3x3(x.to(_dev()), wpk, bias.to(_dev()), True).cpu()
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))
    got = ops.deconv3x3(x.to(_dev()), wpk, bias.to(_dev()), True, skip.to(_dev())).cpu()
    assert abs_err(got, want + skip) < 2e-5 * max(1.0, float((want + skip).abs().max()))
    want = F.conv_transpose2d(x, wt, bias, stride=2, padding=1, output_padding=1)
    got = ops.deconv3x3(x.to(_dev()), wpk, bias.to(_dev()), False).cpu()
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))


def test_feature_net_and_pair_unet_native_vs_oracle():
    """FeatureNet0 and CostRegNet2D with every supported conv on the native kernels (3x3, polyphase 5x5 stride 2,
    transposed, zero-padded 3-channel input) against the oracle's torch-CPU restatement of the reference."""
    from adamvs_b200 import synth
    import models.adamvs as M
    sd = synth.fill_state_dict(synth.state_dict_shapes(48), 5)
    imgs, _, _ = synth.make_sample(1, 128, 192, 5, seed=2)
    m = _model("stream", sd, (48, 32, 8), 192)
    want = O.feature_net(sd, imgs[0])
    assert M._NATIVE_CONV
    from adamvs_b200.cascade import _true_fp32            # the product runs the remaining cuDNN pieces with TF32 off
    with torch.no_grad(), _true_fp32():
        got = m.feature(imgs[0].to(_dev()))
    for k in want:
        assert abs_err(got[k].cpu(), want[k]) < 2e-5 * max(1.0, float(want[k].abs().max())), k
    g = torch.Generator().manual_seed(3)
    score = torch.randn(4, 48, 32, 48, generator=g)
    with torch.no_grad(), _true_fp32():
        got = m.DepthNet[0].reg(score.to(_dev())).cpu()
    want = O.pair_unet(sd, "DepthNet.0.reg", score)
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("cin,cout,h,w", [(8, 16, 128, 192), (16, 32, 64, 96), (16, 32, 36, 52)])
def test_polyphase_5x5_stride2_vs_torch(cin, cout, h, w):
    """5x5 stride-2 convolution = pixel_unshuffle + 3x3 stride-1 convolution with re-laid weights, on the native kernel."""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 5, 5, generator=g) / (5.0 * cin ** 0.5)
    bias = torch.randn(cout, generator=g)
    want = F.relu(F.conv2d(x, wt, bias, 2, 2))
    wpk = ops.pack_conv3x3_weight(ops.polyphase_5x5_s2_weight(wt)).to(_dev())
    got = ops.conv3x3(F.pixel_unshuffle(x.to(_dev()), 2), None, wpk, bias.to(_dev()), True, 1).cpu()
    assert abs_err(got, want) < 2e-5 * max(1.0, float(want.abs().max()))


# ------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE's full sizes (the CPU oracle is too slow there)
# ------------------------------------------------------------------------------------------------

def _random_calibrated_model(cls_name, ndepths, num_depth, H, W, seed, gain=20.0):
    # gain 20: the predict class' un-shifted exp overflows (inf/inf) for logits > 88, like the reference's
    from adamvs_b200 import synth
    import models.adamvs as A
    sd = synth.fill_state_dict(synth.state_dict_shapes(ndepths[0]), seed)
    m = getattr(A, cls_name)(**({"num_depth": num_depth} if cls_name.startswith("Infer") else {}), ndepths=list(ndepths),
                             depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    m = m.to(_dev()).eval()
    imgs, proj, dv = synth.make_sample(1, H, W, 5, seed=seed)
    from adamvs_b200.cascade import _true_fp32
    with torch.no_grad(), _true_fp32():
        f = m.feature(imgs[:, 0].to(_dev()))
    sd = synth.calibrate_state_dict(sd, {k: float(f[k].std()) for k in f}, gain)
    m.load_state_dict(sd)
    return m


def test_batch_items_are_independent_at_full_size():
    """configs[2]: a batch of reference views must give, item by item, exactly what single-view forwards give
    (views are independent units; this is what lets them be sharded over GPUs without a collective).  The batched
    launch picks other tile configurations than the single-view one, so agreement is to fp32 round-off, not bitwise."""
    from adamvs_b200 import synth
    m = _random_calibrated_model("Infer_AdaMVSNet", (48, 32, 8), 192, 384, 768, seed=5)
    imgs, proj, dv = synth.make_sample(3, 384, 768, 5, seed=77)
    with torch.no_grad():
        full = m(imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
        for b in (0, 2):
            one = m(imgs[b:b + 1].to(_dev()), {k: v[b:b + 1].to(_dev()) for k, v in proj.items()}, dv[b:b + 1].to(_dev()))
            for s in ("stage1", "stage2", "stage3"):
                assert rel_err(full[s]["depth"][b:b + 1].cpu(), one[s]["depth"].cpu()) < DEPTH_RTOL
                assert abs_err(full[s]["photometric_confidence"][b:b + 1].cpu(), one[s]["photometric_confidence"].cpu()) < PROB_ATOL


@pytest.mark.parametrize("cls_name", ["Infer_AdaMVSNet", "AdaMVSNet"])
def test_config4_oblique_tile_properties(cls_name):
    """configs[3]: 5-view 1536x1536 tile with widened first-stage hypotheses (ndepths 96/32/8).  Properties that hold
    for any weights: output shapes (H/2, H, H), probabilities in (0, 1], every depth inside the hull of its own
    hypotheses (an expectation over them), stage-1 depth inside [min, max], and repeatability of the forward."""
    from adamvs_b200 import synth
    H = W = 1536
    nd = (96, 32, 8)
    m = _random_calibrated_model(cls_name, nd, 192, H, W, seed=9)
    imgs, proj, dv2 = synth.make_sample(1, H, W, 5, seed=9)
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / 192
    dv = dv2 if cls_name.startswith("Infer") else torch.cat([dv2, torch.full((1, 1), interval)], 1)
    with torch.no_grad():
        out = m(imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
        again = m(imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
    assert tuple(out["stage1"]["depth"].shape) == (1, H // 2, W // 2)
    assert tuple(out["stage2"]["depth"].shape) == (1, H, W) and tuple(out["stage3"]["depth"].shape) == (1, H, W)
    for s in ("stage1", "stage2", "stage3"):
        d, c = out[s]["depth"], out[s]["photometric_confidence"]
        assert torch.isfinite(d).all() and torch.isfinite(c).all()
        assert float(c.min()) > 0.0 and float(c.max()) <= 1.0 + 1e-6
        # repeatable (the cuDNN/cuBLAS pieces of the feature heads are not bitwise deterministic, hence a tolerance)
        assert rel_err(d.cpu(), again[s]["depth"].cpu()) < DEPTH_RTOL and abs_err(c.cpu(), again[s]["photometric_confidence"].cpu()) < PROB_ATOL
    d1 = out["stage1"]["depth"]
    assert float(d1.min()) >= synth.DEPTH_MIN - 1e-2 and float(d1.max()) <= synth.DEPTH_MAX + 1e-2
    # later stages: depth within +-(ndepth/2 * ratio * interval) of the previous stage's (resized) depth
    for s, prev, ndp, ratio in (("stage2", "stage1", 32, 2.0), ("stage3", "stage2", 8, 1.0)):
        p = out[prev]["depth"]
        if p.shape != out[s]["depth"].shape[-2:]:
            pass
        # hypotheses are generated at the feature resolution of the stage and up-sampled x2 in stage 2: compare against
        # the hull over a 3x3 neighbourhood of the previous depth at the output resolution
        hr = ndp / 2 * ratio * interval
        src = p if tuple(p.shape[1:]) == tuple(out[s]["depth"].shape[1:]) else \
            F.interpolate(p.unsqueeze(1), size=out[s]["depth"].shape[1:], mode="nearest").squeeze(1)
        lo = -F.max_pool2d(-src.unsqueeze(1), 5, 1, 2).squeeze(1) - hr - 1e-2
        hi = F.max_pool2d(src.unsqueeze(1), 5, 1, 2).squeeze(1) + hr + 1e-2
        dd = out[s]["depth"]
        assert bool(((dd >= lo) & (dd <= hi)).all()), s


@pytest.mark.parametrize("cls", ["stream", "whole"])
def test_config4_oblique_tile_vs_oracle(cls):
    """configs[3] against the oracle: 5-view 1536x1536, ndepths 96/32/8, calibrated seeded weights (the CPU oracle needs
    ~30-45 s per forward at this size; it runs twice).

    Depth holds the north_star's 1e-4 relative at every stage.  Probability: at x ~ 1500 one fp32 ulp of a sample
    coordinate is 1.2e-4 px, twice that of the 768-wide case, and the REFERENCE's own probabilities move by up to
    1.3e-4 (stage 2) / 1.7e-4 (stage 3) when one 4x4 product of its forward (src_proj @ inverse(ref_proj)) is evaluated in
    fp64 instead of fp32 (`noise`, measured below with the oracle's RELPROJ_FP64 switch).  No implementation can sit
    inside 1e-4 of a function that is only defined to 1.7e-4, so the bar is: 99.9 % of the pixels within the
    north_star's 1e-4 absolute, and the maximum within 1e-4 + 2 x the reference's own noise maximum."""
    nd = (96, 32, 8)
    sd, imgs, proj, dv2, dv3 = make_case(1, 1536, 1536, nd, 192, 60.0, 33, 19, O.feature_net)
    if cls == "whole":
        run = lambda: O.adamvs_forward(sd, imgs, proj, dv3, ndepths=nd)
    else:
        run = lambda: O.infer_adamvs_forward(sd, imgs, proj, dv2, num_depth=192, ndepths=nd)
    want, noise = _arithmetic_noise(run)
    m = _model(cls, sd, nd, 192)
    with torch.no_grad():
        out = m(imgs.to(_dev()), _to_dev(proj), (dv3 if cls == "whole" else dv2).to(_dev()))
    assert tuple(out["stage1"]["depth"].shape) == (1, 768, 768) and tuple(out["depth"].shape) == (1, 1536, 1536)
    report = {}
    for s in ("stage1", "stage2", "stage3"):
        d_err = rel_err(out[s]["depth"].cpu(), want[s]["depth"])
        e = (out[s]["photometric_confidence"].cpu().double() - want[s]["photometric_confidence"].double()).abs().flatten()
        p999, p_max = float(torch.quantile(e[::8], 0.999)), float(e.max())
        report[s] = {"depth_rel": d_err, "prob_p99.9": p999, "prob_max": p_max, "reference_noise(depth_rel, prob_max)": noise[s]}
        assert d_err < DEPTH_RTOL, f"config4/{cls}/{s}: depth rel err {d_err:.3e}"
        assert p999 < PROB_ATOL, f"config4/{cls}/{s}: prob p99.9 {p999:.3e}"
        assert p_max < PROB_ATOL + 2 * noise[s][1], f"config4/{cls}/{s}: prob max {p_max:.3e} (reference noise {noise[s][1]:.3e})"
    print(f"config4/{cls}:", report)
    assert float(want["stage1"]["photometric_confidence"].max()) > 5.0 / 96        # not the vacuous 1/D regime
    assert float(want["stage3"]["photometric_confidence"].max()) > 0.5


def test_ops_follow_the_tensors_device_not_the_current_one():
    """the C library launches on the current device; ops must make the tensors' device current.  With one
    visible GPU the guard is exercised through a non-default current stream on the same device, and tensors of two
    devices in one call are refused (when a second GPU exists, the cross-device launch itself is checked)."""
    ops = _ops()
    x = torch.randn(2, 3, 8, 12, device=_dev())
    want = F.interpolate(x, size=(16, 24), mode="bilinear", align_corners=False)
    side = torch.cuda.Stream(device=_dev())
    side.wait_stream(torch.cuda.current_stream(_dev()))
    with torch.cuda.stream(side):
        got = ops.resize_bilinear(x, 16, 24)
    side.synchronize()
    assert abs_err(got.cpu(), want.cpu()) < 1e-6
    if torch.cuda.device_count() > 1:
        d1 = torch.device("cuda:1")
        x1 = x.to(d1)
        assert torch.cuda.current_device() == 0
        got1 = ops.resize_bilinear(x1, 16, 24)                # current device is cuda:0, the tensor lives on cuda:1
        torch.cuda.synchronize(d1)
        assert got1.device == d1 and abs_err(got1.cpu(), want.cpu()) < 1e-6
        feat = torch.randn(1, 3, 8, 8, 16, device=_dev())
        with pytest.raises(ops.AdamvsError):
            ops.fused_volume(feat, torch.zeros(1, 2, 12, device=d1), ops.Hyp(ops.HYP_PLANES, torch.tensor([[1.0, 2.0]], device=_dev())),
                             torch.ones(1, 2, 8, 16, device=_dev()), ops.EPS_NUMERATOR, 4)


def test_predict_scene_writes_the_reference_output_files(tmp_path):
    """Scene folder -> batches -> model -> `*_init.pfm`, `*_prob.pfm`, `*.txt`: the files hold exactly
    what a direct forward on the same preprocessed inputs returns, in the reference's formats."""
    import os
    from PIL import Image
    from adamvs_b200 import pipeline, sceneio as S, synth
    from models.adamvs import Infer_AdaMVSNet
    scene = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_scene")
    ndepths = (8, 4, 2)
    sd = synth.fill_state_dict(synth.state_dict_shapes(ndepths[0]), 5)
    m = Infer_AdaMVSNet(num_depth=32, ndepths=list(ndepths), depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    m = m.to(_dev()).eval()
    written = pipeline.predict_scene(m, scene, str(tmp_path), view_num=3, num_depth=32, max_h=64, max_w=96, batch=2, device=_dev())
    assert len(written) == 3                                     # view 2 has no source views and is skipped
    cams, poses = S.parse_camera_info(os.path.join(scene, "camera_info.txt")), S.parse_image_info(os.path.join(scene, "image_info.txt"))
    paths, _ = S.parse_image_paths(os.path.join(scene, "image_path.txt"))
    rows = S.parse_view_pairs(os.path.join(scene, "viewpair.txt"), 3)
    for row, path in zip(rows, written):
        images = [np.array(Image.open(os.path.join(scene, paths[i]))) for i in row[:3]]
        imgs, proj, dv, _, blk = S.load_view_sample(row, poses, cams, images, 3, num_depth=32, max_h=64, max_w=96, device=_dev())
        with torch.no_grad():
            out = m(imgs[None], {k: torch.from_numpy(v)[None].to(_dev()) for k, v in proj.items()}, torch.from_numpy(dv)[None].to(_dev()))
        depth, scale = S.read_pfm(path)
        assert scale == 1.0 and depth.shape == (64, 96)
        # batch of 2 vs 1: the device-side normalisation reduces in another order (1e-6 on the inputs)
        assert rel_err(torch.from_numpy(depth.copy()), out["depth"][0].cpu()) < 1e-5
        prob, _ = S.read_pfm(path.replace("_init.pfm", "_prob.pfm"))
        assert abs_err(torch.from_numpy(prob.copy()), out["photometric_confidence"][0].cpu()) < 1e-5
        assert open(path.replace("_init.pfm", ".txt")).read().startswith("extrinsic: XrightYdown, [Rcw|tcw]")
