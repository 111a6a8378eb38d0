"""Training path: forward under model.train() and the backward kernels against the UNMODIFIED
reference's autograd (tests/golden/grad_small_d8.npz, made by tests/golden/make_grad_golden.py from /root/reference:
one train_whu.py step - forward, cas_mvs_vis_loss with dlossw 0.5/1/2, backward - on seeded inputs)."""
import numpy as np
import pytest
import torch

from adamvs_b200 import synth
from tests.helpers import abs_err, load_golden, rel_err

DLOSSW = [0.5, 1.0, 2.0]


def _case():
    from tests.golden.make_grad_golden import ground_truth
    g = load_golden("grad_small_d8")
    B, H, W = int(g["meta_B"]), int(g["meta_H"]), int(g["meta_W"])
    nd = tuple(int(x) for x in g["meta_ndepths"])
    imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=int(g["meta_iseed"]))
    interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / int(g["meta_num_depth"])
    dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
    sd = synth.fill_state_dict(synth.state_dict_shapes(nd[0]), int(g["meta_wseed"]))
    fstd = dict(zip(("stage1", "stage2", "stage3"), [float(x) for x in g["meta_fstd"]]))
    sd = synth.calibrate_state_dict(sd, fstd, float(g["meta_gain"]))
    gt, mask = ground_truth(B, H, W, int(g["meta_gtseed"]))
    return g, sd, imgs, proj, dv3, gt, mask, nd


def test_ground_truth_generator_is_deterministic():
    from tests.golden.make_grad_golden import ground_truth
    a, _ = ground_truth(1, 64, 96, 101)
    b, m = ground_truth(1, 64, 96, 101)
    assert torch.equal(a["stage3"], b["stage3"]) and tuple(a["stage1"].shape) == (1, 16, 24) and float(m["stage2"].min()) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,stride,transposed,relu,h,w", [
    (8, 8, 1, False, True, 20, 36), (16, 16, 1, False, False, 18, 22), (16, 8, 1, False, False, 16, 16), (8, 16, 2, False, True, 20, 28),
    (32, 32, 1, False, False, 10, 14), (16, 8, 2, True, False, 9, 13), (8, 1, 2, True, False, 12, 10), (8, 1, 1, False, False, 17, 19)])
def test_conv3x3_function_forward_and_gradients_vs_torch(cin, cout, stride, transposed, relu, h, w):
    """The run-time-channel convolution kernels (forward, data gradient in its three roles, weight gradient) against
    torch's fp64 convolutions."""
    import torch.nn.functional as F
    from adamvs_b200 import autograd as A
    g = torch.Generator().manual_seed(cin * 100 + cout + stride)
    x = torch.randn(2, cin, h, w, generator=g).cuda().requires_grad_()
    wt = (torch.randn((cin, cout, 3, 3) if transposed else (cout, cin, 3, 3), generator=g) * 0.2).cuda().requires_grad_()
    b = torch.randn(cout, generator=g).cuda().requires_grad_()
    y = A.conv3x3(x, wt, b, stride=stride, transposed=transposed, relu=relu)
    xd, wd, bd = (t.detach().double().requires_grad_() for t in (x, wt, b))
    if transposed:
        yd = F.conv_transpose2d(xd, wd, bd, 2, 1, 1)
    else:
        yd = F.conv2d(xd, wd, bd, stride, 1)
    if relu:
        yd = F.relu(yd)
    assert tuple(y.shape) == tuple(yd.shape)
    assert abs_err(y.detach().cpu(), yd.detach().cpu()) < 2e-5 * max(1.0, float(yd.detach().abs().max()))
    go = torch.randn(y.shape, generator=g).cuda()
    y.backward(go)
    yd.backward(go.double())
    for got, want, name in ((x.grad, xd.grad, "dx"), (wt.grad, wd.grad, "dw"), (b.grad, bd.grad, "db")):
        assert abs_err(got.cpu(), want.cpu()) < 1e-4 * max(1.0, float(want.abs().max())), name


@pytest.mark.gpu
def test_cost_volume_and_regression_gradients_vs_oracle_autograd():
    """K1 / K2 / K4 backward kernels against autograd through the oracle's restatement of the same ops (grid_sample
    warp with the grid under no_grad, weighted aggregation, softmax expectation)."""
    import torch.nn.functional as F
    from adamvs_b200 import autograd as A, ops
    from oracle import adamvs_oracle as O
    B, V, C, D, h, w = 1, 4, 8, 5, 24, 32
    g = torch.Generator().manual_seed(7)
    feat = torch.randn(B, V, C, h, w, generator=g)
    cams = synth.make_cameras(4 * h, 4 * w, V - 1)["stage1"][None]
    wts = torch.rand(B, V - 1, h, w, generator=g) * 0.8 + 0.1
    dv = torch.tensor([[synth.DEPTH_MIN, synth.DEPTH_MAX]])
    hyps = O.depth_hypotheses(dv, D, 1.0, [B, h, w])
    # oracle: fp32 CPU autograd
    fo, wo = feat.clone().requires_grad_(), wts.clone().requires_grad_()
    ref_vol = fo[:, 0].unsqueeze(2)
    num, den, scores = 1e-5, 0, []
    for v in range(V - 1):
        warped = O.homography_warp(fo[:, v + 1], cams[:, v + 1], cams[:, 0], hyps)
        prod = ref_vol * warped
        scores.append(prod.mean(1))
        num = num + prod * wo[:, v].unsqueeze(1).unsqueeze(1)
        den = den + wo[:, v].unsqueeze(1).unsqueeze(1)
    vol_o, score_o = num / den, torch.stack(scores, 1)
    gv, gs = torch.randn(vol_o.shape, generator=g), torch.randn(score_o.shape, generator=g)
    (vol_o * gv).sum().backward(retain_graph=True)
    gf_vol, gw_vol = fo.grad.clone(), wo.grad.clone()
    fo.grad = None
    (score_o * gs).sum().backward()
    gf_score = fo.grad.clone()
    # ours
    dev = torch.device("cuda:0")
    relproj, _ = ops.cascade_prepare([cams.to(dev)] * 3, torch.cat([dv, torch.ones(1, 1)], 1).to(dev), ops.INTERVAL_LAST_COLUMN, 0,
                                     [D, D, D], [1.0, 1.0, 1.0])
    fm, wm = feat.to(dev).requires_grad_(), wts.to(dev).requires_grad_()
    vol = A.FusedVolumeFn.apply(fm, wm, relproj[0], ops.HYP_PLANES, dv.to(dev), None, ops.EPS_NUMERATOR, D)
    assert abs_err(vol.detach().cpu(), vol_o.detach()) < 2e-4 * float(vol_o.abs().max())
    (vol * gv.to(dev)).sum().backward()
    assert abs_err(fm.grad.cpu(), gf_vol) < 2e-4 * float(gf_vol.abs().max())
    assert abs_err(wm.grad.cpu(), gw_vol) < 2e-4 * float(gw_vol.abs().max())
    fm.grad = None
    score = A.PairScoreFn.apply(fm, relproj[0], ops.HYP_PLANES, dv.to(dev), None, D)
    (score * gs.to(dev)).sum().backward()
    assert abs_err(fm.grad.cpu(), gf_score) < 2e-4 * float(gf_score.abs().max())
    # K4
    lo = (torch.randn(2, D, h, w, generator=g) * 2).requires_grad_()
    hy = (600 + 10 * torch.randn(2, D, h, w, generator=g)).requires_grad_()
    p = F.softmax(lo, 1)
    d_o, c_o = (p * hy).sum(1), p.max(1)[0]
    gd, gc = torch.randn(d_o.shape, generator=g), torch.randn(c_o.shape, generator=g)
    ((d_o * gd).sum() + (c_o * gc).sum()).backward()
    lm, hm = lo.detach().to(dev).requires_grad_(), hy.detach().to(dev).requires_grad_()
    d_m, c_m = A.SoftmaxExpectFn.apply(lm, hm)
    assert rel_err(d_m.detach().cpu(), d_o.detach()) < 1e-5 and abs_err(c_m.detach().cpu(), c_o.detach()) < 1e-5
    ((d_m * gd.to(dev)).sum() + (c_m * gc.to(dev)).sum()).backward()
    assert abs_err(lm.grad.cpu(), lo.grad) < 1e-4 * float(lo.grad.abs().max())
    assert abs_err(hm.grad.cpu(), hy.grad) < 1e-5 * max(1.0, float(hy.grad.abs().max()))


@pytest.mark.gpu
def test_training_step_matches_reference_forward_loss_and_every_gradient():
    """One train_whu.py step against the drop-in: model.train(), forward, the reference's loss, backward; outputs, loss
    and the gradient of EVERY parameter the reference gives one to must agree with the unmodified reference (CPU):
    depth 1e-4 relative, probability 3e-4 absolute (train-mode BatchNorm normalises FeatureNet0 and the pair U-Net by
    batch statistics that cuDNN and the CPU reduce in different orders: measured 1.1e-4 at stage 1, against 2e-5 in
    eval mode), loss 1e-4 relative, every gradient tensor within 1e-2 of its largest entry and at
    cosine >= 0.9999 to the reference's, median mismatch < 1e-3 (fp32 atomics, another summation order, BatchNorm's
    batch statistics, and ReLU masks that flip where a pre-activation is ~0: the worst tensor measured 4.4e-3)."""
    from models.adamvs import AdaMVSNet, cas_mvs_vis_loss
    g, sd, imgs, proj, dv3, gt, mask, nd = _case()
    dev = torch.device("cuda:0")
    m = AdaMVSNet(ndepths=list(nd), depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(sd)
    m = m.to(dev).train()
    out = m(imgs.to(dev), {k: v.to(dev) for k, v in proj.items()}, dv3.to(dev))
    loss, _ = cas_mvs_vis_loss(out, {k: v.to(dev) for k, v in gt.items()}, {k: v.to(dev) for k, v in mask.items()}, dlossw=DLOSSW)
    loss.backward()
    for s in ("stage1", "stage2", "stage3"):
        d_err = rel_err(out[s]["depth"].detach().cpu(), g[f"{s}_depth"])
        p_err = abs_err(out[s]["photometric_confidence"].detach().cpu(), g[f"{s}_conf"])
        print(f"train forward {s}: depth rel {d_err:.3e}, prob abs {p_err:.3e}")
        assert d_err < 1e-4, s
        assert p_err < 3e-4, s
    print("loss", float(loss.detach()), "reference", float(g["loss"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * float(g["loss"])
    params = dict(m.named_parameters())
    want = {k[5:]: g[k] for k in g.files if k.startswith("grad_")}
    assert len(want) == 131
    worst = ("", 0.0)
    errs, cos = [], []
    for name, gw in want.items():
        got = params[name].grad
        assert got is not None, f"no gradient for {name}"
        scale = max(float(np.abs(gw).max()), 1e-12)
        err = abs_err(got.cpu(), gw) / scale
        if err > worst[1]:
            worst = (name, err)
        errs.append((err, name, scale))
        a, b2 = got.detach().cpu().double().flatten(), torch.from_numpy(np.asarray(gw)).double().flatten()
        cos.append((float((a * b2).sum() / (a.norm() * b2.norm() + 1e-300)), name, scale))
    errs.sort(reverse=True)
    gmax = max(sc for _, _, sc in errs)
    print("largest gradient mismatches (fraction of the tensor's largest entry, that entry):",
          [(n, f"{e:.2e}", f"{sc:.2e}") for e, n, sc in errs[:10]], "largest gradient entry overall:", gmax)
    # A tensor whose true gradient vanishes (the output layer's bias shifts all D logits of a pixel alike, which the
    # softmax ignores: both sides hold rounding noise ~1e-7 there) is compared on the scale of the other gradients.
    bad = [(n, e, sc) for e, n, sc in errs if e * sc >= 1e-2 * max(sc, 1e-4 * gmax)]
    assert not bad, bad
    med = sorted(e for e, _, sc in errs if sc > 1e-4 * gmax)
    print("median / 90th percentile mismatch:", med[len(med) // 2], med[(9 * len(med)) // 10])
    assert med[len(med) // 2] < 1e-3
    low = [(n, c) for c, n, sc in cos if sc > 1e-4 * gmax and c < 0.9999]
    assert not low, low                                              # direction of every gradient tensor
    unused = [n for n, p in params.items() if n not in want and p.grad is not None and float(p.grad.abs().max()) > 0]
    assert not unused, unused                                        # e.g. DepthNet.1/2.reg are never executed
    print("worst gradient mismatch:", worst)
    # and one optimiser step runs (train_whu.py:116, 283-284)
    opt = torch.optim.RMSprop(m.parameters(), lr=1e-3)
    opt.step()
    assert all(torch.isfinite(p).all() for p in m.parameters())
