"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares, the
drop-in modules keep the reference's names / signatures / state_dict layout, the host-side pieces (BN folding,
polyphase re-layout, losses) are exact, and the product refuses CPU tensors instead of falling back."""
import ctypes
import inspect
import os
import re
import sys

import pytest
import torch
import torch.nn.functional as F

from adamvs_b200 import ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                     reason="the reference tree exists only in the build container")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "adamvs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(adamvs_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from adamvs_b200 import build
    lib_path = build.build(force=False)
    L = ctypes.CDLL(lib_path)                          # no CUDA call happens at load time
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/adamvs_b200.h but not exported"
    assert sorted(ops.EXPORTS) == declared, "ops.EXPORTS and the header disagree"
    L.adamvs_abi_version.restype = ctypes.c_int
    assert L.adamvs_abi_version() == 1
    L.adamvs_conv3x3_supported.argtypes = [ctypes.c_int] * 4
    assert L.adamvs_conv3x3_supported(16, 16, 16, 1) == 1 and L.adamvs_conv3x3_supported(7, 0, 8, 1) == 0


def test_argument_errors_come_back_as_codes_without_touching_a_device():
    L = ops.lib()
    assert L.adamvs_fused_volume_f32(None, None, 0, None, 2, None, None, 0, None, 1, 5, 8, 4, 8, 8, None) == -1   # ADAMVS_EINVAL
    assert int(L.adamvs_regnet_red_workspace_floats(0, 8, 4, 8, 8, 0)) == 0
    assert int(L.adamvs_regnet_red_workspace_floats(1, 8, 4, 64, 96, 0)) > 8 * 64 * 96
    assert int(L.adamvs_regnet_msred_workspace_floats(1, 8, 4, 64, 96)) > 8 * 64 * 96
    # the fused FeatureNet0 head: the three channel combinations of the reference's pyramid and nothing else
    assert [L.adamvs_context_head_supported(*c) for c in ((32, 16, 32), (16, 8, 16), (8, 4, 8), (8, 8, 8), (64, 32, 64))] == [1, 1, 1, 0, 0]
    assert L.adamvs_context_head_f32(None, None, None, None, None, 1, 8, 4, 8, 16, 16, 4, 4, 2, 2, None) == -1
    # the arithmetic mode of the regulariser is validated before anything is launched
    assert L.adamvs_regnet_red_ex_f32(None, None, 0, None, 2, None, 0, 0, 7, None, 0, None, None, None, 1, 8, 4, 8, 8, None) == -1


def test_adamvs_drop_in_names_signatures_and_state_dict():
    import models.adamvs as M
    for name in ("AdaMVSNet", "Infer_AdaMVSNet", "cas_mvs_vis_loss"):
        assert hasattr(M, name)
    sig = inspect.signature(M.AdaMVSNet.__init__)
    assert list(sig.parameters)[1:] == ["ndepths", "depth_intervals_ratio", "share_cr", "cr_base_chs"]
    assert sig.parameters["ndepths"].default == [48, 32, 8] and sig.parameters["depth_intervals_ratio"].default == [4, 2, 1]
    sig = inspect.signature(M.Infer_AdaMVSNet.__init__)
    assert list(sig.parameters)[1:] == ["num_depth", "ndepths", "depth_intervals_ratio", "share_cr", "cr_base_chs"]
    assert sig.parameters["num_depth"].default == 384
    for cls, kw in ((M.AdaMVSNet, {}), (M.Infer_AdaMVSNet, {"num_depth": 192})):
        m = cls(ndepths=[48, 32, 8], depth_intervals_ratio=[4.0, 2.0, 1.0], **kw)
        have = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert have == synth.state_dict_shapes(48)
        # DataParallel checkpoints carry a "module." prefix (train_whu.py:127, predict_whu.py:88)
        sd = synth.fill_state_dict(synth.state_dict_shapes(48), 1)
        torch.nn.DataParallel(m).load_state_dict({"module." + k: v for k, v in sd.items()})


def test_msrednet_drop_in_names_and_signatures():
    import models.msrednet as M
    for name in ("CascadeREDNet", "Infer_CascadeREDNet", "cas_rednet_loss"):
        assert hasattr(M, name)
    sig = inspect.signature(M.Infer_CascadeREDNet.__init__)
    assert list(sig.parameters)[1:] == ["num_depth", "ndepths", "depth_interals_ratio", "share_cr", "cr_base_chs"]
    sig = inspect.signature(M.CascadeREDNet.__init__)
    assert list(sig.parameters)[1:] == ["ndepths", "depth_interals_ratio", "share_cr", "cr_base_chs"]


@needs_reference
def test_state_dict_layouts_equal_the_reference_modules():
    import contextlib
    import importlib
    import io
    import types
    pkg = types.ModuleType("refmodels_t")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["refmodels_t"] = pkg
    with contextlib.redirect_stdout(io.StringIO()):
        ra = importlib.import_module("refmodels_t.adamvs").Infer_AdaMVSNet(num_depth=192, ndepths=[48, 32, 8],
                                                                              depth_intervals_ratio=[4.0, 2.0, 1.0])
        rm = importlib.import_module("refmodels_t.msrednet").Infer_CascadeREDNet(num_depth=192, ndepths=[48, 32, 8],
                                                                                   depth_interals_ratio=[4.0, 2.0, 1.0])
    import models.adamvs as A
    import models.msrednet as M
    ours_a = A.Infer_AdaMVSNet(num_depth=192)
    ours_m = M.Infer_CascadeREDNet(num_depth=192)
    for ref, ours in ((ra, ours_a), (rm, ours_m)):
        r = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        o = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert list(r.keys()) == list(o.keys()) and r == o
        ours.load_state_dict(ref.state_dict())          # strict


@needs_reference
def test_losses_equal_the_reference():
    import importlib
    import types
    pkg = types.ModuleType("refmodels_l")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["refmodels_l"] = pkg
    ref_a = importlib.import_module("refmodels_l.adamvs")
    ref_m = importlib.import_module("refmodels_l.msrednet")
    import models.adamvs as A
    import models.msrednet as M
    g = torch.Generator().manual_seed(0)
    sizes = {"stage1": (24, 32), "stage2": (48, 64), "stage3": (48, 64)}
    gt = {k: 600 + 10 * torch.randn(1, 48, 64, generator=g) for k in sizes}       # the reference loss slices [0:1] (batch 1 only)
    mask = {k: (torch.rand(1, 48, 64, generator=g) > 0.3).float() for k in sizes}
    inputs = {k: {"depth": 600 + 10 * torch.randn(1, *s, generator=g),
                  "pair_result": [600 + 10 * torch.randn(1, *s, generator=g) for _ in range(4 if k == "stage1" else 0)]}
              for k, s in sizes.items()}
    for kw in ({}, {"dlossw": [0.5, 1.0, 2.0]}):
        a, b = ref_a.cas_mvs_vis_loss(inputs, gt, mask, **kw), A.cas_mvs_vis_loss(inputs, gt, mask, **kw)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    flat = {k: {"depth": 600 + 10 * torch.randn(1, 48, 64, generator=g)} for k in sizes}
    for kw in ({}, {"dlossw": [0.5, 1.0, 2.0]}):
        a, b = ref_m.cas_rednet_loss(flat, gt, mask, **kw), M.cas_rednet_loss(flat, gt, mask, **kw)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_bn_folding_and_conv_relayouts_are_exact_enough():
    import models.adamvs as M
    sd = synth.fill_state_dict(synth.state_dict_shapes(8), 3)
    m = M.AdaMVSNet(ndepths=[8, 4, 2])
    m.load_state_dict(sd)
    m.eval()
    x, s = torch.randn(2, 3, 64, 96), torch.randn(2, 8, 16, 24)
    with torch.no_grad():
        M._FOLD_BN = True
        a, ra = m.feature(x), m.DepthNet[0].reg(s)
        M._FOLD_BN = False
        try:
            b, rb = m.feature(x), m.DepthNet[0].reg(s)
        finally:
            M._FOLD_BN = True
    for k in a:
        assert float((a[k] - b[k]).abs().max()) < 1e-6
    assert float((ra - rb).abs().max()) < 2e-6
    # a weight update must invalidate the folded cache
    with torch.no_grad():
        m.feature.conv0[0].conv.weight.mul_(2.0)
        c = m.feature(x)
    assert float((c["stage3"] - a["stage3"]).abs().max()) > 1e-4
    # 5x5 stride-2 conv == pixel_unshuffle + 3x3 conv with the polyphase weights
    w = torch.randn(16, 8, 5, 5)
    xx = torch.randn(2, 8, 12, 20)
    ref = F.conv2d(xx, w, None, 2, 2)
    got = F.conv2d(F.pixel_unshuffle(xx, 2), ops.polyphase_5x5_s2_weight(w), None, 1, 1)
    assert float((ref - got).abs().max()) < 1e-4
    wp = ops.pack_conv3x3_weight(torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3))
    assert tuple(wp.shape) == (3, 9, 2) and float(wp[1, 4, 1]) == float(1 * 27 + 1 * 9 + 4)


def test_cpu_tensors_are_refused_not_emulated():
    import models.adamvs as A
    import models.msrednet as M
    imgs, proj, dv = synth.make_sample(1, 64, 96, 5, seed=1)
    with pytest.raises(ops.AdamvsError):
        A.Infer_AdaMVSNet(num_depth=32, ndepths=[8, 4, 2]).eval()(imgs, proj, dv)
    with pytest.raises(ops.AdamvsError):
        M.Infer_CascadeREDNet(num_depth=32, ndepths=[8, 4, 2]).eval()(imgs, proj, dv)
    # training needs backward kernels that do not exist yet: say so instead of silently running something else
    m = A.AdaMVSNet(ndepths=[8, 4, 2]).train()
    with pytest.raises((NotImplementedError, ops.AdamvsError)):
        m(imgs, proj, torch.cat([dv, torch.full((1, 1), 5.0)], 1))


def test_bench_accounting_matches_the_survey_definitions():
    """bench.py's roofline numerators are SURVEY.md §8(d)'s per-unit figures: 125.4 / 175.2 / 124.2 MB of algorithmic HBM
    bytes for the fused warp + cost volume and 17.45 / 41.11 / 38.39 GFLOP for the recurrent regulariser, per depth map."""
    sys.path.insert(0, ROOT)
    import bench
    shapes = ((32, 48, 96, 192), (16, 32, 192, 384), (8, 8, 384, 768))
    mb = [bench.costvolume_algorithmic_bytes(1, C, D, h, w, weight_px=96 * 192) / 1e6 for C, D, h, w in shapes]
    assert [round(v, 1) for v in mb] == [125.4, 175.2, 124.2]
    gf = [bench.regnet_flops(1, C, D, h, w) / 1e9 for C, D, h, w in shapes]
    assert [round(v, 2) for v in gf] == [17.45, 41.11, 38.39]
    assert all(0.9 < bench.regnet_tc_flops(1, C, D, h, w) / bench.regnet_flops(1, C, D, h, w) < 0.95 for C, D, h, w in shapes)
