"""CPU: pins oracle/msrednet_oracle.py against the reference's own MS-REDNet outputs
(tests/golden/msred_*.npz, generated from /root/reference by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import msrednet_oracle as MO
from tests.helpers import abs_err, load_golden, rebuild_msred_case, rel_err

CASES = ["msred_small_d8", "msred_batch2_d6"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cls", ["whole", "stream"])
def test_outputs_match_reference(name, cls):
    g = load_golden(name)
    sd, imgs, proj, dv2, dv3, meta = rebuild_msred_case(g)
    if cls == "whole":
        out = MO.cascade_rednet_forward(sd, imgs, proj, dv3, ndepths=meta["ndepths"], ratios=(4.0, 2.0, 1.0))
    else:
        out = MO.infer_cascade_rednet_forward(sd, imgs, proj, dv2, num_depth=meta["num_depth"],
                                              ndepths=meta["ndepths"], ratios=(4.0, 2.0, 1.0))
    H, W, B = meta["H"], meta["W"], meta["B"]
    for i, s in enumerate(("stage1", "stage2", "stage3")):
        assert tuple(out[s]["depth"].shape) == (B, H >> (2 - i), W >> (2 - i))      # H/4, H/2, H: no x2 output layer
        assert rel_err(out[s]["depth"], g[f"{cls}_{s}_depth"]) < 3e-6, s
        assert abs_err(out[s]["photometric_confidence"], g[f"{cls}_{s}_conf"]) < 3e-6, s
    assert out["depth"] is out["stage3"]["depth"]


def test_intermediates_match_reference():
    g = load_golden("msred_small_d8")
    sd, imgs, proj, dv2, dv3, meta = rebuild_msred_case(g)
    cap = {}
    MO.cascade_rednet_forward(sd, imgs, proj, dv3, ndepths=meta["ndepths"], ratios=(4.0, 2.0, 1.0), capture=cap)
    for i in range(3):
        want = g[f"whole_s{i + 1}_variance"]
        assert abs_err(cap[f"stage{i + 1}"]["variance"], want) < 2e-6 * float(np.abs(want).max())
        want = g[f"whole_s{i + 1}_logits"]
        assert abs_err(cap[f"stage{i + 1}"]["logits"], want) < 1e-5 * max(1.0, float(np.abs(want).max()))


def test_train_class_first_stage_range_quirk_is_preserved():
    """CascadeREDNet hands [min,max,interval] to the first stage, whose planes therefore run from min to
    `interval` (reference msrednet.py:308, module.py:651-653).  The golden depth of stage 1 must lie in that
    (absurd) range, which shows the fixtures exercise the quirk and the oracle reproduces it."""
    g = load_golden("msred_small_d8")
    d = g["whole_stage1_depth"]
    assert float(d.max()) <= 520.0 + 1e-3 and float(d.min()) >= 5.0 - 1e-3


def test_state_dict_keys_match_drop_in_module():
    from adamvs_b200 import synth
    import models.msrednet as M
    m = M.Infer_CascadeREDNet(num_depth=32, ndepths=[8, 4, 2])
    want = synth.msred_state_dict_shapes()
    have = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert have == want
    assert len(have) == 219                       # the reference's key count
