"""The reference's OWN predict_whu.py, executed unmodified (tools/run_reference_script.py, from the git-ignored copy
baseline/_ref) on a tiny WHU-OMVS style scene: with the reference's models (CPU, the launcher itself) and, on the GPU,
with this repo's drop-in ``models/`` first on sys.path — "drops into predict_whu.py unchanged" (SURVEY.md §8b)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from adamvs_b200 import sceneio as S, synth
from tests.helpers import abs_err, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join(ROOT, "tests", "golden", "io_scene")
HAVE_REF = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "predict_whu.py"))
NDEPTHS, NUM_DEPTH, VIEWS, MAX_H, MAX_W = (8, 4, 2), 32, 3, 64, 96

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref absent (python baseline/install_ref.py in the build container)")


def _state_dict():
    return synth.fill_state_dict(synth.state_dict_shapes(NDEPTHS[0]), 5)


def _run_script(tmp, models, extra_env=None, cpu_shim=False, model="adamvs", sd=None):
    from tools.run_reference_script import make_checkpoint
    ck = os.path.join(tmp, f"model_{model}.ckpt")
    make_checkpoint(ck, _state_dict() if sd is None else sd)
    out = os.path.join(tmp, f"out_{model}_{models}")
    os.makedirs(out, exist_ok=True)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "--script", "predict_whu.py", "--models", models]
    if cpu_shim:
        cmd.append("--cpu-shim")
    cmd += ["--", "--model", model, "--data_folder", SCENE, "--output_folder", out, "--loadckpt", ck,
            "--view_num", str(VIEWS), "--numdepth", str(NUM_DEPTH), "--ndepths", ",".join(map(str, NDEPTHS)),
            "--max_h", str(MAX_H), "--max_w", str(MAX_W), "--resize_scale", "1"]
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "final, total_cnt = 3" in r.stdout, r.stdout[-2000:]
    return out, r.stdout


def _outputs(out_dir):
    found = {}
    for view in sorted(os.listdir(out_dir)):
        for f in sorted(os.listdir(os.path.join(out_dir, view))):
            if f.endswith("_init.pfm"):
                stem = f[: -len("_init.pfm")]
                d, _ = S.read_pfm(os.path.join(out_dir, view, f))
                p, _ = S.read_pfm(os.path.join(out_dir, view, stem + "_prob.pfm"))
                found[f"{view}/{stem}"] = (d.copy(), p.copy(), open(os.path.join(out_dir, view, stem + ".txt")).read())
    return found


@needs_ref
def test_reference_script_with_reference_models_matches_oracle_on_cpu(tmp_path):
    """The launcher, the scene fixture and the checkpoint layout, with nothing of ours on the model path: the unmodified
    script + unmodified models on CPU write PFMs equal to the oracle's forward on inputs prepared by OUR scene reader."""
    from oracle import adamvs_oracle as O
    from PIL import Image
    out, _ = _run_script(str(tmp_path), "reference", cpu_shim=True)
    got = _outputs(out)
    assert len(got) == 3
    sd = _state_dict()
    cams, poses = S.parse_camera_info(os.path.join(SCENE, "camera_info.txt")), S.parse_image_info(os.path.join(SCENE, "image_info.txt"))
    paths, _ = S.parse_image_paths(os.path.join(SCENE, "image_path.txt"))
    rows = S.parse_view_pairs(os.path.join(SCENE, "viewpair.txt"), VIEWS)
    for row in rows:
        images = [np.array(Image.open(os.path.join(SCENE, paths[i]))) for i in row[:VIEWS]]
        imgs, proj, dv, _, _ = S.load_view_sample(row, poses, cams, images, VIEWS, num_depth=NUM_DEPTH, max_h=MAX_H, max_w=MAX_W,
                                                  device=torch.device("cpu"))
        want = O.infer_adamvs_forward(sd, imgs[None], {k: torch.from_numpy(v)[None] for k, v in proj.items()},
                                      torch.from_numpy(dv)[None], num_depth=NUM_DEPTH, ndepths=NDEPTHS)
        name = poses[row[0]].name
        key = f"{os.path.dirname(name).split('/')[-1]}/{os.path.splitext(os.path.basename(name))[0]}"
        d, p, _ = got[key]
        assert rel_err(torch.from_numpy(d), want["depth"][0]) < 1e-4, key
        assert abs_err(torch.from_numpy(p), want["photometric_confidence"][0]) < 1e-4, key


@needs_ref
@pytest.mark.gpu
def test_reference_predict_script_runs_unchanged_on_the_dropin(tmp_path):
    """predict_whu.py (unmodified) with this repo's models/ first on sys.path: DataParallel wrap, 'module.'-prefixed
    checkpoint, DataLoader, tocuda/tensor2numpy and the save code all run as written, and the PFMs it writes equal
    (a) what the SAME script writes with the reference's own models on this GPU in true fp32, within the north_star's
    tolerance, and (b) what adamvs_b200.pipeline.predict_scene writes; the camera files are byte-identical."""
    from adamvs_b200 import pipeline
    from models.adamvs import Infer_AdaMVSNet
    ours_dir, log = _run_script(str(tmp_path), "ours")
    assert os.path.join(ROOT, "models") in log                     # the launcher reports which `models` package was imported
    ref_dir, _ = _run_script(str(tmp_path), "reference", extra_env={"NVIDIA_TF32_OVERRIDE": "0"})
    ours, ref = _outputs(ours_dir), _outputs(ref_dir)
    assert sorted(ours) == sorted(ref) and len(ours) == 3
    for k in ours:
        assert rel_err(torch.from_numpy(ours[k][0]), torch.from_numpy(ref[k][0])) < 1e-4, k
        assert abs_err(torch.from_numpy(ours[k][1]), torch.from_numpy(ref[k][1])) < 1e-4, k
        assert ours[k][2] == ref[k][2], k
    m = Infer_AdaMVSNet(num_depth=NUM_DEPTH, ndepths=list(NDEPTHS), depth_intervals_ratio=[4.0, 2.0, 1.0])
    m.load_state_dict(_state_dict())
    mine_dir = os.path.join(str(tmp_path), "out_predict_scene")
    pipeline.predict_scene(m.cuda().eval(), SCENE, mine_dir, view_num=VIEWS, num_depth=NUM_DEPTH, max_h=MAX_H, max_w=MAX_W,
                           batch=2, device=torch.device("cuda:0"))
    mine = _outputs(mine_dir)
    assert sorted(mine) == sorted(ours)
    for k in ours:                                                   # device-side vs numpy image normalisation: 1e-6 on the inputs
        assert rel_err(torch.from_numpy(mine[k][0]), torch.from_numpy(ours[k][0])) < 1e-5, k
        assert abs_err(torch.from_numpy(mine[k][1]), torch.from_numpy(ours[k][1])) < 1e-5, k
        assert mine[k][2] == ours[k][2], k


@needs_ref
@pytest.mark.gpu
def test_reference_predict_script_runs_unchanged_on_the_msrednet_dropin(tmp_path):
    """The same for `--model msrednet` (BASELINE configs[4]'s class, Infer_CascadeREDNet): the unmodified predict_whu.py
    with this repo's models/msrednet.py against the same script with the reference's own module on this GPU in true fp32.
    MS-REDNet's probabilities carry the reference's own arithmetic noise (5e-4 at these sizes), hence 2e-3."""
    sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), 41)
    ours_dir, log = _run_script(str(tmp_path), "ours", model="msrednet", sd=sd)
    assert os.path.join(ROOT, "models") in log
    ref_dir, _ = _run_script(str(tmp_path), "reference", extra_env={"NVIDIA_TF32_OVERRIDE": "0"}, model="msrednet", sd=sd)
    ours, ref = _outputs(ours_dir), _outputs(ref_dir)
    assert sorted(ours) == sorted(ref) and len(ours) == 3
    for k in ours:
        assert rel_err(torch.from_numpy(ours[k][0]), torch.from_numpy(ref[k][0])) < 1e-4, k
        assert abs_err(torch.from_numpy(ours[k][1]), torch.from_numpy(ref[k][1])) < 2e-3, k
        assert ours[k][2] == ref[k][2], k


@needs_ref
@pytest.mark.gpu
def test_reference_train_script_runs_unchanged_on_the_dropin(tmp_path):
    """train_whu.py --mode train (unmodified, from baseline/_ref) with this repo's models/ first on sys.path: one epoch
    over a two-sample BlendedMVS-layout scene.  Its own loop does model.train() -> forward -> cas_mvs_vis_loss ->
    backward -> RMSprop step (train_whu.py:265-300), saves checkpoints, then evaluates with model.eval() under no_grad
    (the inference kernels) and writes train_record.txt.  Checked: finite losses, a checkpoint whose weights moved
    and stayed finite, the evaluation record."""
    import re
    from tests.helpers import make_blendedmvs_scene
    from tools.run_reference_script import make_checkpoint
    root = make_blendedmvs_scene(str(tmp_path / "train"))
    sd = _state_dict()
    ck = str(tmp_path / "init.ckpt")
    make_checkpoint(ck, sd)
    logdir = str(tmp_path / "log")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "--script", "train_whu.py", "--models", "ours", "--",
           "--mode", "train", "--model", "adamvs", "--set_name", "BlendedMVS", "--dataset", "cas_total_rscv",
           "--trainpath", root, "--testpath", root, "--loadckpt", ck, "--logdir", logdir, "--view_num", "5",
           "--ndepths", ",".join(map(str, NDEPTHS)), "--epochs", "1", "--summary_freq", "1", "--batch_size", "1", "--lr", "0.0005"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    losses = [float(x) for x in re.findall(r"train loss = ([0-9.eE+-]+|nan|inf)", r.stdout)]
    assert len(losses) == 2 and all(np.isfinite(l) and l > 0 for l in losses), r.stdout[-2000:]
    tests_ = re.findall(r"test loss = ([0-9.eE+-]+|nan|inf)", r.stdout)
    assert len(tests_) == 2 and all(np.isfinite(float(x)) for x in tests_), r.stdout[-2000:]
    assert os.path.isfile(os.path.join(logdir, "train_record.txt"))
    saved = torch.load(os.path.join(logdir, "model_000000.ckpt"), map_location="cpu")
    moved, n = 0, 0
    for k, v in saved["model"].items():
        assert torch.isfinite(v.float()).all(), k
        k0 = k[len("module."):]
        if v.dtype.is_floating_point and k0 in sd and "running" not in k0:
            n += 1
            moved += int(not torch.equal(v, sd[k0]))
    # every parameter the forward uses received a gradient and an RMSprop step; DepthNet.1/2.reg are never executed
    assert moved >= 120 and moved < n, (moved, n)
