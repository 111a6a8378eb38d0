#!/usr/bin/env python
"""Run one of the reference's OWN scripts (predict_whu.py / train_whu.py, unmodified, from the git-ignored copy
``baseline/_ref`` made by ``baseline/install_ref.py``) against either this repo's drop-in ``models/`` package or the
reference's own ``models/`` — the proof of "drops into train_whu.py and predict_whu.py unchanged" (SURVEY.md §8b).

    python tools/run_reference_script.py --script predict_whu.py --models ours -- --model adamvs --data_folder ... \
        --output_folder ... --loadckpt ... --view_num 3 --numdepth 32 --ndepths 8,4,2 --max_h 64 --max_w 96 --resize_scale 1

What the launcher adds, and nothing else (the script file itself is executed byte for byte with runpy):
  * ``sys.path``: ``--models ours`` puts this repo's root (our ``models/``) in front of ``baseline/_ref`` (the
    reference's ``datasets/`` and ``utils.py``); ``--models reference`` leaves only ``baseline/_ref``;
  * stub modules for packages this image lacks and the scripts import at the top: ``imageio`` (names only),
    ``matplotlib.pyplot`` (``imsave`` writes the file with PIL), ``tensorboardX.SummaryWriter`` (no-op);
  * ``np.float = float`` (removed in NumPy 2; datasets/predict_oblique.py:83 uses it);
  * the working directory is the scene folder (``image_path.txt`` holds paths relative to it);
  * ``--cpu-shim``: ``Tensor.cuda`` / ``Module.cuda`` become the identity so that the REFERENCE models can be driven in
    the GPU-less build container (our models refuse CPU tensors by design).
Also: ``make_checkpoint(path, state_dict)`` writes a checkpoint in the layout the scripts load
(``{'model': {'module.<key>': tensor}}``, predict_whu.py:86-88).
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def make_checkpoint(path, state_dict, epoch=0):
    import torch
    torch.save({"epoch": epoch, "model": {"module." + k: v for k, v in state_dict.items()}}, path)


def install_stubs():
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float
    if "imageio" not in sys.modules:
        try:
            import imageio  # noqa: F401
        except ImportError:
            sys.modules["imageio"] = types.SimpleNamespace(imread=None, imsave=None, imwrite=None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        def imsave(path, arr, format="png", **kw):
            from PIL import Image
            a = np.asarray(arr)
            if a.dtype != np.uint8:
                a = np.nan_to_num(a.astype(np.float64))
                lo, hi = float(a.min()), float(a.max())
                a = ((a - lo) / (hi - lo + 1e-12) * 255).astype(np.uint8)
            Image.fromarray(a).save(path, format="PNG")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.imsave = imsave
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    try:
        import tensorboardX  # noqa: F401
    except ImportError:
        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None
        tb = types.ModuleType("tensorboardX")
        tb.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = tb


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default="predict_whu.py", choices=["predict_whu.py", "train_whu.py"])
    ap.add_argument("--models", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cwd", default=None, help="working directory for the script (default: its --data_folder / --trainpath)")
    ap.add_argument("--cpu-shim", action="store_true")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    rest = a.rest[1:] if a.rest[:1] == ["--"] else a.rest
    script = os.path.join(REF, a.script)
    if not os.path.isfile(script):
        raise SystemExit(f"{script} is missing: run `python baseline/install_ref.py` in the build container")
    install_stubs()
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        del sys.modules[name]
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, REF)]
    sys.path.insert(0, REF)
    if a.models == "ours":
        sys.path.insert(0, ROOT)
    if a.cpu_shim:
        import torch
        torch.Tensor.cuda = lambda self, *x, **k: self
        torch.nn.Module.cuda = lambda self, *x, **k: self
    cwd = a.cwd
    if cwd is None:
        for flag in ("--data_folder", "--trainpath"):
            if flag in rest:
                cwd = rest[rest.index(flag) + 1]
    if cwd:
        os.chdir(cwd)
    import models.adamvs as M                       # which package did `models` resolve to?
    want = ROOT if a.models == "ours" else REF
    assert os.path.abspath(M.__file__).startswith(want), (M.__file__, want)
    print(f"run_reference_script: {a.script} with models from {os.path.dirname(M.__file__)}", flush=True)
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
