#!/usr/bin/env python
"""Copy the UNMODIFIED reference (Python/PyTorch sources only) from /root/reference into the git-ignored
``baseline/_ref/`` so that it travels to the GPU box with the repo snapshot.

The reference has no setup.py / pyproject, so there is nothing to pip-install: the "install" is a plain copy of
its ``models/``, ``datasets/``, ``utils.py`` and the two scripts.  Nothing under ``baseline/_ref`` is ever
committed (``.gitignore``) and nothing in ``adamvs_b200/`` or ``models/`` imports it: it is the measured
*baseline* arm of ``bench.py`` (``--impl reference``, ``library_bar``) and the script driven by
``tools/run_reference_script.py`` against this repo's drop-in ``models/``.

    python baseline/install_ref.py            # no-op when /root/reference is absent (the GPU box)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ADAMVS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
WANT = ("models", "datasets", "utils.py", "predict_whu.py", "train_whu.py")


def install(verbose: bool = False) -> bool:
    """-> True when baseline/_ref holds the reference afterwards."""
    if not os.path.isdir(SRC):
        return os.path.isdir(os.path.join(DST, "models"))
    os.makedirs(DST, exist_ok=True)
    for name in WANT:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
        if verbose:
            print("copied", s, "->", d)
    return True


if __name__ == "__main__":
    ok = install(verbose=True)
    print("baseline/_ref", "ready" if ok else "absent (no /root/reference here)")
    sys.exit(0)
