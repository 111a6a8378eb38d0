"""Import the unmodified reference's model modules from ``baseline/_ref`` (or, in the build container, straight
from /root/reference) WITHOUT colliding with this repo's drop-in ``models`` package: the directory is bound to a
private package name (``refmodels``), so ``from .module import *`` inside the reference still resolves.

Baseline / test infrastructure only — nothing in ``adamvs_b200/`` or ``models/`` imports this.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def reference_root():
    for root in CANDIDATES:
        if os.path.isfile(os.path.join(root, "models", "adamvs.py")):
            return root
    return None


def load(module: str = "adamvs"):
    """-> the reference's ``models/<module>.py`` as a module object; raises FileNotFoundError when not installed."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError("the reference is not installed: run `python baseline/install_ref.py` in the build "
                                "container (copies /root/reference into the git-ignored baseline/_ref)")
    if "refmodels" not in sys.modules:
        pkg = types.ModuleType("refmodels")
        pkg.__path__ = [os.path.join(root, "models")]
        sys.modules["refmodels"] = pkg
    mod = importlib.import_module("refmodels." + module)
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(root)), mod.__file__
    return mod


@contextlib.contextmanager
def cpu_cuda_shim():
    """The reference hard-codes ``.cuda()`` inside its forwards (models/adamvs.py:175-176, 448-459); on a CPU run
    ``Tensor.cuda`` is the identity for the duration of the block."""
    import torch
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = saved
