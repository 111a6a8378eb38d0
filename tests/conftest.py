import os
import shutil
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_listed():
    """True when the machine has an NVIDIA GPU according to nvidia-smi (independent of this process' CUDA state)."""
    smi = shutil.which("nvidia-smi")
    if not smi:
        return False
    try:
        out = subprocess.run([smi, "-L"], capture_output=True, text=True, timeout=30).stdout
    except Exception:
        return False
    return "GPU " in out


def _wait_for_cuda(tries=10, pause=3.0):
    """On a fresh GPU box the first CUDA initialisation has been seen to fail transiently ("CUDA driver initialization
    failed") - and a failed init is sticky inside a process, which would silently SKIP every gpu test.  Probe in child
    processes until CUDA comes up, before this process touches it."""
    probe = [sys.executable, "-c", "import torch, sys; sys.exit(0 if torch.cuda.is_available() else 1)"]
    for _ in range(tries):
        try:
            if subprocess.run(probe, capture_output=True, timeout=180).returncode == 0:
                return True
        except Exception:
            pass
        time.sleep(pause)
    return False


def pytest_collection_modifyitems(config, items):
    if not any("gpu" in item.keywords for item in items):
        return
    if _gpu_listed() and not _wait_for_cuda():
        pytest.exit("nvidia-smi lists a GPU but CUDA does not initialise: refusing to silently skip the gpu tests", returncode=3)
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
