"""Host -> device input staging for batched inference (the predict_whu.py scenario): the next batch of
views is copied from pinned host memory on a side stream while the current batch computes, so the
H2D transfer (17.7 MB per 5-view 768x384 sample) does not serialise with the kernels.  The reference
copies synchronously inside its loop (utils.tocuda, predict_whu.py:100-104)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

Batch = Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]


class InputPrefetcher:
    def __init__(self, device: torch.device):
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self._staged: Optional[Batch] = None
        self._ready: Optional[torch.cuda.Event] = None

    def stage(self, host_batch: Batch) -> None:
        """Start copying (imgs, proj_matrices, depth_values) — pinned host tensors — to the device."""
        imgs, proj, dv = host_batch
        with torch.cuda.stream(self.copy_stream):
            dev = (imgs.to(self.device, non_blocking=True),
                   {k: v.to(self.device, non_blocking=True) for k, v in proj.items()},
                   dv.to(self.device, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._staged, self._ready = dev, ev

    def take(self) -> Batch:
        """Device tensors of the staged batch; the caller's current stream waits for the copy."""
        if self._staged is None:
            raise RuntimeError("take() without a staged batch")
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready)
        imgs, proj, dv = self._staged
        for t in (imgs, dv, *proj.values()):
            t.record_stream(cur)                      # allocated on the copy stream, consumed on this one
        self._staged = self._ready = None
        return imgs, proj, dv
