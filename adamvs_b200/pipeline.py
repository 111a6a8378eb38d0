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


def predict_scene(model, scene_dir: str, out_dir: str, view_num: int = 5, num_depth: int = 384, max_h: int = 384,
                  max_w: int = 768, batch: int = 8, device: Optional[torch.device] = None, rank: int = 0, world: int = 1):
    """The `predict_whu.py` loop on this stack (reference predict_whu.py:92-161 with datasets/predict_oblique.py): read a
    WHU-OMVS scene folder, run the reference views of this rank (views are independent: contiguous shard, no collective)
    in batches through `model`, write `<out_dir>/<view dir>/<name>_init.pfm`, `_prob.pfm` and `<name>.txt` per view.
    Raw uint8 images are uploaded (3x smaller than the normalised float tensors the reference ships) and normalised on
    the device; the next batch is staged on a side stream while the current one computes.  Returns the written paths."""
    import os

    import numpy as np
    from PIL import Image

    from . import sceneio as S
    from .sharding import shard_range

    device = device or torch.device("cuda", torch.cuda.current_device())
    cams = S.parse_camera_info(os.path.join(scene_dir, "camera_info.txt"))
    poses = S.parse_image_info(os.path.join(scene_dir, "image_info.txt"))
    paths, _ = S.parse_image_paths(os.path.join(scene_dir, "image_path.txt"))
    rows = S.parse_view_pairs(os.path.join(scene_dir, "viewpair.txt"), view_num)
    lo, hi = shard_range(len(rows), rank, world)
    rows = rows[lo:hi]

    def host_batch(chunk):
        raw, projs, dvs, meta = [], {"stage1": [], "stage2": [], "stage3": []}, [], []
        for row in chunk:
            blks, crops = [], []
            for v in range(view_num):
                p = paths[row[v]]
                img = np.array(Image.open(p if os.path.isabs(p) else os.path.join(scene_dir, p)))
                blk = S.camera_block(poses[row[v]], cams[poses[row[v]].camera_id], num_depth)
                img, blk = S.crop_to_network(img, blk, max_h, max_w)
                blks.append(blk)
                crops.append(torch.from_numpy(np.ascontiguousarray(img)))
            pyr = S.projection_pyramid(np.stack([S.projection_matrix(b) for b in blks]))
            for k in projs:
                projs[k].append(torch.from_numpy(pyr[k]))
            if any(c.shape != crops[0].shape for c in crops) or (raw and crops[0].shape != raw[0].shape[1:]):
                raise ValueError("predict_scene: every cropped view of a batch must have one size (got "
                                 f"{sorted({tuple(c.shape) for c in crops})}); use batch=1 for scenes with mixed image sizes")
            raw.append(torch.stack(crops))
            dvs.append(torch.tensor([blks[0][1][3][0], blks[0][1][3][3]], dtype=torch.float32))
            meta.append((poses[row[0]].name, blks[0], paths[row[0]]))
        pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
        return (pin(torch.stack(raw)), {k: pin(torch.stack(v)) for k, v in projs.items()}, pin(torch.stack(dvs))), meta

    written = []
    chunks = [rows[i:i + batch] for i in range(0, len(rows), batch)]
    if not chunks:
        return written
    with torch.cuda.device(device):                   # the C library launches on the current device (ops._guard)
        pre = InputPrefetcher(device)
        nxt, nxt_meta = host_batch(chunks[0])
        pre.stage(nxt)
        model = model.to(device).eval()
        for ci in range(len(chunks)):
            raw, proj, dv = pre.take()
            meta = nxt_meta
            with torch.no_grad():
                out = model(S.center_images(raw), proj, dv)       # asynchronous: the kernels are only enqueued here
            if ci + 1 < len(chunks):
                # decode / crop / pin the next batch on the host WHILE the GPU computes this one, then start its copy
                nxt, nxt_meta = host_batch(chunks[ci + 1])
                pre.stage(nxt)
            depth = out["depth"].cpu().numpy()                    # first host sync of the iteration
            prob = out["photometric_confidence"].cpu().numpy()
            for j, (name, blk, ref_path) in enumerate(meta):
                stem = os.path.splitext(os.path.basename(name))[0]
                view_dir = os.path.join(out_dir, os.path.dirname(name).split("/")[-1])
                os.makedirs(view_dir, exist_ok=True)
                S.write_pfm(os.path.join(view_dir, stem + "_init.pfm"), np.ascontiguousarray(depth[j], dtype=np.float32))
                S.write_pfm(os.path.join(view_dir, stem + "_prob.pfm"), np.ascontiguousarray(prob[j], dtype=np.float32))
                S.write_cam_txt(os.path.join(view_dir, stem + ".txt"), blk, ref_path)
                written.append(os.path.join(view_dir, stem + "_init.pfm"))
    return written
