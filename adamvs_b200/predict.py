"""Command-line front end of `pipeline.predict_scene`: the `predict_whu.py` scenario (reference predict_whu.py:22-90
for the arguments, :92-161 for the loop) with one process per GPU and the reference views sharded over the ranks.

    python -m adamvs_b200.predict --data_folder SCENE --output_folder OUT --loadckpt model.ckpt [--model adamvs|msrednet]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m adamvs_b200.predict ...   # 8 GPUs

Argument names and defaults follow the reference where the option exists there.  `--resize_scale` other than 1 and
`--sample_scale` other than 1 (cv2 resampling of the inputs) are not implemented and are refused."""
from __future__ import annotations

import argparse
import os
import sys
import time


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Predict depth maps for a WHU-OMVS style scene folder on B200 GPUs")
    p.add_argument("--model", default="adamvs", choices=["adamvs", "msrednet"])
    p.add_argument("--data_folder", required=True, help="scene folder with viewpair.txt, image_info.txt, camera_info.txt, image_path.txt")
    p.add_argument("--output_folder", required=True)
    p.add_argument("--loadckpt", default=None, help="reference checkpoint ({'model': state_dict}); random init when omitted")
    p.add_argument("--view_num", type=int, default=5)
    p.add_argument("--numdepth", type=int, default=192)
    p.add_argument("--max_w", type=int, default=3712)
    p.add_argument("--max_h", type=int, default=5504)
    p.add_argument("--resize_scale", type=float, default=1.0)
    p.add_argument("--sample_scale", type=float, default=1.0)
    p.add_argument("--batch_size", type=int, default=8)
    p.add_argument("--share_cr", action="store_true")
    p.add_argument("--ndepths", type=str, default="48,32,8")
    p.add_argument("--depth_inter_r", type=str, default="4,2,1")
    p.add_argument("--cr_base_chs", type=str, default="8,8,8")
    return p


def model_from_args(a):
    """The drop-in module with the reference's constructor arguments (predict_whu.py:66-80)."""
    ndepths = [int(v) for v in a.ndepths.split(",") if v]
    ratios = [float(v) for v in a.depth_inter_r.split(",") if v]
    chs = [int(v) for v in a.cr_base_chs.split(",") if v]
    if a.model == "msrednet":
        from models.msrednet import Infer_CascadeREDNet
        return Infer_CascadeREDNet(num_depth=a.numdepth, ndepths=ndepths, depth_interals_ratio=ratios, share_cr=a.share_cr,
                                   cr_base_chs=chs)
    from models.adamvs import Infer_AdaMVSNet
    return Infer_AdaMVSNet(num_depth=a.numdepth, ndepths=ndepths, depth_intervals_ratio=ratios, share_cr=a.share_cr,
                           cr_base_chs=chs)


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    if a.resize_scale != 1.0 or a.sample_scale != 1.0:
        raise SystemExit("--resize_scale / --sample_scale other than 1 are not implemented (inputs are used at their own size)")
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("adamvs_b200.predict needs a CUDA device (sm_100a); there is no CPU path")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    model = model_from_args(a)
    if a.loadckpt:
        sd = torch.load(a.loadckpt, map_location="cpu")["model"]
        model.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in sd.items()})
    from .pipeline import predict_scene
    t0 = time.time()
    written = predict_scene(model, a.data_folder, a.output_folder, view_num=a.view_num, num_depth=a.numdepth, max_h=a.max_h,
                            max_w=a.max_w, batch=a.batch_size, device=device, rank=rank, world=world)
    torch.cuda.synchronize(device)
    dt = time.time() - t0
    print(f"rank {rank}/{world}: {len(written)} depth maps in {dt:.2f} s ({len(written) / max(dt, 1e-9):.1f} maps/s) -> {a.output_folder}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
