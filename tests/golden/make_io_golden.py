#!/usr/bin/env python
"""Generate the scene-I/O fixtures by running the UNMODIFIED reference dataset / data_io code (/root/reference) on a tiny
synthetic WHU-OMVS style scene.  Build container only; the fixtures are committed.

    python tests/golden/make_io_golden.py

Writes tests/golden/io_scene/{camera_info,image_info,image_path,viewpair}.txt + images/*.png (inputs) and
tests/golden/io_golden.npz (what the reference made of them) + io_ref_depth.pfm / io_ref_cam.txt (its output files).
Accommodations: `imageio` is not installed (stub module; the dataset only imports names from it) and NumPy 2 has no
`np.float` (aliased to float, predict_oblique.py:83)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCENE = os.path.join(HERE, "io_scene")
REF = "/root/reference"


def write_scene():
    rng = np.random.default_rng(42)
    os.makedirs(os.path.join(SCENE, "images"), exist_ok=True)
    from PIL import Image
    H, W, n = 70, 100, 4                      # not multiples of 32: the crop rule must round up, not cut
    with open(os.path.join(SCENE, "camera_info.txt"), "w") as f:
        f.write("# CAMERA_ID WIDTH HEIGHT PIXELSIZE fx fy cx cy k1 k2 k3 p1 p2\n")
        f.write("0 100 70 0.0046 120.5 121.25 49.5 34.75 0.0 0.0 0.0 0.0 0.0\n")
        f.write("\n1 100 70 0.0046 118.0 118.0 50.0 35.0\n")
    with open(os.path.join(SCENE, "image_info.txt"), "w") as f:
        f.write("# IMAGE_ID CAMERA_ID Rwc[9] twc[3] MINDEPTH MAXDEPTH NAME\n")
        for i in range(n):
            a = 0.05 * i
            R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ \
                np.array([[1, 0, 0], [0, np.cos(0.1), -np.sin(0.1)], [0, np.sin(0.1), np.cos(0.1)]])
            t = np.array([431000.5 + 25.0 * i, 3895000.25 - 10.0 * i, 600.0 + i])
            vals = " ".join("%.10f" % v for v in list(R.reshape(-1)) + list(t))
            f.write(f"{i} {i % 2} {vals} {520.0 + i} {680.0 + 2 * i} area1/view{i}/img_{i:03d}.png\n")
    with open(os.path.join(SCENE, "image_path.txt"), "w") as f:
        f.write(f"{n}\n")
        for i in range(n):
            f.write(f"{i} img_{i:03d}.png {os.path.join('images', f'img_{i:03d}.png')}\n")
    with open(os.path.join(SCENE, "viewpair.txt"), "w") as f:
        f.write(f"{n}\n")
        f.write("0\n3 1 0.9 2 0.8 3 0.7\n")
        f.write("1\n1 0 0.9\n")                 # short list: padded with its first entry
        f.write("2\n0\n")                        # no sources: dropped
        f.write("3\n2 2 0.5 0 0.4\n")
    for i in range(n):
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        Image.fromarray(img).save(os.path.join(SCENE, "images", f"img_{i:03d}.png"))


def main():
    write_scene()
    sys.modules.setdefault("imageio", types.SimpleNamespace(imread=None, imsave=None, imwrite=None))
    if not hasattr(np, "float"):
        np.float = float
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(SCENE)                              # image_path.txt holds paths relative to the scene folder
    from datasets.predict_oblique import MVSDataset
    from datasets import data_io
    args = types.SimpleNamespace(min_interval=0.1, interval_scale=1.0, numdepth=192, resize_scale=1, sample_scale=1,
                                 max_h=64, max_w=96)
    ds = MVSDataset(SCENE, 3, args)
    out = {"n_samples": np.array(len(ds)), "sample_rows": np.array(ds.sample_list, dtype=np.int64)}
    for k in range(len(ds)):
        s = ds[k]
        out[f"s{k}_imgs"] = s["imgs"].astype(np.float32)
        for st in ("stage1", "stage2", "stage3"):
            out[f"s{k}_proj_{st}"] = s["proj_matrices"][st]
        out[f"s{k}_depth_values"] = s["depth_values"]
        out[f"s{k}_outcam"] = s["outcam"]
        out[f"s{k}_outimage"] = np.asarray(s["outimage"])
    cams = data_io.read_cameras_text(os.path.join(SCENE, "camera_info.txt"))
    poses = data_io.read_images_text(os.path.join(SCENE, "image_info.txt"))
    out["cam_ids"] = np.array(sorted(cams))
    out["cam_params"] = np.array([[cams[c].size[0], cams[c].size[1], cams[c].pixelsize, *cams[c].focallength, *cams[c].x0y0] for c in sorted(cams)])
    out["cam0_distortion"] = cams[0].distortion
    out["pose_R"] = np.stack([poses[i].rotation_matrix for i in sorted(poses)])
    out["pose_t"] = np.stack([poses[i].project_center for i in sorted(poses)])
    out["pose_depth"] = np.stack([poses[i].depth for i in sorted(poses)])
    # the reference's own output files for one depth map
    depth = (np.arange(64 * 96, dtype=np.float32).reshape(64, 96) * 0.37 + 520.0).astype(np.float32)
    out["pfm_depth"] = depth
    os.chdir(cwd)
    data_io.save_pfm(os.path.join(HERE, "io_ref_depth.pfm"), depth)
    data_io.write_red_cam(os.path.join(HERE, "io_ref_cam.txt"), ds[0]["outcam"] if False else out["s0_outcam"], "images/img_000.png")
    np.savez_compressed(os.path.join(HERE, "io_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "io_golden.npz"), {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
