"""Wire formats either side of the hot path: the WHU-OMVS scene text files that
`predict_whu.py` reads and the depth / probability / camera files it writes, plus the per-view preprocessing
between them and the network (camera conversion, scale, crop, per-image normalisation, the 3-level projection
pyramid).  Same bytes and same numbers as the reference's helpers, written from their behaviour:

    reference                                              here
    datasets/data_io.py:47-72   read_cameras_text          parse_camera_info
    datasets/data_io.py:75-101  read_images_text           parse_image_info
    datasets/data_io.py:104-118 read_images_path_text      parse_image_paths
    datasets/data_io.py:121-136 read_view_pair_text        parse_view_pairs
    datasets/predict_oblique.py:72-112 create_cams         camera_block
    datasets/preprocess.py:22-83 scale_camera / crop_input scale_camera_block / crop_to_network
    datasets/predict_oblique.py:59-64, preprocess.py:101-111 center_image   center_images (torch, any device)
    datasets/predict_oblique.py:150-177 projection pyramid  projection_pyramid
    datasets/data_io.py:160-190 read_pfm                   read_pfm
    datasets/data_io.py:193-222 save_pfm                   write_pfm / pfm_bytes
    datasets/data_io.py:139-157 write_red_cam              write_cam_txt

Host code (numpy / torch); nothing here touches the CUDA library.  `tests/test_sceneio.py` holds it to fixtures made
by the reference's own functions (`tests/golden/make_io_golden.py`)."""
from __future__ import annotations

import math
import os
import re
import sys
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


# ---------------------------------------------------------------------------------------------------------
# scene text files
# ---------------------------------------------------------------------------------------------------------
@dataclass
class CameraModel:
    camera_id: int
    width: int
    height: int
    pixelsize: float
    fx: float
    fy: float
    cx: float
    cy: float
    distortion: np.ndarray        # k1 k2 k3 p1 p2 (possibly empty)


@dataclass
class ImagePose:
    image_id: int
    camera_id: int
    rotation_wc: np.ndarray       # [3,3] float64, X-right-Y-up camera axes
    centre_wc: np.ndarray         # [3] float64
    depth_min: float
    depth_max: float
    name: str


def _records(path: str):
    """Non-empty, non-comment lines split on whitespace."""
    with open(path, "r") as fh:
        for raw in fh:
            line = raw.strip()
            if line and not line.startswith("#"):
                yield line.split()


def parse_camera_info(path: str) -> Dict[int, CameraModel]:
    """`CAMERA_ID WIDTH HEIGHT PIXELSIZE fx fy cx cy [k1 k2 k3 p1 p2]` per line."""
    out = {}
    for f in _records(path):
        cid = int(f[0])
        out[cid] = CameraModel(cid, int(f[1]), int(f[2]), float(f[3]), float(f[4]), float(f[5]), float(f[6]), float(f[7]),
                               np.asarray([float(v) for v in f[8:]], dtype=np.float64))
    return out


def parse_image_info(path: str) -> Dict[int, ImagePose]:
    """`IMAGE_ID CAMERA_ID Rwc[9] twc[3] MINDEPTH MAXDEPTH NAME` per line."""
    out = {}
    for f in _records(path):
        iid = int(f[0])
        vals = [float(v) for v in f[2:16]]
        out[iid] = ImagePose(iid, int(f[1]), np.asarray(vals[0:9], dtype=np.float64).reshape(3, 3),
                             np.asarray(vals[9:12], dtype=np.float64), vals[12], vals[13], f[16])
    return out


def parse_image_paths(path: str) -> Tuple[Dict[int, str], Dict[int, str]]:
    """A count followed by `index name path` triples (any whitespace): ({index: path}, {index: name})."""
    tok = open(path).read().split()
    n = int(tok[0])
    paths, names = {}, {}
    for i in range(n):
        idx, name, p = int(tok[1 + 3 * i]), tok[2 + 3 * i], tok[3 + 3 * i]
        paths[idx], names[idx] = p, name
    return paths, names


def parse_view_pairs(path: str, view_num: int) -> List[List[int]]:
    """`viewpair.txt`: a count, then per reference view its id line and a `n id score id score ...` line.  Views
    without sources are dropped; short source lists are padded with their first entry (the dataset then takes the
    first `view_num` entries of each row)."""
    rows = []
    with open(path) as fh:
        n = int(fh.readline())
        for _ in range(n):
            ref = int(fh.readline().rstrip())
            src = [int(v) for v in fh.readline().rstrip().split()[1::2]]
            if not src:
                continue
            if len(src) < view_num:
                src = src + [src[0]] * (view_num - len(src))
            rows.append([ref] + src)
    return rows


# ---------------------------------------------------------------------------------------------------------
# cameras
# ---------------------------------------------------------------------------------------------------------
def camera_block(pose: ImagePose, cam: CameraModel, num_depth: int = 384) -> np.ndarray:
    """The reference's `[2,4,4]` float32 camera: `[0]` = Tcw of the X-right-Y-down camera (inverse of `[Rwc·diag(1,-1,-1) | twc]`,
    inverted in float32 as the reference does), `[1][:3,:3]` = K, `[1][3]` = (depth_min, interval, num_depth, depth_max)."""
    blk = np.zeros((2, 4, 4), dtype=np.float32)
    twc = np.zeros((4, 4), dtype=np.float32)
    flip = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)
    twc[0:3, 0:3] = np.matmul(pose.rotation_wc, flip)
    twc[0:3, 3] = pose.centre_wc
    twc[3, 3] = 1.0
    blk[0] = np.linalg.inv(twc)
    blk[1][0][0], blk[1][1][1], blk[1][0][2], blk[1][1][2], blk[1][2][2] = cam.fx, cam.fy, cam.cx, cam.cy, 1
    blk[1][3][0] = pose.depth_min
    blk[1][3][1] = (pose.depth_max - pose.depth_min) / num_depth
    blk[1][3][2] = num_depth
    blk[1][3][3] = pose.depth_max
    return blk


def scale_camera_block(blk: np.ndarray, scale: float = 1) -> np.ndarray:
    out = np.copy(blk)
    for r, c in ((0, 0), (1, 1), (0, 2), (1, 2)):
        out[1][r][c] = blk[1][r][c] * scale
    return out


def crop_to_network(image: np.ndarray, blk: np.ndarray, max_h: int = 384, max_w: int = 768, resize_scale: float = 1,
                    multiple: int = 32) -> Tuple[np.ndarray, np.ndarray]:
    """Top-left crop to at most (max_h, max_w)·resize_scale; smaller images get the next multiple of 32 as their
    target size (the slice then keeps them whole).  The principal point moves with the crop origin (0, 0)."""
    max_h, max_w = int(max_h * resize_scale), int(max_w * resize_scale)
    h, w = image.shape[0:2]
    new_h = max_h if h > max_h else int(math.ceil(h / multiple) * multiple)
    new_w = max_w if w > max_w else int(math.ceil(w / multiple) * multiple)
    return image[0:new_h, 0:new_w], blk           # crop origin is (0, 0): the camera is unchanged


def projection_matrix(blk: np.ndarray) -> np.ndarray:
    """`K·[R|t]` in the top three rows of the 4x4 extrinsic (float32)."""
    p = blk[0].copy()
    p[:3, :4] = np.matmul(blk[1, 0:3, 0:3], p[:3, :4])
    return p


def projection_pyramid(proj: np.ndarray) -> Dict[str, np.ndarray]:
    """Per-view full-resolution projections `[V,4,4]` -> the model's dict: rows 0-1 divided by 4 / 2 / 1."""
    s1, s2 = proj.copy(), proj.copy()
    s2[:, :2, :] = proj[:, :2, :] / 2
    s1[:, :2, :] = proj[:, :2, :] / 4
    return {"stage1": s1, "stage2": s2, "stage3": proj}


def center_images(imgs: torch.Tensor) -> torch.Tensor:
    """Per image and channel `(x - mean) / (sqrt(var) + 1e-8)` over H x W with the population variance
    (`np.var`): imgs `[..., H, W, 3]` uint8/float -> float32 `[..., 3, H, W]` on the same device."""
    x = imgs.to(torch.float32)
    mean = x.mean(dim=(-3, -2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(-3, -2), keepdim=True)
    y = (x - mean) / (var.sqrt() + 0.00000001)
    return y.movedim(-1, -3).contiguous()


# ---------------------------------------------------------------------------------------------------------
# outputs
# ---------------------------------------------------------------------------------------------------------
def pfm_bytes(image: np.ndarray, scale: float = 1) -> bytes:
    """`Pf` / `PF` header, `W H`, `%f` scale (negative = little endian), rows bottom-up."""
    if image.dtype.name != "float32":
        raise ValueError("PFM data must be float32")
    if image.ndim == 3 and image.shape[2] == 3:
        tag = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        tag = b"Pf\n"
    else:
        raise ValueError("PFM data must be H x W, H x W x 1 or H x W x 3")
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little")
    head = tag + ("%d %d\n" % (image.shape[1], image.shape[0])).encode() + ("%f\n" % (-scale if little else scale)).encode()
    return head + np.ascontiguousarray(image[::-1]).tobytes()


def write_pfm(path: str, image: np.ndarray, scale: float = 1) -> None:
    with open(path, "wb") as fh:
        fh.write(pfm_bytes(image, scale))


def read_pfm(path: str) -> Tuple[np.ndarray, float]:
    with open(path, "rb") as fh:
        tag = fh.readline().decode("utf-8").rstrip()
        if tag not in ("PF", "Pf"):
            raise ValueError("not a PFM file")
        m = re.match(r"^(\d+)\s(\d+)\s$", fh.readline().decode("utf-8"))
        if not m:
            raise ValueError("malformed PFM header")
        w, h = int(m.group(1)), int(m.group(2))
        scale = float(fh.readline().rstrip())
        data = np.frombuffer(fh.read(), dtype=("<f4" if scale < 0 else ">f4"))
    shape = (h, w, 3) if tag == "PF" else (h, w)
    return np.flipud(data.reshape(shape)), abs(scale)


def write_cam_txt(path: str, blk: np.ndarray, ref_path: str) -> None:
    """The `*.txt` camera next to every depth map: extrinsic rows, intrinsic rows, the depth line, the image path."""
    lines = ["extrinsic: XrightYdown, [Rcw|tcw]\n"]
    for i in range(4):
        lines.append("".join(str(blk[0][i][j]) + " " for j in range(4)) + "\n")
    lines.append("\n")
    lines.append("intrinsic\n")
    for i in range(3):
        lines.append("".join(str(blk[1][i][j]) + " " for j in range(3)) + "\n")
    lines.append("\n" + " ".join(str(blk[1][3][j]) for j in range(4)) + "\n")
    lines.append("\n")
    lines.append(str(ref_path) + "\n")
    with open(path, "w") as fh:
        fh.writelines(lines)


def load_view_sample(rows: Sequence[int], poses: Dict[int, ImagePose], cams: Dict[int, CameraModel], images: Sequence[np.ndarray],
                     view_num: int, num_depth: int = 384, max_h: int = 384, max_w: int = 768, device=None):
    """One reference view with its sources, as `predict_oblique.MVSDataset.__getitem__` assembles it (resize_scale and
    sample_scale 1): returns (imgs `[V,3,H,W]` float32 on `device`, proj pyramid dict of `[V,4,4]`, depth_values `[2]`,
    the reference view's cropped image and camera block)."""
    projs, crops, out_img, out_blk = [], [], None, None
    for v in range(view_num):
        blk = camera_block(poses[rows[v]], cams[poses[rows[v]].camera_id], num_depth)
        img, blk = crop_to_network(np.asarray(images[v]), blk, max_h, max_w)
        if v == 0:
            out_img, out_blk = img, blk
        projs.append(projection_matrix(blk))
        crops.append(torch.from_numpy(np.ascontiguousarray(img)))
    imgs = center_images(torch.stack(crops).to(device) if device is not None else torch.stack(crops))
    depth_values = np.array([out_blk[1][3][0], out_blk[1][3][3]], dtype=np.float32)
    return imgs, projection_pyramid(np.stack(projs)), depth_values, out_img, out_blk
