"""Scene wire formats and per-view preprocessing against fixtures produced by the reference's own
dataset / data_io code (tests/golden/make_io_golden.py, run in the build container)."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from adamvs_b200 import sceneio as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENE = os.path.join(GOLD, "io_scene")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "io_golden.npz")))


def test_text_files_parse_like_the_reference(gold):
    cams = S.parse_camera_info(os.path.join(SCENE, "camera_info.txt"))
    assert sorted(cams) == list(gold["cam_ids"])
    got = np.array([[cams[c].width, cams[c].height, cams[c].pixelsize, cams[c].fx, cams[c].fy, cams[c].cx, cams[c].cy] for c in sorted(cams)])
    assert np.array_equal(got, gold["cam_params"])
    assert np.array_equal(cams[0].distortion, gold["cam0_distortion"]) and cams[1].distortion.size == 0
    poses = S.parse_image_info(os.path.join(SCENE, "image_info.txt"))
    assert np.array_equal(np.stack([poses[i].rotation_wc for i in sorted(poses)]), gold["pose_R"])
    assert np.array_equal(np.stack([poses[i].centre_wc for i in sorted(poses)]), gold["pose_t"])
    assert np.array_equal(np.array([[poses[i].depth_min, poses[i].depth_max] for i in sorted(poses)]), gold["pose_depth"])
    assert poses[2].name == "area1/view2/img_002.png" and poses[3].camera_id == 1
    paths, names = S.parse_image_paths(os.path.join(SCENE, "image_path.txt"))
    assert paths[1] == os.path.join("images", "img_001.png") and names[3] == "img_003.png" and len(paths) == 4
    # view 2 has no sources and is dropped, view 1's single source is repeated, view 0 keeps all three
    rows = S.parse_view_pairs(os.path.join(SCENE, "viewpair.txt"), 3)
    assert np.array_equal(np.array(rows), gold["sample_rows"])


def test_samples_equal_the_reference_dataset(gold):
    cams = S.parse_camera_info(os.path.join(SCENE, "camera_info.txt"))
    poses = S.parse_image_info(os.path.join(SCENE, "image_info.txt"))
    paths, _ = S.parse_image_paths(os.path.join(SCENE, "image_path.txt"))
    rows = S.parse_view_pairs(os.path.join(SCENE, "viewpair.txt"), 3)
    assert len(rows) == int(gold["n_samples"])
    for k, row in enumerate(rows):
        images = [np.array(Image.open(os.path.join(SCENE, paths[i]))) for i in row[:3]]
        imgs, proj, dv, out_img, out_blk = S.load_view_sample(row, poses, cams, images, 3, num_depth=192, max_h=64, max_w=96)
        for st in ("stage1", "stage2", "stage3"):
            assert np.array_equal(proj[st], gold[f"s{k}_proj_{st}"]), (k, st)       # same float32 operations: bit equal
        assert np.array_equal(dv, gold[f"s{k}_depth_values"])
        assert np.array_equal(out_blk, gold[f"s{k}_outcam"]) and np.array_equal(out_img, gold[f"s{k}_outimage"])
        assert tuple(imgs.shape) == (3, 3, 64, 96) and imgs.dtype == torch.float32
        # fp32 mean / variance over 6144 values of 0..255: numpy's pairwise sums and torch's differ by a few 1e-6 relative
        assert float(np.abs(imgs.numpy() - gold[f"s{k}_imgs"]).max()) < 2e-5


def test_small_images_are_kept_whole_by_the_crop_rule():
    img = np.zeros((70, 100, 3), np.uint8)
    blk = np.zeros((2, 4, 4), np.float32)
    out, _ = S.crop_to_network(img, blk, max_h=384, max_w=768)
    assert out.shape[:2] == (70, 100)                                   # target 96 x 128 exceeds the image: slice is a no-op
    out, _ = S.crop_to_network(img, blk, max_h=64, max_w=96)
    assert out.shape[:2] == (64, 96)


def test_pfm_and_camera_files_are_byte_identical_to_the_reference(gold, tmp_path):
    depth = gold["pfm_depth"]
    ref_bytes = open(os.path.join(GOLD, "io_ref_depth.pfm"), "rb").read()
    assert S.pfm_bytes(depth) == ref_bytes
    p = tmp_path / "d.pfm"
    S.write_pfm(str(p), depth)
    back, scale = S.read_pfm(str(p))
    assert scale == 1.0 and np.array_equal(back, depth)
    rgb = np.random.default_rng(0).random((5, 7, 3), dtype=np.float32)
    S.write_pfm(str(p), rgb, scale=2)
    back, scale = S.read_pfm(str(p))
    assert scale == 2.0 and np.array_equal(back, rgb)
    with pytest.raises(ValueError):
        S.pfm_bytes(depth.astype(np.float64))
    c = tmp_path / "c.txt"
    S.write_cam_txt(str(c), gold["s0_outcam"], "images/img_000.png")
    assert c.read_text() == open(os.path.join(GOLD, "io_ref_cam.txt")).read()


def test_center_images_matches_numpy_definition():
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(2, 24, 40, 3), dtype=np.uint8)
    got = S.center_images(torch.from_numpy(x)).numpy()
    f = x.astype(np.float32)
    want = (f - f.mean(axis=(1, 2), keepdims=True)) / (np.sqrt(f.var(axis=(1, 2), keepdims=True)) + 0.00000001)
    assert got.shape == (2, 3, 24, 40)
    assert float(np.abs(got - want.transpose(0, 3, 1, 2)).max()) < 2e-5
