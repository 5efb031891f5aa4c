"""Blender-synthetic scene reader under the reference's module name (models.data_loader.load_blender_data,
/root/reference/src/models/data_loader.py:8-64): `transforms_{split}.json` + PNG frames ->
(images (N,3,H,W) in [0,1], poses (N,4,4), (H, W, focal)).

Disk I/O that runs once per run, outside the hot path; it exists so that the trainer shell (training/train.py)
finds a loader.  Behaviour kept from the reference: `img_size` wins over `half_res`; the focal length is derived
from the RESIZED width and then multiplied by the resize factor once more (data_loader.py:40,62 - a quirk, kept so
that the same configs give the same cameras)."""
import json
import os

import numpy as np
import torch


def _plan_resize(width, height, img_size, half_res):
    """-> ((new_h, new_w), focal factor)."""
    if img_size:
        return (img_size, img_size), img_size / width
    if half_res:
        return (height // 2, width // 2), 0.5
    return (height, width), 1.0


def _read_rgb(path, img_size, half_res):
    from PIL import Image
    if not os.path.exists(path):
        raise FileNotFoundError("Image not found: %s" % path)
    with Image.open(path) as im:
        im = im.convert("RGB")
        (new_h, new_w), factor = _plan_resize(im.size[0], im.size[1], img_size, half_res)
        if (new_w, new_h) != im.size:
            im = im.resize((new_w, new_h), Image.LANCZOS)
        chw = np.ascontiguousarray(np.asarray(im, dtype=np.float32).transpose(2, 0, 1)) / np.float32(255.0)
    return torch.from_numpy(chw), factor


def load_blender_data(basedir, split="train", img_size=None, half_res=False):
    with open(os.path.join(basedir, "transforms_%s.json" % split)) as fh:
        scene = json.load(fh)
    frames = scene["frames"]
    loaded = [_read_rgb(os.path.join(basedir, fr["file_path"] + ".png"), img_size, half_res) for fr in frames]
    images = torch.stack([img for img, _ in loaded])
    poses = torch.from_numpy(np.stack([np.asarray(fr["transform_matrix"], dtype=np.float32) for fr in frames]))
    factor = loaded[-1][1]
    H, W = int(images.shape[2]), int(images.shape[3])
    focal = 0.5 * W / np.tan(0.5 * scene["camera_angle_x"]) * factor
    return images, poses, (H, W, focal)
