"""Drop-in for /root/reference/src/models/data_loader.py: Blender `transforms_{split}.json` + PNG frames ->
(images (N,3,H,W), poses (N,4,4), (H, W, focal)).  Disk I/O that runs once per run (outside the hot path);
kept so that the trainer shell finds the loader under the reference's name.  Same semantics as
data_loader.py:8-64, including the focal length being scaled by focal_scale on top of the resized width
(data_loader.py:40,62)."""
import json
import os

import numpy as np
import torch


def load_blender_data(basedir, split="train", img_size=None, half_res=False):
    from PIL import Image
    with open(os.path.join(basedir, "transforms_%s.json" % split), "r") as f:
        meta = json.load(f)
    images, poses = [], []
    for frame in meta["frames"]:
        img_path = os.path.join(basedir, frame["file_path"] + ".png")
        if not os.path.exists(img_path):
            raise FileNotFoundError("Image not found: %s" % img_path)
        img = Image.open(img_path).convert("RGB")
        w_orig, h_orig = img.size
        if img_size:
            dims, focal_scale = (img_size, img_size), img_size / w_orig
        elif half_res:
            dims, focal_scale = (h_orig // 2, w_orig // 2), 0.5
        else:
            dims, focal_scale = (h_orig, w_orig), 1.0
        img = img.resize((dims[1], dims[0]), Image.LANCZOS)          # T.Resize((h, w)) -> PIL size (w, h)
        images.append(torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1))
        poses.append(torch.from_numpy(np.array(frame["transform_matrix"], dtype=np.float32)))
    images = torch.stack(images)
    poses = torch.stack(poses)
    _, _, H, W = images.shape
    focal = 0.5 * W / np.tan(0.5 * meta["camera_angle_x"]) * focal_scale
    return images, poses, (H, W, focal)
