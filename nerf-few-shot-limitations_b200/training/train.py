"""Repaired drop-in for /root/reference/src/training/train.py.

The reference's NeRFDINOTrainer does not run as shipped.  This module keeps its
class name, method names, config schema, progressive schedule, loss, optimiser / scheduler semantics and
checkpoint keys (train.py:46-406) and applies the minimal repairs:
  B1/B2  models.nerf_model.NeRFMLP accepts train.py's keyword form and returns (rgb, density);
  B3     models.ray_sampler.sample_points_along_rays accepts (N,3) rays;
  B4     near / far fall back config['near'] -> config['rendering'] -> config['data'];
  B5     utils.ray_utils.get_rays builds its grid on the pose's device;
  B6     torchmetrics / lpips / imageio / wandb are absent offline: PSNR = -10 log10(mse), SSIM from a Gaussian
         window in torch, LPIPS not computed (NaN), PNGs through PIL when present, no wandb.
Beyond the reference (which is single-device and launches every op eagerly): each optimisation step is captured
once per batch shape into a CUDA graph and replayed (training.cuda_graph, default on), and under torchrun the
batch of every step is sharded over the ranks (one process per GPU; every rank walks the same permutation, the
weight gradient is summed in the fused peer-memory Adam step of nfs_b200.dist or by NCCL).
Everything per ray runs on the CUDA kernels of this package (there is no CPU path); the Dinov2 feature
extractor (per-view preprocessing, needs downloaded weights) is outside the hot path: with use_dino the
trainer takes precomputed (1, Hp, Wp, C) feature maps through set_feature_maps().
"""
import math
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from models.data_loader import load_blender_data  # noqa: E402
from models.nerf_mlp import VolumeRenderer  # noqa: E402
from models.nerf_model import NeRFMLP  # noqa: E402
from models.ray_sampler import sample_points_along_rays  # noqa: E402
from nfs_b200 import dist as _nd  # noqa: E402
from nfs_b200 import ops as _ops  # noqa: E402
from nfs_b200 import pipeline as _pipeline  # noqa: E402
from nfs_b200.optim import FusedAdam  # noqa: E402
from utils.ray_utils import get_rays  # noqa: E402


class NeRFLoss(nn.Module):
    """train.py:27-44: only the weighted rgb MSE is returned (depth / regularisation weights are stored)."""

    def __init__(self, rgb_weight=1.0, depth_weight=0.1, reg_weight=0.01):
        super().__init__()
        self.rgb_weight, self.depth_weight, self.reg_weight = rgb_weight, depth_weight, reg_weight

    def forward(self, predictions, targets):
        return {"rgb": self.rgb_weight * F.mse_loss(predictions["rgb"], targets["rgb"])}


def near_far(config):
    """B4: train.py:192-193 reads config['near'] / ['far'], which only lora.yaml / multiscale.yaml define."""
    out = []
    for key in ("near", "far"):
        for scope in (config, config.get("rendering", {}), config.get("data", {})):
            if key in scope:
                out.append(float(scope[key]))
                break
        else:
            raise KeyError("config has no '%s' (looked at top level, rendering.*, data.*)" % key)
    return tuple(out)


def multistep_lr(base_lr, milestones, gamma, epoch):
    """optim.lr_scheduler.MultiStepLR (train.py:120-124) after `epoch` scheduler steps."""
    return base_lr * gamma ** sum(1 for m in milestones if m <= epoch)


def ssim(img, ref, window=11, sigma=1.5):
    """Mean SSIM of (1,3,H,W) images in [0,1] (Gaussian 11x11 window, the torchmetrics default)."""
    k = torch.arange(window, dtype=torch.float32, device=img.device) - window // 2
    g = torch.exp(-(k ** 2) / (2 * sigma ** 2))
    g = (g / g.sum())[:, None] * (g / g.sum())[None, :]
    w = g.expand(3, 1, window, window).contiguous()
    mu_x, mu_y = F.conv2d(img, w, groups=3), F.conv2d(ref, w, groups=3)
    sxx = F.conv2d(img * img, w, groups=3) - mu_x ** 2
    syy = F.conv2d(ref * ref, w, groups=3) - mu_y ** 2
    sxy = F.conv2d(img * ref, w, groups=3) - mu_x * mu_y
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return float((((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x ** 2 + mu_y ** 2 + c1) * (sxx + syy + c2))).mean())


def synthetic_scene(n_views, resolution, seed=0):
    """A procedural stand-in for data/nerf_synthetic/<scene> (no dataset on the box): cameras on the Blender
    sphere (r = 4.0311, phi = -30 deg) looking at the origin, images of a shaded unit sphere on black (what
    Image.convert('RGB') makes of the Blender PNGs' transparent background, data_loader.py:35)."""
    g = torch.Generator().manual_seed(seed)
    H = W = resolution
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    images, poses = [], []
    for v in range(n_views):
        th, ph = math.radians(float(torch.rand((), generator=g)) * 360.0 - 180.0), math.radians(-30.0)
        t = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0311], [0, 0, 0, 1]], dtype=torch.float32)
        rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0],
                           [0, 0, 0, 1]], dtype=torch.float32)
        rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0],
                           [0, 0, 0, 1]], dtype=torch.float32)
        flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
        c2w = flip @ rt @ rp @ t
        i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="xy")
        dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
        rd = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
        rd = rd / rd.norm(dim=-1, keepdim=True)
        ro = c2w[:3, 3]
        b = (rd * ro).sum(-1)
        disc = b * b - (ro.dot(ro) - 1.0)
        hit = disc > 0
        tt = -b - torch.sqrt(disc.clamp_min(0))
        n = ro + rd * tt[..., None]
        shade = (0.25 + 0.75 * (n @ torch.tensor([0.4, 0.5, 0.76])).clamp_min(0))[..., None]
        col = shade * (0.5 + 0.5 * n)
        img = torch.where(hit[..., None], col, torch.zeros_like(col))
        images.append(img.permute(2, 0, 1))
        poses.append(c2w)
    return torch.stack(images), torch.stack(poses), (H, W, focal)


class NeRFDINOTrainer:
    """train.py:46-389 with the repairs listed in the module docstring."""

    def __init__(self, config, device=None):
        self.config = config
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("NeRFDINOTrainer: the B200 path needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda")
        self.device = torch.device(device)
        self.use_dino = config["model"].get("use_dino", True)
        self.dino_model = None                      # the Dinov2 extractor is per-view preprocessing (out of scope)
        self.dino_features_precomputed = None
        dino_dim = int(config.get("dino_model", {}).get("feature_dim", 64)) if self.use_dino else 0
        nc = config["nerf_model"]
        self.nerf_model = NeRFMLP(pos_freq=nc["pos_freq"], dir_freq=nc["dir_freq"], hidden_dim=nc["hidden_dim"],
                                  num_density_layers=nc["num_layers"], use_dino=self.use_dino,
                                  dino_dim=dino_dim).to(self.device)                      # train.py:82-89
        self.volume_renderer = VolumeRenderer().to(self.device)
        lc = config["loss"]
        self.criterion = NeRFLoss(rgb_weight=lc["rgb_weight"], depth_weight=lc["depth_weight"], reg_weight=lc["reg_weight"])
        oc = config["optimizer"]
        self.base_lr = float(oc["lr"])
        self.optimizer = FusedAdam(self.nerf_model.parameters(), lr=self.base_lr,
                                   weight_decay=float(oc.get("weight_decay", 0.0)))        # optim.Adam, train.py:114
        self.milestones, self.gamma = list(oc["lr_milestones"]), float(oc["lr_gamma"])
        self.sched_epoch = 0
        self.near, self.far = near_far(config)
        self.epoch = 0
        self.best_psnr = 0.0
        # data parallel: one process per GPU under torchrun, rays sharded per batch
        self.rank, self.world = 0, 1
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.exchange = _nd.make_allreduce(self.optimizer) if self.world > 1 else None
        self.use_graph = bool(config["training"].get("cuda_graph", True))
        self._graphs = {}
        self._perm_gen = torch.Generator(device=self.device)
        self._perm_gen.manual_seed(int(config.get("experiment", {}).get("seed", 0)))      # same permutations on every rank

    # ------------------------------------------------------------------ data
    def load_data(self, data_path, split="train", max_views=None):
        images, poses, hwf = load_blender_data(data_path, split=split, img_size=self.config["data"]["resolution"])
        self._install(images, poses, hwf, split, max_views)

    def load_synthetic(self, n_train=None, n_test=2, seed=0):
        res = self.config["data"]["resolution"]
        n_train = n_train or self.config["data"]["num_views"]
        self._install(*synthetic_scene(n_train, res, seed), "train", None)
        self._install(*synthetic_scene(n_test, res, seed + 1), "test", None)

    def _install(self, images, poses, hwf, split, max_views):
        if split == "train":
            if max_views is not None:
                images, poses = images[:max_views], poses[:max_views]
            self.H, self.W, self.focal = hwf
            self.images = [img.permute(1, 2, 0).float().to(self.device) for img in images]
            self.poses = [p.float().to(self.device) for p in poses]
            self.pose_invs = [torch.inverse(p).contiguous() for p in self.poses]
        else:
            self.test_images = [img.permute(1, 2, 0).float().to(self.device) for img in images]
            self.test_poses = [p.float().to(self.device) for p in poses]

    def set_feature_maps(self, maps):
        """Precomputed per-view feature maps (1, Hp, Wp, C) (what precompute_dino_features, train.py:158-169,
        would produce)."""
        self.dino_features_precomputed = [m.float().to(self.device) for m in maps]

    def get_rays_for_view(self, view_idx, split="train"):
        pose = self.poses[view_idx] if split == "train" else self.test_poses[view_idx]
        target = self.images[view_idx] if split == "train" else self.test_images[view_idx]
        rays_o, rays_d = get_rays(self.H, self.W, self.focal, pose)
        if target.shape[-1] == 4:
            target = target[..., :3] * target[..., 3:4] + (1.0 - target[..., 3:4])
        return rays_o, rays_d, target

    # ------------------------------------------------------------------ hot path (train.py:188-242)
    def render_rays(self, rays_o, rays_d, view_idx, N_samples=64):
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        pts, z_vals = sample_points_along_rays(rays_o, rays_d, self.near, self.far, N_samples,
                                               perturb=self.nerf_model.training)
        n_rays = rays_o.shape[0]
        pts_flat = pts.reshape(-1, 3)
        feats = None
        dirs = rays_d.unsqueeze(1).expand(-1, N_samples, -1).reshape(-1, 3)
        plan = self.nerf_model._get_plan() if hasattr(self.nerf_model, "_get_plan") else None
        if self.use_dino:
            if self.dino_features_precomputed is None:
                raise RuntimeError("use_dino: call set_feature_maps() with the per-view (1,Hp,Wp,C) feature maps")
            idx = view_idx if self.nerf_model.training else 0
            if hasattr(plan, "chain_a_bwd") and plan.D > 0 and os.environ.get("NFS_G3_OPERAND", "1") != "0":
                # projection + feature lookup + encoding as the producer of the model's first operand (nfs_g3_operand)
                from nfs_b200 import mlp_g3
                rgb, density = mlp_g3.g3_forward_from_map(plan, pts_flat, dirs, self.dino_features_precomputed[idx],
                                                          self.pose_invs[idx], self.focal, self.H, self.W)
                feats = False
            else:
                _, _, _, feats = _ops.project_gather(pts_flat, self.poses[idx], self.focal, self.H, self.W,
                                                     features=self.dino_features_precomputed[idx], want_projection=False,
                                                     pose_inv=self.pose_invs[idx])
        if feats is not False:
            rgb, density = self.nerf_model(pts_flat, dirs, feats)
        rgb_r, depth_r, weights = self.volume_renderer(rgb.reshape(n_rays, N_samples, 3),
                                                       density.reshape(n_rays, N_samples, 1), z_vals, rays_d)
        return {"rgb": rgb_r, "depth": depth_r, "weights": weights}

    def train_step(self, epoch):
        self.nerf_model.train()
        sched = self.config["training"]["progressive_schedule"]
        bs = self.config["training"]["batch_size"]
        if epoch < 50:
            (h_t, w_t, n_samples), batch = sched["epochs_0_50"], bs * 2
        elif epoch < 100:
            (h_t, w_t, n_samples), batch = sched["epochs_50_100"], bs
        else:
            (h_t, w_t, n_samples), batch = sched["epochs_100_plus"], bs // 2
        total = torch.zeros((), device=self.device)
        n_batches = 0
        for view_idx in range(len(self.images)):
            # train.py:262-278 builds the full (H,W,3) ray images of the view and gathers a randperm slice of them
            # per batch; nfs_rays_generate evaluates the same rays for the batch's pixels only and gathers the
            # target colours in the same launch (bit-identical values, no (H*W,3) round trips).
            target = self.images[view_idx]
            if target.shape[-1] == 4:
                target = target[..., :3] * target[..., 3:4] + (1.0 - target[..., 3:4])
            focal_t = self.focal
            if h_t != self.H or w_t != self.W:
                focal_t = self.focal * (h_t / self.H)
                target = F.interpolate(target.permute(2, 0, 1).unsqueeze(0), size=(h_t, w_t), mode="bilinear",
                                       align_corners=False).squeeze(0).permute(1, 2, 0)
            target = target.contiguous()
            perm = torch.randperm(h_t * w_t, device=self.device, generator=self._perm_gen)
            for i in range(0, perm.shape[0], batch):
                n_global = min(batch, perm.shape[0] - i)
                b = _nd.shard_batch(perm, i, batch, self.rank, self.world)       # this rank's rays of the global batch
                loss = self._optimise(b, n_global, view_idx, h_t, w_t, focal_t, target, n_samples)
                total += loss                       # one host sync per epoch instead of one per batch (train.py:289)
                n_batches += 1
        return float(total) / n_batches if n_batches else 0.0

    # ------------------------------------------------------------------ one optimisation step (train.py:280-287)
    def _batch_loss(self, idx, view_idx, h_t, w_t, focal_t, target, n_samples, pose=None):
        ro_b, rd_b, tg_b = _ops.generate_rays(h_t, w_t, focal_t, self.poses[view_idx] if pose is None else pose,
                                              pix_idx=idx, image=target)
        pred = self.render_rays(ro_b, rd_b, view_idx, n_samples)
        return sum(self.criterion(pred, {"rgb": tg_b}).values())

    def _optimise(self, idx, n_global, view_idx, h_t, w_t, focal_t, target, n_samples):
        """zero_grad / loss / backward / (gradient exchange) / optimizer.step for this rank's `idx` pixels of a global
        batch of n_global; the mean over the global batch is sum over ranks of local_mean * local/global.  Every batch
        shape (the full batches' and the tail's of a permutation) replays a CUDA graph captured once; training.cuda_graph
        = false and the feature-conditioned model launch eagerly.
        Returns the detached local loss (a device tensor)."""
        n_local = int(idx.numel())
        scale = _nd.loss_scale(n_local, n_global)
        ex = self.exchange
        fused = bool(getattr(ex, "in_graph", False))
        graphable = self.use_graph and n_local > 0 and not self.use_dino      # conditioned model: per-view maps vary in shape
        if graphable:
            key = (h_t, w_t, n_samples, n_local, n_global)
            g = self._graphs.get(key)
            if g is None:
                st = {"idx": idx.clone(), "pose": self.poses[view_idx].clone(), "target": target.clone(), "view": -1}
                closure = lambda: self._batch_loss(st["idx"], 0, h_t, w_t, focal_t, st["target"], n_samples, pose=st["pose"])
                st["step"] = _pipeline.GraphedStep(self.optimizer, closure, loss_scale=scale, allreduce=ex)
                g = self._graphs[key] = st
            g["idx"].copy_(idx)
            if g["view"] != (view_idx, self.epoch):
                g["pose"].copy_(self.poses[view_idx])
                g["target"].copy_(target)
                g["view"] = (view_idx, self.epoch)
            return g["step"].replay().clone()
        opt = self.optimizer
        opt.zero_grad()
        if fused:
            ex.wait_readers()
        if n_local > 0:
            loss = self._batch_loss(idx, view_idx, h_t, w_t, focal_t, target, n_samples)
            (loss * scale if scale != 1.0 else loss).backward()
            out = loss.detach()
        else:
            out = torch.zeros((), device=self.device)
        opt.gather_grads()
        if fused:
            ex.fused_step()
        else:
            if ex is not None:
                ex(opt.grad)
            opt.step(gathered=True)
        return out

    @torch.no_grad()
    def evaluate(self, epoch):
        self.nerf_model.eval()
        psnrs, ssims = [], []
        out_dir = os.path.join(self.config["output"]["save_dir"], "epoch_%d" % epoch)
        n_eval = self.config["training"]["progressive_schedule"]["epochs_100_plus"][2]
        chunk = self.config["rendering"]["chunk_size"]
        for i in range(len(self.test_images)):
            rays_o, rays_d, target = self.get_rays_for_view(i, "test")
            ro, rd = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
            parts = [self.render_rays(ro[j:j + chunk], rd[j:j + chunk], 0, n_eval)["rgb"] for j in range(0, ro.shape[0], chunk)]
            img = torch.cat(parts, 0).reshape(self.H, self.W, 3)
            # metrics on the UNCLAMPED rendering, as train.py:321-326 computes them (data_range 1.0); only the PNG is clamped
            mse = F.mse_loss(img, target)
            psnrs.append(float(-10.0 * torch.log10(mse.clamp_min(1e-12))))
            ssims.append(ssim(img.permute(2, 0, 1).unsqueeze(0), target.permute(2, 0, 1).unsqueeze(0)))
            if i < 5 and self.rank == 0:
                try:
                    from PIL import Image
                    os.makedirs(out_dir, exist_ok=True)
                    Image.fromarray((img.clamp(0, 1).cpu().numpy() * 255).astype("uint8")).save(
                        os.path.join(out_dir, "render_%d.png" % i))
                except ImportError:
                    pass
        return {"psnr": sum(psnrs) / len(psnrs), "ssim": sum(ssims) / len(ssims), "lpips": float("nan")}

    def train(self, epochs):
        for epoch in range(self.epoch, epochs):
            self.epoch = epoch
            loss = self.train_step(epoch)
            self.sched_epoch += 1                                                          # scheduler.step()
            self.optimizer.lr = multistep_lr(self.base_lr, self.milestones, self.gamma, self.sched_epoch)
            if self.rank == 0:
                print("Epoch %d/%d | Train Loss: %.4f | LR: %.2e" % (epoch + 1, epochs, loss, self.optimizer.lr))
            if (epoch + 1) % self.config["output"]["val_freq"] == 0:
                m = self.evaluate(epoch)
                if self.rank == 0:
                    print("Validation PSNR: %.2f, SSIM: %.2f" % (m["psnr"], m["ssim"]))
                if m["psnr"] > self.best_psnr:
                    self.best_psnr = m["psnr"]
                    if self.rank == 0:
                        self.save_checkpoint("best_%s.pth" % self.config["experiment"]["name"])
            if (epoch + 1) % self.config["output"]["save_freq"] == 0 and self.rank == 0:
                self.save_checkpoint("epoch_%d.pth" % (epoch + 1))
        if self.rank == 0:
            print("Training completed. Best PSNR: %.2f" % self.best_psnr)

    def save_checkpoint(self, filename):
        """Same keys as train.py:374-389 (no 'dino_model_state_dict': the extractor is not part of this trainer).
        optimizer_state_dict / scheduler_state_dict have the layouts of torch.optim.Adam.state_dict() and
        MultiStepLR.state_dict(), so a checkpoint loads into the reference's optimizer / scheduler and vice versa."""
        from collections import Counter
        ckpt = {"epoch": self.epoch, "best_psnr": self.best_psnr,
                "nerf_model_state_dict": self.nerf_model.state_dict(),
                "optimizer_state_dict": self.optimizer.state_dict(torch_format=True),
                "scheduler_state_dict": {"milestones": Counter(self.milestones), "gamma": self.gamma,
                                         "base_lrs": [self.base_lr], "last_epoch": self.sched_epoch,
                                         "_step_count": self.sched_epoch + 1, "_get_lr_called_within_step": False,
                                         "_last_lr": [self.optimizer.lr]},
                "config": self.config}
        path = os.path.join(self.config["output"]["save_dir"], filename)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(ckpt, path)
        return path

    def load_checkpoint(self, path):
        """Resume from a checkpoint written by save_checkpoint() or by the reference's trainer (train.py:374-389)."""
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.nerf_model.load_state_dict(ckpt["nerf_model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        sd = ckpt.get("scheduler_state_dict", {})
        self.sched_epoch = int(sd.get("last_epoch", 0))
        self.optimizer.set_lr(multistep_lr(self.base_lr, self.milestones, self.gamma, self.sched_epoch))
        self.epoch = int(ckpt.get("epoch", -1)) + 1
        self.best_psnr = float(ckpt.get("best_psnr", 0.0))
        return ckpt


def main():
    import argparse
    import yaml
    ap = argparse.ArgumentParser(description="Train a NeRF model on the B200 render path.")
    ap.add_argument("--config", type=str, required=True, help="path to an experiments/*.yaml of the reference")
    ap.add_argument("--synthetic", action="store_true", help="procedural scene instead of data/<dataset>/<scene>")
    ap.add_argument("--epochs", type=int, default=None)
    args = ap.parse_args()
    with open(args.config) as f:
        config = yaml.safe_load(f)
    rank, world, local_rank = _nd.world()
    if world > 1:                                    # torchrun --nproc-per-node G: one process per GPU
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    trainer = NeRFDINOTrainer(config, device=torch.device("cuda", local_rank) if world > 1 else None)
    if args.synthetic:
        trainer.load_synthetic()
    else:
        root = os.path.join("data", config["data"]["dataset"], config["data"]["scene"])
        trainer.load_data(root, "train", max_views=config["data"]["num_views"])
        trainer.load_data(root, "test")
    trainer.train(args.epochs or config["training"]["epochs"])


if __name__ == "__main__":
    main()
