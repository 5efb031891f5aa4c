"""CPU oracle for the NeRF render hot path  ---  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  Nothing under nerf-few-shot-limitations_b200/ imports it, and the product
path has no CPU fallback.

What it is: a restatement, on ATen CPU tensors, of the reference's algorithm for every row
of SURVEY.md section 8a.  The reference (ANKITSANJYAL/nerf-few-shot-limitations) is pure
Python over PyTorch; its arithmetic *is* the sequence of ATen ops it issues, so the oracle
issues the same ops in the same order (that is what makes stratified z-values, searchsorted
indices and bins bit-comparable).  The only third-party arithmetic is PyTorch itself
(requirements.txt:5 `torch>=2.0.0`, unpinned; this image's 2.11.0+cu128 is the de-facto pin).

Pinning: the reference ships no tests and no golden vectors, so the
oracle is pinned against outputs of the reference ITSELF, imported from /root/reference/src
in the build container by tests/golden/make_golden.py; the resulting fixtures are committed
under tests/golden/ and tests/test_oracle_golden.py checks every function below against them.

Each function cites the reference lines it follows (paths relative to the reference root).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- R1 / R2
def _alpha_and_weights(sigma, z_vals, rays_d):
    """sigma, z_vals (...,S); rays_d (...,3) -> weights (...,S).
    src/models/nerf_mlp.py:181-202 == src/models/volume_renderer.py:21-38."""
    gaps = z_vals[..., 1:] - z_vals[..., :-1]                                  # nerf_mlp.py:181
    gaps = torch.cat([gaps, torch.full_like(gaps[..., :1], 1e10)], dim=-1)     # :182
    gaps = gaps * torch.norm(rays_d[..., None, :], dim=-1)                     # :185
    alpha = 1.0 - torch.exp(-F.relu(sigma) * gaps)                             # :193
    survive = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], dim=-1)
    trans = torch.cumprod(survive, dim=-1)[..., :-1]                           # :196-199
    return alpha * trans                                                       # :202


def render(rgb, density, z_vals, rays_d, noise=None, noise_std=0.0, white_bkgd=False):
    """VolumeRenderer.forward, src/models/nerf_mlp.py:165-215.
    rgb (N,S,3), density (N,S,1), z_vals (N,S), rays_d (N,3).  `noise` is the N(0,1) tensor
    the reference would draw with randn_like(density) (:189) - passed in so the CUDA path
    can be given the same draws."""
    if noise is not None and noise_std > 0.0:
        density = density + noise * noise_std                                  # :189-190
    w = _alpha_and_weights(density[..., 0], z_vals, rays_d)
    rgb_out = torch.sum(w[..., None] * rgb, dim=-2)                            # :205
    depth_out = torch.sum(w * z_vals, dim=-1)                                  # :208
    if white_bkgd:
        acc = torch.sum(w[..., None], dim=-2)                                  # :212
        rgb_out = rgb_out + (1.0 - acc) * 1.0                                  # :213
    return rgb_out, depth_out, w


def render_packed(rgb_sigma, z_vals, rays_d):
    """volume_render_radiance, src/models/volume_renderer.py:4-43 (noise_std = 0 branch;
    the noise branch is an in-place ATen add on the caller's tensor before this)."""
    w = _alpha_and_weights(rgb_sigma[..., 3], z_vals, rays_d)                  # :18,21-38
    return torch.sum(w[..., None] * rgb_sigma[..., :3], dim=-2)                # :41


def render_backward_closed_form(rgb, density, z_vals, rays_d, g_rgb, g_depth=None, g_w=None,
                                white_bkgd=False):
    """Closed form of autograd through render() (SURVEY.md section 8a footnote); used to
    cross-check the formula the CUDA backward implements.  Works in the dtype of the inputs."""
    sig = density[..., 0]
    gaps = z_vals[..., 1:] - z_vals[..., :-1]
    gaps = torch.cat([gaps, torch.full_like(gaps[..., :1], 1e10)], dim=-1)
    gaps = gaps * torch.norm(rays_d[..., None, :], dim=-1)
    e = torch.exp(-F.relu(sig) * gaps)
    alpha = 1.0 - e
    q = 1.0 - alpha + 1e-10
    T = torch.cumprod(torch.cat([torch.ones_like(q[..., :1]), q], dim=-1), dim=-1)[..., :-1]
    w = alpha * T
    G = (g_rgb[..., None, :] * rgb).sum(-1)
    if g_depth is not None:
        G = G + g_depth[..., None] * z_vals
    if g_w is not None:
        G = G + g_w
    if white_bkgd:
        G = G - g_rgb.sum(-1, keepdim=True)
    gw = G * w
    incl = torch.flip(torch.cumsum(torch.flip(gw, [-1]), -1), [-1])             # sum_{k>=i} G_k w_k
    suffix = torch.cat([incl[..., 1:], torch.zeros_like(incl[..., :1])], -1)   # sum_{k>i}: shifted, never subtracted
    d_alpha = G * T - suffix / q
    d_sigma = d_alpha * gaps * e * (sig > 0).to(sig.dtype)
    return w[..., None] * g_rgb[..., None, :], d_sigma[..., None]


# ----------------------------------------------------------------------------- R3 / R4
def frequency_bands(num_freqs, log_sampling=True):
    """src/models/positional_encoding.py:13-18, src/models/nerf_mlp.py:14."""
    if log_sampling:
        return 2.0 ** torch.linspace(0.0, num_freqs - 1, steps=num_freqs)
    return torch.linspace(2.0 ** 0.0, 2.0 ** (num_freqs - 1), steps=num_freqs)


def encode(x, bands, include_input=True):
    """src/models/positional_encoding.py:27-33 (x*f) / src/models/nerf_mlp.py:24-33 (f*x)."""
    parts = [x] if include_input else []
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


# ----------------------------------------------------------------------------- R7 / R8
def stratified(rays_o, rays_d, near, far, n_samples, t_rand=None, lindisp=False):
    """src/utils/ray_utils.py:55-84 and src/models/ray_sampler.py:47-61 (any leading shape).
    t_rand = the uniform draws of ray_utils.py:78 (None <=> perturb=False)."""
    t = torch.linspace(0.0, 1.0, n_samples)                                    # ray_utils.py:60/64
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)                     # :61
    else:
        z = near * (1.0 - t) + far * t                                         # :65
    z = z.expand(*rays_o.shape[:-1], n_samples)                                # :67
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])                                # :72
        upper = torch.cat([mids, z[..., -1:]], -1)                             # :73
        lower = torch.cat([z[..., :1], mids], -1)                              # :74
        z = lower + (upper - lower) * t_rand                                   # :79
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]        # :82
    return pts, z


# ----------------------------------------------------------------------------- R9
def hierarchical(rays_o, rays_d, z_vals, weights, u):
    """src/utils/ray_utils.py:101-143 with the uniform draws `u` (N,Ni) passed in.
    Valid (as in the reference) only for weights.shape[-1] == z_vals.shape[-1] - 1.
    Returns a dict with every intermediate the parity tests look at."""
    n_rays = z_vals.shape[0]
    n_imp = u.shape[-1]
    w = weights + 1e-5                                                         # :105
    pdf = w / torch.sum(w, dim=-1, keepdim=True)                               # :108
    cdf = torch.cumsum(pdf, dim=-1)                                            # :109
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)             # :110
    u = u.expand(n_rays, n_imp).contiguous()                                   # :117,120
    idx = torch.searchsorted(cdf, u, right=True)                               # :121
    below = torch.max(torch.zeros_like(idx - 1), idx - 1)                      # :122
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(idx), idx)         # :123
    pair = torch.stack([below, above], dim=-1)                                 # :124
    shape = [n_rays, n_imp, cdf.shape[-1]]                                     # :127
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, pair)              # :128
    bins_g = torch.gather(z_vals.unsqueeze(1).expand(shape), 2, pair)          # :129
    den = cdf_g[..., 1] - cdf_g[..., 0]                                        # :132
    den = torch.where(den < 1e-5, torch.ones_like(den), den)                   # :133
    t = (u - cdf_g[..., 0]) / den                                              # :134
    samples = bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])           # :135
    z_all, _ = torch.sort(torch.cat([z_vals, samples], dim=-1), dim=-1)        # :138
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]    # :141
    return dict(cdf=cdf, idx=idx, below=below, above=above, cdf_g=cdf_g, bins_g=bins_g,
                samples=samples, z=z_all, pts=pts)


# ----------------------------------------------------------------------------- R5
class PlainNeRF(nn.Module):
    """nerf_model.NeRFMLP, src/models/nerf_model.py:5-24 (G1).  Parameter creation order,
    names and shapes match so that a fixed torch.manual_seed gives the reference's weights."""

    def __init__(self, pos_dim=63, hidden_dim=256, n_layers=8):
        super().__init__()
        self.layers = nn.ModuleList(
            [nn.Linear(pos_dim if i == 0 else hidden_dim, hidden_dim) for i in range(n_layers)])   # :8-11
        self.sigma_out = nn.Linear(hidden_dim, 1)                              # :13
        self.rgb_out = nn.Linear(hidden_dim, 3)                                # :14

    def forward(self, x, dir_enc=None):
        h = x
        for lin in self.layers:
            h = F.relu(lin(h))                                                 # :18-20
        return torch.cat([torch.sigmoid(self.rgb_out(h)), self.sigma_out(h)], dim=-1)   # :22-24


# ----------------------------------------------------------------------------- R6
class _Encoder(nn.Module):
    """nerf_mlp.PositionalEncoding, src/models/nerf_mlp.py:6-39 (bands are a buffer)."""

    def __init__(self, num_freqs):
        super().__init__()
        self.num_freqs = num_freqs
        self.register_buffer("freq_bands", frequency_bands(num_freqs))

    def forward(self, x):
        return encode(x, self.freq_bands, True)


class _Fusion(nn.Module):
    """NeRFDINOFusion, src/models/dino_feature_model.py:150-197 (== lora_dino.py:146-193)."""

    def __init__(self, pos_dim, dino_dim, hidden_dim):
        super().__init__()
        self.fusion = nn.Sequential(nn.Linear(pos_dim + dino_dim, hidden_dim), nn.ReLU(),
                                    nn.Linear(hidden_dim, hidden_dim), nn.ReLU())          # :157-162
        self.attention = nn.Sequential(nn.Linear(hidden_dim, hidden_dim // 4), nn.ReLU(),
                                       nn.Linear(hidden_dim // 4, 2), nn.Softmax(dim=-1))  # :165-170
        self.output_proj = nn.Linear(hidden_dim, hidden_dim)                               # :172

    def forward(self, pe, feat):
        fused = self.fusion(torch.cat([pe, feat], dim=-1))                     # :182-185
        gate = self.attention(fused)                                           # :188
        again = self.fusion(torch.cat([pe * gate[:, 0:1], feat * gate[:, 1:2]], dim=-1))   # :191-195
        return self.output_proj(again)                                         # :197


class _Density(nn.Module):
    """DensityMLP, src/models/nerf_mlp.py:41-66."""

    def __init__(self, input_dim, hidden_dim, num_layers):
        super().__init__()
        mods = []
        for i in range(num_layers):
            mods += [nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim), nn.ReLU()]
        self.density_layers = nn.Sequential(*mods)
        self.density_head = nn.Linear(hidden_dim, 1)
        self.feature_head = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, x):
        h = self.density_layers(x)
        return F.relu(self.density_head(h)), self.feature_head(h)              # :60-66


class _Color(nn.Module):
    """ColorMLP, src/models/nerf_mlp.py:68-84."""

    def __init__(self, feature_dim, dir_dim, hidden_dim):
        super().__init__()
        self.color_layers = nn.Sequential(nn.Linear(feature_dim + dir_dim, hidden_dim), nn.ReLU(),
                                          nn.Linear(hidden_dim, hidden_dim // 2), nn.ReLU(),
                                          nn.Linear(hidden_dim // 2, 3), nn.Sigmoid())

    def forward(self, feat, dirs):
        return self.color_layers(torch.cat([feat, dirs], dim=-1))


class ConditionedNeRF(nn.Module):
    """NeRFWithDINO, src/models/nerf_mlp.py:86-158 (G3).  Sub-module names and creation
    order match the reference so state_dicts and seeded initialisations interchange."""

    def __init__(self, pos_freq=10, dir_freq=4, dino_dim=64, hidden_dim=256, num_density_layers=8):
        super().__init__()
        self.pos_encoder = _Encoder(pos_freq)                                  # :101
        self.dir_encoder = _Encoder(dir_freq)                                  # :102
        pos_dim, dir_dim = 3 + 6 * pos_freq, 3 + 6 * dir_freq                  # :105-106
        self.dino_fusion = _Fusion(pos_dim, dino_dim, hidden_dim)              # :111-115
        self.density_mlp = _Density(hidden_dim, hidden_dim, num_density_layers)   # :118-122
        self.color_mlp = _Color(hidden_dim, dir_dim, hidden_dim // 2)          # :125-129

    def forward(self, positions, directions, dino_features):
        pe = self.pos_encoder(positions)                                       # :146
        de = self.dir_encoder(directions)                                      # :147
        density, feat = self.density_mlp(self.dino_fusion(pe, dino_features))  # :150-153
        return self.color_mlp(feat, de), density                               # :156-158


# ----------------------------------------------------------------------------- R10
def nerf_loss(pred, target, rgb_weight=1.0, depth_weight=0.1, reg_weight=0.01):
    """nerf_mlp.NeRFLoss.forward, src/models/nerf_mlp.py:225-258."""
    out = {"rgb": F.mse_loss(pred["rgb"], target["rgb"])}                      # :235-236
    if "depth" in target:
        out["depth"] = F.l1_loss(pred["depth"], target["depth"])               # :239-241
    if "weights" in pred:
        out["regularization"] = torch.mean(pred["weights"] ** 2)               # :244-246
    total = rgb_weight * out["rgb"]                                            # :249
    if "depth" in out:
        total = total + depth_weight * out["depth"]                            # :251-252
    if "regularization" in out:
        total = total + reg_weight * out["regularization"]                     # :254-255
    out["total"] = total
    return out


# ----------------------------------------------------------------------------- 8f rank 1
def project_points(points_3d, pose, focal, H, W):
    """utils.ray_utils.project_points_to_image, src/utils/ray_utils.py:176-210."""
    pose_inv = torch.inverse(pose)                                                         # :192
    homo = torch.cat([points_3d, torch.ones_like(points_3d[..., :1])], dim=-1)             # :193
    cam = torch.matmul(homo, pose_inv.T)[..., :3]                                          # :194
    valid = cam[..., 2] > 0                                                                # :198
    x = cam[..., 0] / (cam[..., 2] + 1e-8) * focal + W / 2                                 # :201
    y = cam[..., 1] / (cam[..., 2] + 1e-8) * focal + H / 2                                 # :202
    return torch.stack([(x / W) * 2 - 1, (y / H) * 2 - 1], dim=-1), cam[..., 2], valid     # :205-210


def sample_features(features, points_2d):
    """SpatialDINOFeatures.sample_features_at_points, src/models/dino_feature_model.py:114-148:
    features (B,Hp,Wp,C), points_2d (N,2) -> (N,C) for B == 1."""
    B = features.shape[0]
    grid = points_2d.unsqueeze(0).unsqueeze(2).expand(B, -1, 1, -1)                        # :131-132
    out = F.grid_sample(features.permute(0, 3, 1, 2), grid, mode="bilinear", padding_mode="zeros",
                        align_corners=False)                                               # :135-140
    out = out.squeeze(-1).permute(0, 2, 1)                                                 # :143
    return out.squeeze(0) if B == 1 else out                                               # :145-148


# ----------------------------------------------------------------------------- synthetic rays
def pose_spherical(theta_deg, phi_deg, radius):
    """Blender-style camera-to-world looking at the origin."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    t = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=torch.float32)
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0],
                       [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0],
                       [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ rt @ rp @ t


def pixel_rays(H, W, focal, c2w):
    """src/models/ray_sampler.py:18-30: rays_o, rays_d (H,W,3); directions NOT normalised."""
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32),
                          indexing="xy")
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    return c2w[:3, 3].expand(rays_d.shape), rays_d


def ray_batch(H, W, focal, c2w, idx, image=None):
    """The batch assembly of NeRFDINOTrainer.train_step, src/training/train.py:272-278: rays (and target colours)
    of the pixels idx of one view, by gathering from the view's full ray images."""
    ro, rd = pixel_rays(H, W, focal, c2w)
    out = (ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx])                                # :275-276
    return out if image is None else out + (image.reshape(-1, 3)[idx],)                  # :277


def pixel_rays_scalar(H, W, focal, c2w):
    """pixel_rays with the ATen arithmetic spelled out in numpy fp32 (what the CUDA kernel implements): true
    division by the focal length, each product rounded, the three terms added left to right.  Equal to pixel_rays
    bit for bit (tests/test_oracle_golden.py::test_rays)."""
    import numpy as np
    f32 = np.float32
    R = c2w[:3, :3].numpy().astype(f32)
    i, j = np.meshgrid(np.arange(W, dtype=f32), np.arange(H, dtype=f32), indexing="xy")
    dx = (i - f32(W * 0.5)) / f32(focal)
    dy = -((j - f32(H * 0.5)) / f32(focal))
    dz = -np.ones_like(i)
    rd = np.stack([(dx * R[k, 0] + dy * R[k, 1]) + dz * R[k, 2] for k in range(3)], -1)
    return c2w[:3, 3].expand(H, W, 3), torch.from_numpy(rd)


def lego_rays(n_rays, H=800, W=800, seed=0):
    """n_rays rays drawn from one synthetic Blender-lego-shaped view (camera_angle_x =
    0.6911112, r = 4.0311, phi = -30 deg) - the generator of SURVEY.md section 8d."""
    g = torch.Generator().manual_seed(seed)
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    theta = float(torch.rand((), generator=g) * 360.0 - 180.0)
    ro, rd = pixel_rays(H, W, focal, pose_spherical(theta, -30.0, 4.0311))
    pick = torch.randint(0, H * W, (n_rays,), generator=g)
    return ro.reshape(-1, 3)[pick].contiguous(), rd.reshape(-1, 3)[pick].contiguous()
