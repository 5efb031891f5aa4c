"""Generate tests/golden/*.pt from the UNMODIFIED reference, imported from /root/reference/src.

Run in the build container only (the reference does not travel to the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Each fixture holds seeded inputs and the reference's outputs (and autograd gradients where
the path is differentiable).  The reference has no tests / golden vectors of its own
(SURVEY.md section 4), so these files are what pins oracle/nerf_oracle.py.
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
REF = "/root/reference/src"
sys.path[:0] = [REF, REF + "/models", REF + "/utils"]

import torch  # noqa: E402

from models.nerf_mlp import NeRFLoss, NeRFWithDINO, VolumeRenderer  # noqa: E402
from models.nerf_mlp import PositionalEncoding as PE4  # noqa: E402
from models.nerf_model import NeRFMLP  # noqa: E402
from models.positional_encoding import PositionalEncoding as PE3  # noqa: E402
from models.ray_sampler import get_rays  # noqa: E402
from models.ray_sampler import sample_points_along_rays as sample_hw  # noqa: E402
from models.volume_renderer import volume_render_radiance  # noqa: E402
from utils import ray_utils  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print("%-28s %8.1f KB" % (name + ".pt", os.path.getsize(path) / 1024))


def rays(n, gen):
    o = torch.tensor([0.5, -3.2, 2.1]).expand(n, 3).contiguous()
    d = torch.randn(n, 3, generator=gen)
    d = d / d.norm(dim=-1, keepdim=True) * (1.0 + 0.2 * torch.rand(n, 1, generator=gen))
    return o, d


def render_cases():
    gen = torch.Generator().manual_seed(101)
    cases = []
    for (n, s, scale, white) in [(48, 64, 10.0, False), (48, 64, 1.0, True), (6, 1, 5.0, False),
                                 (6, 2, 5.0, True), (7, 33, 100.0, False), (5, 192, 10.0, True),
                                 (9, 48, 3.0, False), (4, 128, 30.0, False)]:
        _, rd = rays(n, gen)
        rgb = torch.rand(n, s, 3, generator=gen).requires_grad_()
        den = (torch.randn(n, s, 1, generator=gen) * scale).requires_grad_()
        z = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=gen), dim=-1).values
        t_rgb, t_depth = torch.rand(n, 3, generator=gen), 2.0 + 4.0 * torch.rand(n, generator=gen)
        vr = VolumeRenderer().eval()
        o_rgb, o_depth, o_w = vr(rgb, den, z, rd, white_bkgd=white)
        loss = ((o_rgb - t_rgb) ** 2).mean() + 0.1 * (o_depth - t_depth).abs().mean() + 0.01 * (o_w ** 2).mean()
        g_rgb, g_depth, g_w = torch.autograd.grad(loss, [o_rgb, o_depth, o_w], retain_graph=True)
        d_rgb, d_den = torch.autograd.grad(loss, [rgb, den])
        cases.append(dict(rgb=rgb.detach(), density=den.detach(), z_vals=z, rays_d=rd, white_bkgd=white,
                          out_rgb=o_rgb.detach(), out_depth=o_depth.detach(), out_weights=o_w.detach(),
                          g_rgb=g_rgb, g_depth=g_depth, g_weights=g_w, d_rgb=d_rgb, d_density=d_den))
    # training-mode noise (nerf_mlp.py:188-190): same seed -> same randn_like draw
    n, s = 16, 64
    _, rd = rays(n, gen)
    rgb = torch.rand(n, s, 3, generator=gen)
    den = torch.randn(n, s, 1, generator=gen) * 4.0
    z = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=gen), dim=-1).values
    torch.manual_seed(7)
    o = VolumeRenderer().train()(rgb, den, z, rd, noise_std=1.0)
    torch.manual_seed(7)
    noise = torch.randn_like(den)
    cases.append(dict(rgb=rgb, density=den, z_vals=z, rays_d=rd, white_bkgd=False, noise=noise, noise_std=1.0,
                      out_rgb=o[0], out_depth=o[1], out_weights=o[2]))
    save("render", cases)


def packed_cases():
    gen = torch.Generator().manual_seed(202)
    cases = []
    for shape, s in [((5, 7), 32), ((12,), 64), ((2, 3), 5)]:
        rs = torch.rand(*shape, s, 4, generator=gen)
        rs[..., 3] = torch.randn(*shape, s, generator=gen) * 8.0
        rs.requires_grad_()
        z = torch.sort(2.0 + 4.0 * torch.rand(*shape, s, generator=gen), dim=-1).values
        rd = torch.randn(*shape, 3, generator=gen)
        out = volume_render_radiance(rs, z, rd)
        tgt = torch.rand(*shape, 3, generator=gen)
        (g,) = torch.autograd.grad(((out - tgt) ** 2).mean(), [rs])
        cases.append(dict(rgb_sigma=rs.detach(), z_vals=z, rays_d=rd, out=out.detach(), target=tgt, d_rgb_sigma=g))
    save("render_packed", cases)


def posenc_cases():
    gen = torch.Generator().manual_seed(303)
    x = (torch.rand(200, 3, generator=gen) - 0.5) * 12.0
    cases = []
    for kw in [dict(num_freqs=10), dict(num_freqs=4), dict(num_freqs=6, log_sampling=False),
               dict(num_freqs=10, include_input=False), dict(num_freqs=0)]:
        cases.append(dict(kind="positional_encoding", kwargs=kw, x=x, out=PE3(**kw)(x)))
    x5 = torch.randn(3, 4, 5, generator=gen)
    cases.append(dict(kind="positional_encoding", kwargs=dict(num_freqs=3), x=x5, out=PE3(num_freqs=3)(x5)))
    for L in (12, 4):
        cases.append(dict(kind="nerf_mlp", kwargs=dict(num_freqs=L), x=x, out=PE4(L)(x)))
    save("posenc", cases)


def stratified_cases():
    gen = torch.Generator().manual_seed(404)
    cases = []
    # image-shaped (ray_sampler.py)
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.1, 0.2, 4.0])
    ro, rd = get_rays(6, 5, 7.5, c2w)
    ro = ro.contiguous()
    for perturb in (True, False):
        torch.manual_seed(11)
        pts, z = sample_hw(ro, rd, 2.0, 6.0, 64, perturb=perturb)
        torch.manual_seed(11)
        t_rand = torch.rand(6, 5, 64) if perturb else None
        cases.append(dict(kind="ray_sampler", rays_o=ro, rays_d=rd, near=2.0, far=6.0, n_samples=64,
                          t_rand=t_rand, lindisp=False, pts=pts, z=z.contiguous()))
    # flat (ray_utils.py)
    for (n, s, near, far, lindisp, perturb) in [(40, 64, 2.0, 6.0, False, True), (40, 64, 2.0, 6.0, True, True),
                                                (9, 1, 2.0, 6.0, False, True), (9, 2, 0.5, 3.0, False, True),
                                                (5, 37, 2.0, 6.0, False, True), (8, 192, 2.0, 6.0, False, False),
                                                (8, 128, 1.0, 9.0, True, False)]:
        o, d = rays(n, gen)
        torch.manual_seed(n * 1000 + s)
        pts, z = ray_utils.sample_points_along_rays(o, d, near, far, s, perturb=perturb, lindisp=lindisp)
        torch.manual_seed(n * 1000 + s)
        t_rand = torch.rand(n, s) if perturb else None
        cases.append(dict(kind="ray_utils", rays_o=o, rays_d=d, near=near, far=far, n_samples=s,
                          t_rand=t_rand, lindisp=lindisp, pts=pts, z=z.contiguous()))
    save("stratified", cases)


def hierarchical_cases():
    gen = torch.Generator().manual_seed(505)
    cases = []
    for (n, m1, ni, perturb, peaky) in [(64, 64, 128, True, True), (64, 64, 128, False, True),
                                        (16, 64, 128, True, False), (10, 17, 40, True, True),
                                        (6, 2, 5, True, False), (12, 64, 128, True, "zero")]:
        o, d = rays(n, gen)
        z = torch.sort(2.0 + 4.0 * torch.rand(n, m1, generator=gen), dim=-1).values
        w = torch.rand(n, m1 - 1, generator=gen)
        if peaky is True:
            w = w ** 8
        elif peaky == "zero":
            w = torch.zeros(n, m1 - 1)          # all mass from the +1e-5 floor; many ties
            w[:, 5] = 1.0
        grabbed = {}
        real = torch.searchsorted

        def spy(cdf, u, **kw):
            idx = real(cdf, u, **kw)
            grabbed.update(cdf=cdf.clone(), u=u.clone(), idx=idx.clone())
            return idx

        torch.searchsorted = spy
        try:
            torch.manual_seed(n + ni)
            pts, zc = ray_utils.hierarchical_sampling(o, d, z, w, ni, perturb=perturb)
        finally:
            torch.searchsorted = real
        cases.append(dict(rays_o=o, rays_d=d, z_vals=z, weights=w, n_importance=ni, perturb=perturb,
                          u=grabbed["u"], cdf=grabbed["cdf"], idx=grabbed["idx"], pts=pts, z=zc))
    save("hierarchical", cases)


def mlp_cases():
    gen = torch.Generator().manual_seed(606)
    cases = []
    for kw in [dict(), dict(pos_dim=75, hidden_dim=256, n_layers=8), dict(pos_dim=63, hidden_dim=128, n_layers=4)]:
        torch.manual_seed(1234)
        m = NeRFMLP(**kw)
        x = torch.randn(96, kw.get("pos_dim", 63), generator=gen)
        out = m(x)
        tgt = torch.rand_like(out)
        loss = ((out - tgt) ** 2).mean()
        grads = torch.autograd.grad(loss, list(m.parameters()))
        names = [k for k, _ in m.named_parameters()]
        cases.append(dict(kind="g1", seed=1234, kwargs=kw, x=x, out=out.detach(), target=tgt,
                          param_sums={k: float(v.double().sum()) for k, v in m.state_dict().items()},
                          grad_norms={k: float(g.double().norm()) for k, g in zip(names, grads)},
                          grad_sigma_out_w=grads[names.index("sigma_out.weight")],
                          grad_layer0_w_row0=grads[0][0].clone()))
    for kw in [dict(), dict(pos_freq=12, dino_dim=64), dict(dino_dim=0), dict(num_density_layers=2, hidden_dim=128)]:
        torch.manual_seed(4321)
        m = NeRFWithDINO(**kw)
        dd = kw.get("dino_dim", 64)
        pos = torch.randn(80, 3, generator=gen) * 2.0
        dirs = torch.randn(80, 3, generator=gen)
        feat = torch.randn(80, dd, generator=gen)
        rgb, den = m(pos, dirs, feat)
        cases.append(dict(kind="g3", seed=4321, kwargs=kw, positions=pos, directions=dirs, dino=feat,
                          rgb=rgb.detach(), density=den.detach(),
                          keys=list(m.state_dict().keys()),
                          param_sums={k: float(v.double().sum()) for k, v in m.state_dict().items()}))
    save("mlp", cases)


def gather_cases():
    """project_points_to_image + SpatialDINOFeatures.sample_features_at_points (the method body is pure
    torch; it is called unbound on a stand-in object so that no Dinov2 weights are needed)."""
    from models.dino_feature_model import SpatialDINOFeatures
    gen = torch.Generator().manual_seed(707)
    import math

    class _Self:
        pass

    cases = []
    for (n, hw, hp, c, spread) in [(257, 128, 9, 64, 1.0), (100, 100, 7, 32, 3.0), (64, 800, 16, 128, 0.5)]:
        th = float(torch.rand((), generator=gen)) * 2 * math.pi
        pose = torch.eye(4)
        pose[:3, :3] = torch.tensor([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
        pose[:3, 3] = torch.tensor([0.3, -0.2, 4.0])
        focal = 0.5 * hw / math.tan(0.5 * 0.6911112)
        pts = (torch.rand(n, 3, generator=gen) - 0.5) * 2 * spread
        pts[::7] += torch.tensor([0.0, 0.0, 9.0])          # some points behind the camera / far outside the image
        feats = torch.randn(1, hp, hp, c, generator=gen)
        p2d, depth, valid = ray_utils.project_points_to_image(pts, pose, focal, hw, hw)
        sampled = SpatialDINOFeatures.sample_features_at_points(_Self(), feats, p2d)
        cases.append(dict(points=pts, pose=pose, focal=focal, H=hw, W=hw, features=feats, points_2d=p2d, depths=depth,
                          valid=valid, sampled=sampled))
    save("gather", cases)


def loss_cases():
    gen = torch.Generator().manual_seed(707)
    pred = dict(rgb=torch.rand(50, 3, generator=gen), depth=torch.rand(50, generator=gen) * 6,
                weights=torch.rand(50, 64, generator=gen))
    tgt = dict(rgb=torch.rand(50, 3, generator=gen), depth=torch.rand(50, generator=gen) * 6)
    full = NeRFLoss()(pred, tgt)
    rgb_only = NeRFLoss(2.0, 0.5, 0.1)({"rgb": pred["rgb"]}, {"rgb": tgt["rgb"]})
    save("loss", dict(pred=pred, target=tgt, full={k: v.clone() for k, v in full.items()},
                      rgb_only={k: v.clone() for k, v in rgb_only.items()}))


def ray_cases():
    """get_rays of both modules (ray_sampler.py:4-30, ray_utils.py:4-37) and the batch gather of
    train.py:272-278 (rays / target colours of a randperm slice of the view's pixels)."""
    gen = torch.Generator().manual_seed(808)
    cases = []
    for (H, W, focal, n_batch) in [(100, 100, 138.88887889922103, 1024), (37, 53, 77.7, 200), (1, 1, 3.0, 1),
                                   (64, 48, 55.5 * (64 / 100), 512), (800, 800, 1111.1110311937682, 4096), (200, 200, 277.77775779844205, 4096)]:
        c2w = torch.eye(4)
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
        c2w[:3, :3] = q
        c2w[:3, 3] = torch.randn(3, generator=gen) * 3
        ro, rd = get_rays(H, W, focal, c2w)
        ro2, rd2 = ray_utils.get_rays(H, W, focal, c2w)
        assert torch.equal(ro, ro2) and torch.equal(rd, rd2)
        image = torch.rand(H, W, 3, generator=gen)
        idx = torch.randperm(H * W, generator=gen)[:n_batch].clone()
        full = H * W <= 4096                              # whole ray images only for the small views (fixture size)
        cases.append(dict(H=H, W=W, focal=focal, c2w=c2w, rays_o=ro.contiguous() if full else None,
                          rays_d=rd.contiguous() if full else None, rays_d_sum=rd.double().sum(), image=image if full else None,
                          idx=idx, batch_o=ro.reshape(-1, 3)[idx].clone(), batch_d=rd.reshape(-1, 3)[idx].clone(),
                          batch_target=image.view(-1, 3)[idx].clone()))
    c34 = cases[1]["c2w"][:3].clone()                     # (3,4) pose form (ray_sampler.py docstring)
    ro, rd = get_rays(37, 53, 77.7, c34)
    assert torch.equal(rd, cases[1]["rays_d"])
    save("rays", cases)


def ray_batch_cases():
    """utils.ray_utils.get_ray_batch (ray_utils.py:145-174): the generator of (rays_o, rays_d, pixel indices) slices."""
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.3, -0.2, 4.0])
    ro, rd = ray_utils.get_rays(7, 5, 9.5, c2w)
    ro = ro.contiguous()
    cases = []
    for bs in (8, 35, 100):
        batches = [(o.clone(), d.clone(), i.clone()) for o, d, i in ray_utils.get_ray_batch(ro, rd, batch_size=bs)]
        cases.append(dict(rays_o=ro, rays_d=rd, batch_size=bs, batches=batches))
    save("ray_batch", cases)


def render_loss_cases():
    """VolumeRenderer.forward followed by NeRFLoss (nerf_mlp.py:165-258) on rgb [+ depth] targets, with the
    autograd gradients back to the per-sample inputs: what nfs_composite_loss_fwd + nfs_composite_bwd fuse."""
    gen = torch.Generator().manual_seed(909)
    cases = []
    for (n, s, scale, white, with_depth, wr, wd) in [(40, 64, 10.0, False, True, 1.0, 0.1), (33, 192, 3.0, True, False, 1.0, 0.1),
                                                     (100, 64, 1.0, False, True, 2.0, 0.5), (5, 37, 30.0, False, False, 0.7, 0.0)]:
        _, rd = rays(n, gen)
        rgb = torch.rand(n, s, 3, generator=gen).requires_grad_()
        den = (torch.randn(n, s, 1, generator=gen) * scale).requires_grad_()
        z = torch.sort(2.0 + 4.0 * torch.rand(n, s, generator=gen), dim=-1).values
        tgt = dict(rgb=torch.rand(n, 3, generator=gen))
        if with_depth:
            tgt["depth"] = 2.0 + 4.0 * torch.rand(n, generator=gen)
        o_rgb, o_depth, o_w = VolumeRenderer().eval()(rgb, den, z, rd, white_bkgd=white)
        losses = NeRFLoss(wr, wd, 0.01)({"rgb": o_rgb, "depth": o_depth}, tgt)        # no 'weights' key: no reg term
        d_rgb, d_den = torch.autograd.grad(losses["total"], [rgb, den])
        cases.append(dict(rgb=rgb.detach(), density=den.detach(), z=z, rays_d=rd, white_bkgd=white, target=tgt,
                          rgb_weight=wr, depth_weight=wd, losses={k: v.detach().clone() for k, v in losses.items()},
                          out_rgb=o_rgb.detach(), out_depth=o_depth.detach(), out_w=o_w.detach(), d_rgb=d_rgb, d_density=d_den))
    save("render_loss", cases)


if __name__ == "__main__":
    if len(sys.argv) > 1:                                 # regenerate selected fixtures only
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    render_cases()
    packed_cases()
    posenc_cases()
    stratified_cases()
    hierarchical_cases()
    mlp_cases()
    loss_cases()
    gather_cases()
    ray_cases()
    ray_batch_cases()
    render_loss_cases()
