"""8f rank 2 parity: K6 (ray generation + batch assembly, nfs_rays_generate) is bit-exact against the
reference-generated fixture and the CPU oracle; the loss epilogue of the compositing kernel
(nfs_composite_loss_fwd + nfs_composite_bwd) against the reference's VolumeRenderer + NeRFLoss.

Tolerances: rays / gathered colours bit-exact.  Loss terms: relative 1e-5 (the kernel sums the per-ray squared
errors in fp64; the reference's mse_loss is an fp32 cascade sum).  Gradients: the compositing tolerance of
tests/test_gpu_composite.py (per ray 1e-5 of max(max_s|ref|, 3 % of the batch maximum))."""
import math

import pytest
import torch

from helpers import bit_equal, per_ray_err, record, rel_err
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def test_rays_golden_bit_exact(golden, cuda):
    from models.ray_sampler import get_rays
    from nfs_b200 import ops
    from utils.ray_utils import get_rays as get_rays_u
    for c in golden("rays"):
        H, W, focal, c2w = c["H"], c["W"], c["focal"], c["c2w"].to(cuda)
        ro, rd = get_rays(H, W, focal, c2w)
        ro_u, rd_u = get_rays_u(H, W, focal, c2w)
        assert ro.shape == rd.shape == (H, W, 3) and bit_equal(rd, rd_u) and bit_equal(ro, ro_u)
        assert float(rd.double().sum()) == float(c["rays_d_sum"])
        if c["rays_d"] is not None:
            assert bit_equal(rd, c["rays_d"]) and bit_equal(ro, c["rays_o"])
        idx = c["idx"].to(cuda)
        if c["image"] is not None:
            bo, bd, bt = ops.generate_rays(H, W, focal, c2w, pix_idx=idx, image=c["image"].to(cuda))
            assert bit_equal(bt, c["batch_target"])
        else:
            bo, bd = ops.generate_rays(H, W, focal, c2w, pix_idx=idx)
        assert bit_equal(bo, c["batch_o"]) and bit_equal(bd, c["batch_d"])
    c = golden("rays")[1]                                        # (3,4) pose form
    _, rd34 = get_rays(c["H"], c["W"], c["focal"], c["c2w"][:3].to(cuda))
    assert bit_equal(rd34, c["rays_d"])


def test_rays_full_view_vs_oracle_and_edges(cuda):
    """The 800 x 800 benchmark view against the oracle, bit for bit; then the edge cases."""
    from nfs_b200 import ops
    H = W = 800
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    c2w = O.pose_spherical(37.0, -30.0, 4.0311)
    ro_ref, rd_ref = O.pixel_rays(H, W, focal, c2w)
    ro, rd = ops.generate_rays(H, W, focal, c2w.to(cuda))
    assert bit_equal(rd.reshape(H, W, 3), rd_ref) and bit_equal(ro.reshape(H, W, 3), ro_ref.contiguous())
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, H * W, (4096,), generator=g)          # with repeats, as torch indexing allows
    img = torch.rand(H, W, 3, generator=g)
    bo, bd, bt = ops.generate_rays(H, W, focal, c2w.to(cuda), pix_idx=idx.to(cuda), image=img.to(cuda))
    ref = O.ray_batch(H, W, focal, c2w, idx, img)
    assert bit_equal(bo, ref[0]) and bit_equal(bd, ref[1]) and bit_equal(bt, ref[2])
    # empty batch; int32 indices are accepted; an index outside the image poisons that ray only
    e = ops.generate_rays(H, W, focal, c2w.to(cuda), pix_idx=torch.zeros(0, dtype=torch.int64, device=cuda))
    assert e[0].shape == (0, 3) and e[1].shape == (0, 3)
    b32 = ops.generate_rays(H, W, focal, c2w.to(cuda), pix_idx=idx[:7].int().to(cuda))
    assert bit_equal(b32[1], ref[1][:7])
    bad = ops.generate_rays(H, W, focal, c2w.to(cuda), pix_idx=torch.tensor([5, H * W, -1, 6], device=cuda))
    assert bool(torch.isnan(bad[1][1:3]).all()) and not bool(torch.isnan(bad[1][[0, 3]]).any())
    with pytest.raises(RuntimeError):
        ops.generate_rays(H, W, focal, c2w)                      # CPU pose: no fallback
    with pytest.raises(RuntimeError):
        ops.generate_rays(H, W, focal, c2w.to(cuda), pix_idx=idx[:4].to(cuda), image=img[:10].to(cuda))


def _fused(c, dev, packed):
    from nfs_b200 import ops
    rgb = c["rgb"].to(dev).requires_grad_()
    den = c["density"].to(dev).requires_grad_()
    tgt = c["target"]
    td = tgt["depth"].to(dev) if "depth" in tgt else None
    if packed:
        raw = torch.cat([rgb, den], -1)
        out = ops.composite_loss(raw, None, c["z"].to(dev), c["rays_d"].to(dev), tgt["rgb"].to(dev), td,
                                 c["rgb_weight"], c["depth_weight"], c["white_bkgd"], want_weights=True)
    else:
        out = ops.composite_loss(rgb, den, c["z"].to(dev), c["rays_d"].to(dev), tgt["rgb"].to(dev), td,
                                 c["rgb_weight"], c["depth_weight"], c["white_bkgd"], want_weights=True)
    d_rgb, d_den = torch.autograd.grad(out["total"], [rgb, den])
    return out, d_rgb, d_den


@pytest.mark.parametrize("packed", [False, True])
def test_composite_loss_golden(golden, cuda, packed):
    for c in golden("render_loss"):
        out, d_rgb, d_den = _fused(c, cuda, packed)
        assert set(k for k in out if k in ("total", "rgb", "depth")) == set(c["losses"])
        for k, v in c["losses"].items():
            assert rel_err(out[k], v, floor=1e-6) <= 1e-5, k
        assert rel_err(out["rgb_map"], c["out_rgb"], floor=0.03) <= 1e-5
        assert rel_err(out["depth_map"], c["out_depth"], floor=0.03 * float(c["out_depth"].abs().max())) <= 1e-5
        assert per_ray_err(out["weights"], c["out_w"], floor=0.03) <= 1e-5
        e1 = per_ray_err(d_rgb, c["d_rgb"], floor_frac=0.03)
        e2 = per_ray_err(d_den, c["d_density"], floor_frac=0.03)
        record("composite_loss_golden", packed=packed, n=c["z"].shape[0], s=c["z"].shape[1], d_rgb=e1, d_density=e2)
        assert e1 <= 1e-5 and e2 <= 1e-5, (e1, e2)
        assert not out["rgb_map"].requires_grad and out["total"].requires_grad


def test_composite_loss_vs_unfused_route_large(cuda):
    """65 536 rays x 64: the fused loss equals VolumeRenderer + NeRFLoss on the same kernels (same compositing
    arithmetic, so the renderings are bit-identical and the gradients differ only by the rounding of the
    upstream gradient), an upstream scale is honoured, and the multi-block fp64 reduction is deterministic."""
    from models.nerf_mlp import NeRFLoss, VolumeRenderer
    from nfs_b200 import ops
    g = torch.Generator().manual_seed(11)
    N, S = 65536, 64
    rgb = torch.rand(N, S, 3, generator=g).to(cuda).requires_grad_()
    den = (torch.randn(N, S, 1, generator=g) * 5).to(cuda).requires_grad_()
    z = torch.sort(2 + 4 * torch.rand(N, S, generator=g), -1).values.to(cuda)
    rd = torch.randn(N, 3, generator=g).to(cuda)
    t_rgb, t_d = torch.rand(N, 3, generator=g).to(cuda), (2 + 4 * torch.rand(N, generator=g)).to(cuda)
    out = ops.composite_loss(rgb, den, z, rd, t_rgb, t_d, 1.0, 0.1)
    assert "weights" not in out
    d_rgb, d_den = torch.autograd.grad(out["total"] * 3.0, [rgb, den])
    o_rgb, o_depth, o_w = VolumeRenderer().eval()(rgb, den, z, rd)
    ref = NeRFLoss(1.0, 0.1, 0.0)({"rgb": o_rgb, "depth": o_depth}, {"rgb": t_rgb, "depth": t_d})
    r_rgb, r_den = torch.autograd.grad(ref["total"] * 3.0, [rgb, den])
    assert bit_equal(out["rgb_map"], o_rgb) and bit_equal(out["depth_map"], o_depth)
    for k in ("total", "rgb", "depth"):
        assert rel_err(out[k], ref[k]) <= 1e-5, k
    assert per_ray_err(d_rgb, r_rgb, floor_frac=0.03) <= 1e-5 and per_ray_err(d_den, r_den, floor_frac=0.03) <= 1e-5
    again = ops.composite_loss(rgb, den, z, rd, t_rgb, t_d, 1.0, 0.1)
    assert abs(float(again["total"].detach()) - float(out["total"].detach())) <= 1e-7 * float(out["total"].detach())
    with pytest.raises(RuntimeError):
        ops.composite_loss(rgb.detach().cpu(), den.detach().cpu(), z.cpu(), rd.cpu(), t_rgb.cpu())
    with pytest.raises(RuntimeError):
        ops.composite_loss(rgb[:0], den[:0], z[:0], rd[:0], t_rgb[:0])


def test_render_rays_fused_loss_matches_unfused(cuda):
    """pipeline.render_rays(target=...) - the loss of both passes in the compositing epilogues - against the torch-op
    loss on the same draws: same loss, same parameter gradients (up to the order of the fp32 atomics in wgrad)."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import pipeline
    bands = O.frequency_bands(10)
    ro, rd = O.lego_rays(1024, seed=4)
    ro, rd = ro.to(cuda), rd.to(cuda)
    tgt = torch.rand(1024, 3, generator=torch.Generator().manual_seed(1)).to(cuda)
    torch.manual_seed(0)
    model = NeRFMLP().to(cuda)
    res = []
    for fused in (True, False):
        torch.manual_seed(7)
        model.zero_grad()
        out = pipeline.render_rays(model, bands, ro, rd, 2.0, 6.0, 64, 128, target=tgt if fused else None)
        if fused:
            loss = out["loss"]
            assert not out["rgb"].requires_grad and out["weights"] is None
        else:
            loss = ((out["rgb"] - tgt) ** 2).mean() + ((out["rgb_coarse"] - tgt) ** 2).mean()
        loss.backward()
        res.append((float(loss), out["rgb"].detach().clone(), [p.grad.detach().clone() for p in model.parameters()]))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * abs(res[1][0])
    assert bit_equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-9
