"""oracle/nerf_oracle.py against the fixtures produced by the reference itself
(tests/golden/make_golden.py).  Same ATen ops in the same order: the results are expected to
be bit-identical on the machine that generated the fixtures; a 2e-6 relative allowance
covers hosts whose ATen CPU kernels use another vector width (sum / exp / linspace are
width-dependent).  Integer outputs (searchsorted indices) must match
exactly except where that allowance moves a cdf edge across a draw."""
import torch

from helpers import bit_equal, per_ray_err, rel_err
from oracle import nerf_oracle as O

TOL = 2e-6


def close(a, b, tol=TOL):
    return bit_equal(a, b) or rel_err(a, b, floor=1e-3) <= tol


def test_render_forward_and_autograd(golden):
    for c in golden("render"):
        rgb = c["rgb"].clone().requires_grad_()
        den = c["density"].clone().requires_grad_()
        o_rgb, o_depth, o_w = O.render(rgb, den, c["z_vals"], c["rays_d"], noise=c.get("noise"),
                                       noise_std=c.get("noise_std", 0.0), white_bkgd=c["white_bkgd"])
        assert close(o_rgb, c["out_rgb"]) and close(o_depth, c["out_depth"])
        assert per_ray_err(o_w, c["out_weights"]) <= TOL
        if "g_rgb" in c:
            d_rgb, d_den = torch.autograd.grad([o_rgb, o_depth, o_w], [rgb, den],
                                               [c["g_rgb"], c["g_depth"], c["g_weights"]])
            assert per_ray_err(d_rgb, c["d_rgb"]) <= TOL
            assert per_ray_err(d_den, c["d_density"]) <= TOL


def test_render_closed_form_backward_matches_autograd_fp64(golden):
    for c in golden("render"):
        if "g_rgb" not in c or c["z_vals"].shape[-1] == 1:   # S=1: the reference renders nothing (empty dists)
            continue
        dd = lambda t: t.double()
        rgb, den = dd(c["rgb"]).requires_grad_(), dd(c["density"]).requires_grad_()
        outs = O.render(rgb, den, dd(c["z_vals"]), dd(c["rays_d"]), white_bkgd=c["white_bkgd"])
        gs = [dd(c["g_rgb"]), dd(c["g_depth"]), dd(c["g_weights"])]
        a_rgb, a_den = torch.autograd.grad(list(outs), [rgb, den], gs)
        f_rgb, f_den = O.render_backward_closed_form(rgb.detach(), den.detach(), dd(c["z_vals"]), dd(c["rays_d"]),
                                                     gs[0], gs[1], gs[2], white_bkgd=c["white_bkgd"])
        assert per_ray_err(f_rgb, a_rgb) < 1e-12
        assert per_ray_err(f_den, a_den) < 1e-6   # saturated rays: suffix/q with q ~ 1e-10 amplifies fp64 rounding


def test_render_packed(golden):
    for c in golden("render_packed"):
        rs = c["rgb_sigma"].clone().requires_grad_()
        out = O.render_packed(rs, c["z_vals"], c["rays_d"])
        assert close(out, c["out"])
        (g,) = torch.autograd.grad(((out - c["target"]) ** 2).mean(), [rs])
        n = g.shape[:-2].numel()
        assert per_ray_err(g.reshape(n, -1), c["d_rgb_sigma"].reshape(n, -1)) <= TOL


def test_posenc(golden):
    for c in golden("posenc"):
        kw = c["kwargs"]
        bands = O.frequency_bands(kw["num_freqs"], kw.get("log_sampling", True))
        out = O.encode(c["x"], bands, kw.get("include_input", True))
        assert out.shape == c["out"].shape
        assert bit_equal(out, c["out"]) or float((out - c["out"]).abs().max()) <= 1e-6


def test_stratified(golden):
    for c in golden("stratified"):
        pts, z = O.stratified(c["rays_o"], c["rays_d"], c["near"], c["far"], c["n_samples"],
                              t_rand=c["t_rand"], lindisp=c["lindisp"])
        assert close(z.contiguous(), c["z"]) and close(pts, c["pts"])


def test_hierarchical(golden):
    for c in golden("hierarchical"):
        r = O.hierarchical(c["rays_o"], c["rays_d"], c["z_vals"], c["weights"], c["u"])
        assert close(r["cdf"], c["cdf"])
        same_cdf = bit_equal(r["cdf"], c["cdf"])
        mismatch = float((r["idx"] != c["idx"]).float().mean())
        assert mismatch == 0.0 if same_cdf else mismatch < 1e-3
        if same_cdf:
            assert bit_equal(r["z"], c["z"]) and bit_equal(r["pts"], c["pts"])
        else:
            assert float((r["z"] - c["z"]).abs().max()) < 1e-4


def test_mlp_g1(golden):
    for c in golden("mlp"):
        if c["kind"] != "g1":
            continue
        torch.manual_seed(c["seed"])
        m = O.PlainNeRF(**c["kwargs"])
        for k, v in m.state_dict().items():
            assert abs(float(v.double().sum()) - c["param_sums"][k]) < 1e-9, k
        out = m(c["x"])
        assert float((out - c["out"]).abs().max()) <= 1e-6
        loss = ((out - c["target"]) ** 2).mean()
        grads = dict(zip([k for k, _ in m.named_parameters()], torch.autograd.grad(loss, list(m.parameters()))))
        for k, g in grads.items():
            assert abs(float(g.double().norm()) - c["grad_norms"][k]) <= 1e-5 * max(c["grad_norms"][k], 1e-6), k


def test_mlp_g3(golden):
    for c in golden("mlp"):
        if c["kind"] != "g3":
            continue
        torch.manual_seed(c["seed"])
        m = O.ConditionedNeRF(**{k: v for k, v in c["kwargs"].items()})
        assert list(m.state_dict().keys()) == c["keys"]
        for k, v in m.state_dict().items():
            assert abs(float(v.double().sum()) - c["param_sums"][k]) < 1e-9, k
        rgb, den = m(c["positions"], c["directions"], c["dino"])
        assert float((rgb - c["rgb"]).abs().max()) <= 1e-6
        assert float((den - c["density"]).abs().max()) <= 1e-5


def test_loss(golden):
    c = golden("loss")
    full = O.nerf_loss(c["pred"], c["target"])
    for k, v in c["full"].items():
        assert close(full[k], v), k
    only = O.nerf_loss({"rgb": c["pred"]["rgb"]}, {"rgb": c["target"]["rgb"]}, 2.0, 0.5, 0.1)
    for k, v in c["rgb_only"].items():
        assert close(only[k], v), k


def test_project_and_sample_features(golden):
    """8f rank 1: project_points_to_image + sample_features_at_points against the reference's outputs."""
    for c in golden("gather"):
        p2d, depth, valid = O.project_points(c["points"], c["pose"], c["focal"], c["H"], c["W"])
        assert torch.equal(valid, c["valid"])
        assert float((p2d - c["points_2d"]).abs().max()) <= 1e-5 * max(1.0, float(c["points_2d"].abs().max()))
        assert float((depth - c["depths"]).abs().max()) <= 1e-6
        sampled = O.sample_features(c["features"], c["points_2d"])
        assert float((sampled - c["sampled"]).abs().max()) <= 1e-6


def test_rays(golden):
    """8f rank 2: get_rays (both reference modules) and the per-batch gather of train.py:272-278 - bit-exact,
    also for the scalar fp32 restatement the CUDA kernel follows."""
    for c in golden("rays"):
        H, W = c["H"], c["W"]
        ro, rd = O.pixel_rays(H, W, c["focal"], c["c2w"])
        ro_s, rd_s = O.pixel_rays_scalar(H, W, c["focal"], c["c2w"])
        assert torch.equal(rd, rd_s) and torch.equal(ro, ro_s)
        assert float(rd.double().sum()) == float(c["rays_d_sum"])
        if c["rays_d"] is not None:
            assert torch.equal(rd, c["rays_d"]) and torch.equal(ro.contiguous(), c["rays_o"])
        img = c["image"]
        got = O.ray_batch(H, W, c["focal"], c["c2w"], c["idx"], img)
        assert torch.equal(got[0], c["batch_o"]) and torch.equal(got[1], c["batch_d"])
        if img is not None:
            assert torch.equal(got[2], c["batch_target"])


def test_render_loss(golden):
    """VolumeRenderer + NeRFLoss (rgb [+ depth] terms) and their autograd gradients: the composition the fused
    compositing + loss kernel is checked against."""
    for c in golden("render_loss"):
        rgb, den = c["rgb"].clone().requires_grad_(), c["density"].clone().requires_grad_()
        o_rgb, o_depth, o_w = O.render(rgb, den, c["z"], c["rays_d"], white_bkgd=c["white_bkgd"])
        losses = O.nerf_loss({"rgb": o_rgb, "depth": o_depth}, c["target"], c["rgb_weight"], c["depth_weight"], 0.01)
        assert set(losses) == set(c["losses"])
        for k, v in c["losses"].items():
            assert close(losses[k], v), k
        d_rgb, d_den = torch.autograd.grad(losses["total"], [rgb, den])
        assert per_ray_err(d_rgb, c["d_rgb"]) <= TOL and per_ray_err(d_den, c["d_density"]) <= TOL
