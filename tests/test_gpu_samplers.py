"""K4a / K4b parity (bit-exact z, indices and bins for fixed uniform draws) and K2 parity."""
import pytest
import torch

from helpers import bit_equal, record, ulp_diff
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def test_stratified_golden_bit_exact(golden, cuda):
    """Fixtures come from the reference on the build machine; z and pts must match bit for bit
    as long as this host's torch.linspace agrees with the fixture's (checked first)."""
    from nfs_b200 import ops
    for c in golden("stratified"):
        pts_o, z_o = O.stratified(c["rays_o"], c["rays_d"], c["near"], c["far"], c["n_samples"],
                                  t_rand=c["t_rand"], lindisp=c["lindisp"])
        same_host_arith = bit_equal(z_o.contiguous(), c["z"])
        t = None if c["t_rand"] is None else c["t_rand"].to(cuda)
        pts, z = ops.sample_stratified(c["rays_o"].to(cuda), c["rays_d"].to(cuda), c["near"], c["far"],
                                       c["n_samples"], t_rand=t, lindisp=c["lindisp"])
        assert pts.shape == c["pts"].shape and z.shape == c["z"].shape
        assert bit_equal(z, z_o.contiguous()) and bit_equal(pts, pts_o)
        if same_host_arith:
            assert bit_equal(z, c["z"]) and bit_equal(pts, c["pts"])


def test_stratified_api_draws_like_reference(cuda):
    """The drop-in functions consume one torch.rand(z.shape, device) from the global generator."""
    from models.ray_sampler import sample_points_along_rays as hw
    from utils.ray_utils import sample_points_along_rays as flat
    ro, rd = O.lego_rays(4096)
    ro, rd = ro.to(cuda), rd.to(cuda)
    torch.manual_seed(21)
    pts, z = flat(ro, rd, 2.0, 6.0, 64, perturb=True)
    torch.manual_seed(21)
    t = torch.rand(4096, 64, device=cuda)
    pts_o, z_o = O.stratified(ro.cpu(), rd.cpu(), 2.0, 6.0, 64, t_rand=t.cpu())
    assert bit_equal(z, z_o.contiguous()) and bit_equal(pts, pts_o)
    # image-shaped and flat inputs through models.ray_sampler
    torch.manual_seed(22)
    p2, z2 = hw(ro.reshape(64, 64, 3), rd.reshape(64, 64, 3), 2.0, 6.0, 48, perturb=True)
    torch.manual_seed(22)
    t = torch.rand(64, 64, 48, device=cuda)
    pts_o, z_o = O.stratified(ro.cpu().reshape(64, 64, 3), rd.cpu().reshape(64, 64, 3), 2.0, 6.0, 48, t_rand=t.cpu())
    assert p2.shape == (64, 64, 48, 3) and bit_equal(z2, z_o.contiguous()) and bit_equal(p2, pts_o)
    p3, z3 = hw(ro, rd, 2.0, 6.0, 64, perturb=False)
    pts_o, z_o = O.stratified(ro.cpu(), rd.cpu(), 2.0, 6.0, 64)
    assert bit_equal(z3, z_o.contiguous()) and bit_equal(p3, pts_o)
    # strata are ordered and inside [near, far]
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 2.0 and float(z.max()) <= 6.0
    assert flat(ro[:0], rd[:0], 2.0, 6.0, 64)[1].shape == (0, 64)


def _hier_inputs(n, m1, ni, seed, peaky=True):
    g = torch.Generator().manual_seed(seed)
    ro, rd = O.lego_rays(n, seed=seed)
    z = torch.sort(2 + 4 * torch.rand(n, m1, generator=g), -1).values
    w = torch.rand(n, m1 - 1, generator=g)
    if peaky:
        w = w ** 8
    u = torch.rand(n, ni, generator=g)
    return ro, rd, z, w, u


def _hier_kernel_level(c, dev):
    """Given the reference's cdf and draws: indices, samples, merged z and pts bit-identical."""
    from nfs_b200 import ops
    to = lambda t: t.to(dev)
    pts, z, dbg = ops.sample_hierarchical(to(c["rays_o"]), to(c["rays_d"]), to(c["z_vals"]), to(c["weights"]),
                                          c["u"].shape[-1], u=to(c["u"]), cdf=to(c["cdf"]), debug=True)
    assert bit_equal(dbg["idx"], c["idx"])
    assert bit_equal(z, c["z"]) and bit_equal(pts, c["pts"])
    return dbg


def test_hierarchical_golden_kernel_level(golden, cuda):
    for c in golden("hierarchical"):
        _hier_kernel_level(c, cuda)


@pytest.mark.parametrize("shape", [(4096, 64, 128), (513, 64, 64), (100, 33, 17), (64, 128, 256), (7, 2, 3)])
def test_hierarchical_live_oracle(cuda, shape):
    from nfs_b200 import ops
    n, m1, ni = shape
    ro, rd, z, w, u = _hier_inputs(n, m1, ni, seed=n + ni)
    r = O.hierarchical(ro, rd, z, w, u)
    c = dict(rays_o=ro, rays_d=rd, z_vals=z, weights=w, u=u, cdf=r["cdf"], idx=r["idx"], z=r["z"], pts=r["pts"])
    dbg = _hier_kernel_level(c, cuda)
    assert bit_equal(dbg["samples"], r["samples"])
    # End to end with the kernel's own cdf.  ATen's row sum is a vectorised cascade whose rounding
    # cannot be reproduced; the kernel's fp64-accumulated sum differs from
    # it by <= 2 ulp, which moves every cdf entry by a few ulp (measured: 4).  Consequences:
    #   - an index can differ only where a draw sits within that distance of a cdf edge;
    #   - a sample moves by at most (|d cdf| / den) * bin width with den >= 1e-5 (ray_utils.py:133),
    #     i.e. ill-conditioned (nearly empty) bins amplify the ulp difference - bound checked below,
    #     and all but a 1e-3 fraction of the samples agree to 1e-5.
    to = lambda t: t.to(cuda)
    pts, zz, own = ops.sample_hierarchical(to(ro), to(rd), to(z), to(w), ni, u=to(u), debug=True)
    cdf_ulps = ulp_diff(own["cdf"], r["cdf"])
    assert cdf_ulps <= 6, cdf_ulps
    mism = float((own["idx"].cpu() != r["idx"]).float().mean())
    assert mism <= 1e-4, mism
    dz = (zz.cpu() - r["z"]).abs()
    cdf_abs = float((own["cdf"].cpu() - r["cdf"]).abs().max()) + 2.0 ** -23
    bound = cdf_abs / 1e-5 * float((z[:, 1:] - z[:, :-1]).max()) + 1e-6
    frac_off = float((dz > 1e-5).float().mean())
    record("hierarchical_end_to_end", shape=list(shape), cdf_ulps=cdf_ulps, idx_mismatch_rate=mism,
           max_dz=float(dz.max()), bound=bound, frac_gt_1e5=frac_off)
    assert float(dz.max()) <= bound and frac_off <= 1e-3
    assert bool((zz[:, 1:] >= zz[:, :-1]).all())            # sortedness
    # every coarse depth survives the merge (multiset containment)
    both = torch.sort(torch.cat([z, own["samples"].cpu()], -1), -1).values
    assert bit_equal(zz, both)


@pytest.mark.parametrize("shape", [(300, 64, 128), (50, 33, 17), (9, 2, 3)])
def test_hierarchical_unsorted_coarse_depths(cuda, shape):
    """The reference sorts the concatenation of the coarse depths and the new samples (ray_utils.py:139), so a row of
    z_vals that is NOT ascending still yields the sorted multiset (and the bins are gathered in the given order).  The
    kernel merges by rank and therefore sorts such a row first: bit-identical z / pts given the oracle's cdf."""
    n, m1, ni = shape
    ro, rd, z, w, u = _hier_inputs(n, m1, ni, seed=n + 3 * ni)
    g = torch.Generator().manual_seed(5)
    perm = torch.stack([torch.randperm(m1, generator=g) for _ in range(n)])
    z_shuffled = torch.gather(z, 1, perm)
    z_shuffled[::3] = z[::3]                                  # every third ray stays sorted
    r = O.hierarchical(ro, rd, z_shuffled, w, u)
    c = dict(rays_o=ro, rays_d=rd, z_vals=z_shuffled, weights=w, u=u, cdf=r["cdf"], idx=r["idx"], z=r["z"], pts=r["pts"])
    _hier_kernel_level(c, cuda)


def test_hierarchical_api(cuda):
    from utils.ray_utils import hierarchical_sampling
    ro, rd, z, w, _ = _hier_inputs(256, 64, 128, seed=3)
    to = lambda t: t.to(cuda)
    torch.manual_seed(8)
    pts, zz = hierarchical_sampling(to(ro), to(rd), to(z), to(w), 128, perturb=True)
    torch.manual_seed(8)
    u = torch.rand(256, 128, device=cuda).cpu()
    r = O.hierarchical(ro, rd, z, w, u)
    dz = (zz.cpu() - r["z"]).abs()          # conditioning: see test_hierarchical_live_oracle
    assert pts.shape == (256, 192, 3) and float((dz > 1e-5).float().mean()) <= 1e-3 and float(dz.max()) <= 0.3
    pts, zz = hierarchical_sampling(to(ro), to(rd), to(z), to(w), 128, perturb=False)
    r = O.hierarchical(ro, rd, z, w, torch.linspace(0., 1., 128).expand(256, 128))
    dz = (zz.cpu() - r["z"]).abs()
    assert float((dz > 1e-5).float().mean()) <= 1e-3 and float(dz.max()) <= 0.3
    with pytest.raises(RuntimeError):       # the documented (N,S)/(N,S) call raises in the reference too
        hierarchical_sampling(to(ro), to(rd), to(z), to(z), 128)


def test_posenc_golden_and_live(golden, cuda):
    from models.nerf_mlp import PositionalEncoding as PE4
    from models.positional_encoding import PositionalEncoding as PE3
    for c in golden("posenc"):
        mod = PE3(**c["kwargs"]) if c["kind"] == "positional_encoding" else PE4(**c["kwargs"]).to(cuda)
        out = mod(c["x"].to(cuda))
        assert out.shape == c["out"].shape
        if out.numel():
            assert float((out.cpu() - c["out"]).abs().max()) <= 2e-6      # abs, SURVEY.md section 8c
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(100003, 3, generator=g) - 0.5) * 12
    ref = O.encode(x, O.frequency_bands(10))
    out = PE3(10)(x.to(cuda))
    assert float((out.cpu() - ref).abs().max()) <= 2e-6
    assert bit_equal(out[:, :3], x)
    assert PE3(10)(x[:0].to(cuda)).shape == (0, 63)
    assert "freq_bands" in PE4(4).state_dict() and "freq_bands" not in PE3(4).state_dict()


def test_project_gather_golden_and_live(golden, cuda):
    """K5 (8f rank 1): world points -> image coordinates -> bilinear feature lookup in one kernel, against the
    reference-generated fixture and the oracle on a larger live case.  Tolerances: coordinates / depths
    1e-5 relative to the coordinate scale (the reference's matmul sums in another order), the in-front mask
    exact away from cam_z = 0, sampled features 1e-4 absolute (a coordinate error of 1e-6 moves a bilinear
    tap weight by the same amount; feature values are O(1))."""
    from nfs_b200 import ops
    from models.dino_feature_model import sample_features_at_points
    from utils.ray_utils import project_points_to_image
    from oracle import nerf_oracle as O
    for c in golden("gather"):
        p2d, depth, valid, sampled = ops.project_gather(c["points"].to(cuda), c["pose"].to(cuda), c["focal"], c["H"],
                                                        c["W"], features=c["features"].to(cuda))
        scale = max(1.0, float(c["points_2d"].abs().max()))
        assert float((p2d.cpu() - c["points_2d"]).abs().max()) <= 1e-5 * scale
        assert float((depth.cpu() - c["depths"]).abs().max()) <= 1e-5
        safe = c["depths"].abs() > 1e-4
        assert torch.equal(valid.cpu()[safe], c["valid"][safe])
        assert float((sampled.cpu() - c["sampled"]).abs().max()) <= 1e-4 * max(1.0, scale)
        # the two reference entry points separately
        q2d, qd, qv = project_points_to_image(c["points"].to(cuda), c["pose"].to(cuda), c["focal"], c["H"], c["W"])
        assert torch.equal(q2d, p2d) and torch.equal(qd, depth) and torch.equal(qv, valid)
        s2 = sample_features_at_points(c["features"].to(cuda), c["points_2d"].to(cuda))
        assert float((s2.cpu() - c["sampled"]).abs().max()) <= 1e-5
    g = torch.Generator().manual_seed(5)
    P = 40000
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 3
    pose = torch.eye(4); pose[2, 3] = 4.0
    feats = torch.randn(1, 9, 9, 64, generator=g)
    r2d, rdp, rv = O.project_points(pts, pose, 180.0, 128, 128)
    rs = O.sample_features(feats, r2d)
    p2d, depth, valid, sampled = ops.project_gather(pts.to(cuda), pose.to(cuda), 180.0, 128, 128, features=feats.to(cuda))
    assert float((p2d.cpu() - r2d).abs().max()) <= 1e-5 and torch.equal(valid.cpu(), rv)
    assert float((sampled.cpu() - rs).abs().max()) <= 1e-4
    e2d, ed, ev = ops.project_gather(torch.zeros(0, 3, device=cuda), pose.to(cuda), 180.0, 128, 128)
    assert e2d.shape == (0, 2) and ev.shape == (0,)
    with pytest.raises(RuntimeError):
        ops.project_gather(pts, pose, 180.0, 128, 128)            # CPU tensors: no fallback
