"""End-to-end render path (sampler -> fused encoding + MLP -> compositing -> hierarchical -> ...)
against the same chain of oracle functions on the CPU, the fused optimiser against torch.optim,
and a short training run.  Stated tolerance: rendered pixel abs <= 1e-2."""
import pytest
import torch

from helpers import bit_equal, record
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def _models(cuda, seed=5, **kw):
    from models.nerf_model import NeRFMLP
    torch.manual_seed(seed)
    ref = O.PlainNeRF(**kw)
    mod = NeRFMLP(**kw)
    mod.load_state_dict(ref.state_dict())
    return ref, mod.to(cuda)


def _oracle_render(ref, ro, rd, t_rand, u, n_imp):
    bands = O.frequency_bands(10)
    pts, z = O.stratified(ro, rd, 2.0, 6.0, t_rand.shape[-1], t_rand=t_rand)
    raw = ref(O.encode(pts.reshape(-1, 3), bands)).reshape(*z.shape, 4)
    rgb_c, depth_c, w_c = O.render(raw[..., :3], raw[..., 3:4], z, rd)
    h = O.hierarchical(ro, rd, z, w_c[:, :-1].detach(), u)
    raw_f = ref(O.encode(h["pts"].reshape(-1, 3), bands)).reshape(*h["z"].shape, 4)
    rgb_f, depth_f, w_f = O.render(raw_f[..., :3], raw_f[..., 3:4], h["z"], rd)
    return rgb_c, rgb_f, depth_f


def test_render_rays_vs_oracle(cuda):
    from nfs_b200 import pipeline
    ref, mod = _models(cuda)
    n, S, Ni = 512, 64, 128
    g = torch.Generator().manual_seed(1)
    ro, rd = O.lego_rays(n, seed=2)
    t_rand, u = torch.rand(n, S, generator=g), torch.rand(n, Ni, generator=g)
    with torch.no_grad():
        rgb_c, rgb_f, depth_f = _oracle_render(ref, ro, rd, t_rand, u, Ni)
        out = pipeline.render_rays(mod, O.frequency_bands(10), ro.to(cuda), rd.to(cuda), 2.0, 6.0, S, Ni,
                                   perturb=True, t_rand=t_rand.to(cuda), u=u.to(cuda))
    e_c = float((out["rgb_coarse"].cpu() - rgb_c).abs().max())
    e_f = float((out["rgb"].cpu() - rgb_f).abs().max())
    e_d = float((out["depth"].cpu() - depth_f).abs().max())
    record("pipeline_vs_oracle", rgb_coarse_abs=e_c, rgb_fine_abs=e_f, depth_fine_abs=e_d)
    assert out["rgb"].shape == (n, 3) and out["z_vals"].shape == (n, S + Ni)
    assert e_c <= 1e-2 and e_f <= 1e-2 and e_d <= 5e-2, (e_c, e_f, e_d)


def test_full_size_train_step_vs_oracle(cuda):
    """The bench workload itself (BASELINE config 3: 4096 rays, 64 coarse + 192 fine evaluations, 1 048 576 points, the
    merged backward kernel) against the CPU oracle on the same draws: loss and every parameter gradient.  The draws are
    torch's CUDA generator's, reproduced by seeding (render_rays draws t_rand (N,64), then u (N,128))."""
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    ref, mod = _models(cuda, seed=9)
    with torch.no_grad():
        for m in (ref, mod):
            m.sigma_out.bias.fill_(0.3)                   # a live density at random init
    n, S, Ni = 4096, 64, 128
    ro, rd = O.lego_rays(n, seed=3)
    tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(6))
    torch.manual_seed(21)
    t_rand = torch.rand(n, S, device=cuda)
    u = torch.rand(n, Ni, device=cuda)
    opt = FusedAdam(mod.parameters(), lr=0.0)
    torch.manual_seed(21)
    loss = pipeline.train_step(mod.train(), opt, O.frequency_bands(10), ro.to(cuda), rd.to(cuda), tgt.to(cuda), 2.0, 6.0, S, Ni)
    sess = pipeline._session_for(mod, opt)
    assert sess is not None and sess.merged, "the bench-size step must take the merged backward kernel"
    g = opt.grad.detach().cpu()
    rgb_c, rgb_f, _ = _oracle_render(ref, ro, rd, t_rand.cpu(), u.cpu(), Ni)
    loss_ref = torch.mean((rgb_f - tgt) ** 2) + torch.mean((rgb_c - tgt) ** 2)
    g_ref = torch.autograd.grad(loss_ref, list(ref.parameters()))
    e_loss = abs(float(loss) - float(loss_ref)) / float(loss_ref)
    off, worst, worst_name = 0, 0.0, ""
    top = max(float(t.norm()) for t in g_ref)
    for (name, p), t in zip(ref.named_parameters(), g_ref):
        a = g[off:off + p.numel()].view_as(t)
        off += p.numel()
        rel = float((a - t).norm() / (t.norm() + 2e-4 * top))
        if rel > worst:
            worst, worst_name = rel, name
    flat_rel = float((g - torch.cat([t.reshape(-1) for t in g_ref])).norm() / torch.cat([t.reshape(-1) for t in g_ref]).norm())
    record("full_size_train_step_vs_oracle", loss_rel=e_loss, grad_flat_rel_l2=flat_rel, grad_worst_tensor_rel_l2=worst,
           worst=worst_name)
    assert e_loss <= 2e-3, e_loss
    assert flat_rel <= 5e-2 and worst <= 1e-1, (flat_rel, worst, worst_name)


def test_fused_adam_matches_torch(cuda):
    from nfs_b200.optim import FusedAdam
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(37, 5, device=cuda)), torch.nn.Parameter(torch.randn(11, device=cuda))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    for decoupled, wd in ((False, 0.0), (False, 1e-2), (True, 1e-2)):
        fa = FusedAdam(a, lr=1e-2, weight_decay=wd, decoupled=decoupled)
        ta = (torch.optim.AdamW if decoupled else torch.optim.Adam)(b, lr=1e-2, weight_decay=wd)
        for step in range(5):
            gs = [torch.randn_like(p) for p in a]
            for p, q, g_ in zip(a, b, gs):
                p.grad, q.grad = g_.clone(), g_.clone()
            fa.step(); ta.step()
        for p, q in zip(a, b):
            assert float((p - q).abs().max()) <= 2e-6, (decoupled, wd)
    assert a[0].data_ptr() == fa.flat.data_ptr()          # parameters are views of the flat buffer


def test_train_step_learns(cuda):
    """BASELINE config 3 shape at a small batch: coarse 64 + fine 192 evaluations per ray."""
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    _, mod = _models(cuda, seed=9)
    mod.train()
    with torch.no_grad():
        mod.sigma_out.bias.fill_(0.3)      # default init can start with sigma <= 0 everywhere: relu-dead, zero gradient
    opt = FusedAdam(mod.parameters(), lr=5e-4)
    n = 1024
    ro, rd = O.lego_rays(n, seed=3)
    ro, rd = ro.to(cuda), rd.to(cuda)
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(1)).to(cuda) * 0.2 + 0.6
    bands = O.frequency_bands(10)
    losses = [float(pipeline.train_step(mod, opt, bands, ro, rd, target, 2.0, 6.0, 64, 128)) for _ in range(40)]
    record("train_step_learns", first=losses[0], last=losses[-1])
    assert losses[-1] < 0.5 * losses[0], losses[::8]
    assert all(torch.isfinite(p).all() for p in mod.parameters())
    img = pipeline.render_image(mod.eval(), bands, ro, rd, 2.0, 6.0, 64, 128, chunk=300)
    assert img.shape == (n, 3) and bool(torch.isfinite(img).all())


def test_graphed_train_step_matches_eager(cuda):
    """The CUDA-graph replay of the training step (one cudaGraphLaunch per step) follows the eager
    step exactly: same kernels, same order, same deterministic draws (perturb=False)."""
    import bench
    from models.nerf_model import NeRFMLP
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    N = 1024
    ro, rd = bench.lego_rays(N, seed=3)
    ro, rd = ro.to(cuda), rd.to(cuda)
    g = torch.Generator().manual_seed(0)
    targets = [torch.rand(N, 3, generator=g).to(cuda) for _ in range(4)]
    bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
    losses = {}
    finals = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(7)
        model = NeRFMLP().to(cuda).train()
        opt = FusedAdam(model.parameters(), lr=5e-4)
        step = None
        if mode == "graph":
            step = pipeline.GraphedTrainStep(model, opt, bands, N, 2.0, 6.0, 64, 128, perturb=False)
        ls = []
        for t in targets:
            if step is None:
                ls.append(float(pipeline.train_step(model, opt, bands, ro, rd, t, 2.0, 6.0, 64, 128, perturb=False)))
            else:
                ls.append(float(step(ro, rd, t)))
        losses[mode] = ls
        finals[mode] = opt.flat.clone()
        assert opt.step_count == len(targets)
    for a, b in zip(losses["eager"], losses["graph"]):
        assert abs(a - b) <= 2e-3 * abs(a), (losses["eager"], losses["graph"])      # wgrad atomics reorder fp32 sums
    assert float((finals["eager"] - finals["graph"]).abs().max()) <= 5e-3
    assert losses["graph"][-1] < losses["graph"][0] * 1.5


def test_conditioned_render_vs_oracle(cuda):
    """BASELINE config 4 path (train.py:188-242 with use_dino): sampler -> projection + feature lookup ->
    NeRFWithDINO (pos_freq 12, 64-d features) -> compositing against the oracle chain on the same draws."""
    import math
    import bench
    from models.nerf_mlp import NeRFWithDINO
    from nfs_b200 import pipeline
    from oracle import nerf_oracle as O
    from helpers import record
    N, S = 512, 64
    ro, rd = bench.lego_rays(N, H=128, W=128, seed=2)
    g = torch.Generator().manual_seed(4)
    t_rand = torch.rand(N, S, generator=g)
    feats = torch.randn(1, 9, 9, 64, generator=g)
    pose = torch.eye(4); pose[:3, 3] = torch.tensor([0.2, -0.1, 4.0])
    focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
    torch.manual_seed(21)
    ref = O.ConditionedNeRF(pos_freq=12, dino_dim=64)
    mod = NeRFWithDINO(pos_freq=12, dino_dim=64)
    sd = ref.state_dict()
    sd["density_mlp.density_head.bias"].fill_(0.3)      # the default init leaves relu(density) = 0 everywhere here:
    ref.load_state_dict(sd)                              # nothing would be composited and the comparison be vacuous
    mod.load_state_dict(sd)
    mod = mod.to(cuda).eval()
    # oracle chain
    pts, z = O.stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
    p2d, _, _ = O.project_points(pts.reshape(-1, 3), pose, focal, 128, 128)
    f = O.sample_features(feats, p2d)
    rgb, den = ref(pts.reshape(-1, 3), rd.unsqueeze(1).expand(-1, S, -1).reshape(-1, 3), f)
    ref_out = O.render(rgb.reshape(N, S, 3), den.reshape(N, S, 1), z, rd)
    with torch.no_grad():
        out = pipeline.render_rays_conditioned(mod, ro.to(cuda), rd.to(cuda), 2.0, 6.0, S, pose.to(cuda), focal, 128, 128,
                                               feats.to(cuda), perturb=True, t_rand=t_rand.to(cuda))
    e_rgb = float((out["rgb"].cpu() - ref_out[0].detach()).abs().max())
    e_depth = float((out["depth"].cpu() - ref_out[1].detach()).abs().max())
    record("conditioned_pipeline_vs_oracle", rgb_abs=e_rgb, depth_abs=e_depth)
    assert torch.equal(out["z_vals"].cpu(), z)
    assert float(ref_out[0].abs().max()) > 0.1 and float(ref_out[1].abs().max()) > 1.0       # a live rendering
    assert e_rgb <= 1e-2 and e_depth <= 5e-2, (e_rgb, e_depth)


@pytest.mark.parametrize("n_rays,n_coarse,n_imp,kwargs", [(1024, 64, 128, {}), (1000, 33, 20, {}), (77, 64, 0, {}),
                                                         (600, 64, 128, dict(n_layers=4)), (300, 48, 0, dict(n_layers=2)),
                                                         (1, 64, 128, {})])
def test_step_session_matches_per_call_gradients(cuda, monkeypatch, n_rays, n_coarse, n_imp, kwargs):
    """mlp.StepSession (one weight-gradient launch per layer over the coarse + fine points, written straight into
    the optimizer's flat gradient) against the per-call route (autograd accumulation + gather_grads), same draws:
    identical forward, gradients equal up to the fp32 summation order (rel L2 <= 1e-4); ragged point counts put
    zero-gradient rows between the calls."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import mlp, pipeline
    from nfs_b200.optim import FusedAdam
    ro, rd = O.lego_rays(n_rays, seed=5)
    ro, rd = ro.to(cuda), rd.to(cuda)
    tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(2)).to(cuda)
    bands = O.frequency_bands(10)
    res = {}
    for mode in ("session", "session_merged", "per_call"):
        monkeypatch.setenv("NFS_MLP_SESSION", "0" if mode == "per_call" else "1")
        # "session": dgrad chain per call + a weight-gradient launch per layer (default); "session_merged": the whole
        # backward pass is ONE launch (nfs_mlp_backward_fused: dgrad chain on producer CTA pairs, weight gradients on
        # consumer CTAs fed through L2, opt-in)
        monkeypatch.setenv("NFS_BWD_MERGED", "1" if mode == "session_merged" else "0")
        torch.manual_seed(3)
        model = NeRFMLP(**kwargs).to(cuda).train()
        with torch.no_grad():
            model.sigma_out.bias.fill_(0.3)
        opt = FusedAdam(model.parameters(), lr=5e-4)
        assert (pipeline._session_for(model, opt) is not None) == (mode != "per_call")
        torch.manual_seed(11)
        loss = pipeline.train_step(model, opt, bands, ro, rd, tgt, 2.0, 6.0, n_coarse, n_imp)
        res[mode] = (float(loss), opt.grad.clone(), opt.flat.clone())
        assert getattr(model._get_plan(), "_session", None) is None
    assert res["session"][0] == res["per_call"][0] == res["session_merged"][0]
    g_s, g_x, g_p = res["session"][1], res["session_merged"][1], res["per_call"][1]
    assert float(g_p.norm()) > 0
    rel = float((g_s - g_p).norm() / g_p.norm())
    rel_x = float((g_x - g_p).norm() / g_p.norm())
    rel_sx = float((g_s - g_x).norm() / g_x.norm())
    record("step_session_vs_per_call", n_rays=n_rays, n_coarse=n_coarse, n_imp=n_imp, kwargs=str(kwargs), grad_rel_l2=rel,
           merged_grad_rel_l2=rel_x, merged_vs_split_rel_l2=rel_sx)
    assert rel <= 1e-4 and rel_x <= 1e-4 and rel_sx <= 1e-4, (rel, rel_x, rel_sx)


@pytest.mark.parametrize("S,white", [(64, False), (192, True), (37, False)])
def test_composite_bwd_dy_bit_identical_to_two_kernel_route(cuda, S, white):
    """nfs_composite_bwd_dy (compositing backward + derivative of the MLP head, bf16 operand rows written in place)
    against nfs_composite_bwd (packed) followed by nfs_act_grad_bf16: bit-identical columns 0..3, other columns untouched."""
    from nfs_b200 import _lib, mlp, ops
    from nfs_b200._lib import ptr
    g = torch.Generator().manual_seed(S)
    n = 1500
    raw = torch.cat([torch.rand(n, S, 3, generator=g), torch.randn(n, S, 1, generator=g) * 5], -1).to(cuda)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), -1).values.to(cuda)
    _, rd = O.lego_rays(n, seed=S)
    rd = rd.to(cuda)
    g_rgb, g_depth = torch.randn(n, 3, generator=g).to(cuda), torch.randn(n, generator=g).to(cuda)
    d_raw = torch.empty_like(raw)
    stream = ops._stream()
    _lib.call("nfs_composite_bwd", ptr(raw), None, ptr(z), ptr(rd), None, 0.0, ptr(g_rgb), ptr(g_depth), None, n, S, int(white), 1,
              ptr(d_raw), None, stream)
    want = mlp.act_grad(raw.reshape(-1, 4), d_raw.reshape(-1, 4), 2, 64)
    got = torch.full((n * S, 64), 7.0, device=cuda, dtype=torch.bfloat16)
    _lib.call("nfs_composite_bwd_dy", ptr(raw), ptr(z), ptr(rd), ptr(g_rgb), ptr(g_depth), None, n, S, int(white), ptr(got), 64,
              stream)
    assert torch.equal(got[:, :4].view(torch.int16), want[:, :4].view(torch.int16))
    assert bool((got[:, 4:] == 7.0).all())


def test_session_with_and_without_in_place_dy(cuda, monkeypatch):
    """The training step with the compositing backward writing dY in place (default) against the fp32 d(rgb_sigma) +
    nfs_act_grad_bf16 route: same loss, gradients equal up to the weight-gradient kernels' atomics."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    n = 700
    ro, rd = O.lego_rays(n, seed=8)
    ro, rd = ro.to(cuda), rd.to(cuda)
    tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(4)).to(cuda)
    bands = O.frequency_bands(10)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NFS_K1_BWD_DY", mode)
        torch.manual_seed(3)
        model = NeRFMLP().to(cuda).train()
        with torch.no_grad():
            model.sigma_out.bias.fill_(0.3)
        opt = FusedAdam(model.parameters(), lr=5e-4)
        torch.manual_seed(11)
        loss = pipeline.train_step(model, opt, bands, ro, rd, tgt, 2.0, 6.0, 48, 80)
        res[mode] = (float(loss), opt.grad.clone())
    assert res["1"][0] == res["0"][0]
    rel = float((res["1"][1] - res["0"][1]).norm() / res["0"][1].norm())
    record("step_in_place_dy_vs_act_grad", grad_rel_l2=rel)
    assert rel <= 1e-5, rel


@pytest.mark.parametrize("in_place", ["1", "0"])
def test_merged_backward_reads_no_stale_rows(cuda, monkeypatch, in_place):
    """Hand-off check of the merged backward kernel (nfs_mlp_backward_fused): the same step is run again from the same
    seeds after the dY arena has been filled with NaN.  A weight-gradient consumer that loads a row before the dgrad
    chain has stored it - or a drain that overwrites a staged slab another warp is still summing - shows up as NaN / a
    changed gradient; equal seeds alone would hide it (the stale rows of the previous run hold the same values)."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    monkeypatch.setenv("NFS_K1_BWD_DY", in_place)
    monkeypatch.setenv("NFS_BWD_MERGED", "1")
    n = 700
    ro, rd = O.lego_rays(n, seed=8)
    ro, rd = ro.to(cuda), rd.to(cuda)
    tgt = torch.rand(n, 3, generator=torch.Generator().manual_seed(4)).to(cuda)
    bands = O.frequency_bands(10)
    torch.manual_seed(3)
    model = NeRFMLP().to(cuda).train()
    with torch.no_grad():
        model.sigma_out.bias.fill_(0.3)
    opt = FusedAdam(model.parameters(), lr=0.0)
    ref, worst = None, 0.0
    for rep in range(4):
        sess = pipeline._session_for(model, opt)
        if rep > 0:
            sess.dys.fill_(float("nan"))
        torch.manual_seed(11)
        pipeline.train_step(model, opt, bands, ro, rd, tgt, 2.0, 6.0, 48, 80)
        assert sess.merged
        g = opt.grad.clone()
        if ref is None:
            ref = g
            continue
        assert not bool(torch.isnan(g).any()), "a consumer read dY rows the chain had not stored yet"
        off = 0
        for p in model.parameters():        # per tensor: the head's bias gradients are tiny next to the weights'
            a, b = g[off:off + p.numel()], ref[off:off + p.numel()]
            worst = max(worst, float((a - b).norm() / b.norm().clamp_min(1e-20)))
            off += p.numel()
    record("merged_backward_poisoned_arena", in_place_dy=in_place, worst_tensor_rel_l2=worst)
    assert worst <= 1e-5, worst


@pytest.mark.parametrize("n_rays,S", [(300, 64), (257, 192), (5, 7)])
def test_forward_rays_bit_identical_to_forward_points(cuda, n_rays, S):
    """nfs_mlp_chain_rays (sampler o + d z evaluated inside the chain kernel) against sampling the positions first
    (nfs_sample_stratified + nfs_mlp_chain_points): bit-identical outputs in inference and in a training forward, equal
    gradients (up to the weight-gradient kernels' atomics)."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import ops
    torch.manual_seed(1)
    model = NeRFMLP().to(cuda)
    bands = O.frequency_bands(10)
    ro, rd = O.lego_rays(n_rays, seed=S)
    ro, rd = ro.to(cuda), rd.to(cuda)
    t_rand = torch.rand(n_rays, S, generator=torch.Generator().manual_seed(2)).to(cuda)
    pts, z = ops.sample_stratified(ro, rd, 2.0, 6.0, S, t_rand=t_rand)
    with torch.no_grad():
        a = model.forward_rays(ro, rd, z, bands)
        b = model.forward_points(pts.reshape(-1, 3), bands).reshape(n_rays, S, 4)
    assert bit_equal(a, b)
    ga = torch.autograd.grad((model.forward_rays(ro, rd, z, bands) ** 2).mean(), list(model.parameters()))
    gb = torch.autograd.grad((model.forward_points(pts.reshape(-1, 3), bands) ** 2).mean(), list(model.parameters()))
    for x, y in zip(ga, gb):
        assert float((x - y).norm()) <= 1e-5 * float(y.norm()) + 1e-12


@pytest.mark.parametrize("n_imp,perturb", [(128, False), (40, True), (0, True)])
def test_render_fused_matches_kernel_by_kernel_route(cuda, monkeypatch, n_imp, perturb):
    """nfs_render_fused_fwd (one C call; sampler + encoding + MLP in one kernel) against the launch-by-launch route
    with materialised positions: bit-identical depths, weights and pixels."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import pipeline
    torch.manual_seed(2)
    model = NeRFMLP().to(cuda).eval()
    with torch.no_grad():
        model.sigma_out.bias.fill_(0.3)
    bands = O.frequency_bands(10)
    n = 777
    ro, rd = O.lego_rays(n, seed=3)
    ro, rd = ro.to(cuda), rd.to(cuda)
    g = torch.Generator().manual_seed(5)
    t_rand = torch.rand(n, 64, generator=g).to(cuda)
    u = torch.rand(n, n_imp, generator=g).to(cuda) if (perturb and n_imp) else None
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NFS_RENDER_FUSED", mode)
        with torch.no_grad():
            outs[mode] = pipeline.render_rays(model, bands, ro, rd, 2.0, 6.0, 64, n_imp, perturb=perturb, t_rand=t_rand, u=u,
                                              white_bkgd=True)
    for k in ("rgb", "depth", "weights", "z_vals") + (("rgb_coarse", "weights_coarse") if n_imp else ()):
        assert bit_equal(outs["1"][k], outs["0"][k]), k
