"""K3 parity: tcgen05 dense layers against a plain PyTorch fp32 evaluation of the same bf16
operands, then the MLP modules against the CPU oracle (fp32) within the stated bf16 tolerance."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("P,K,N", [(128, 64, 256), (1000, 256, 256), (4096 * 5 + 77, 256, 256), (300, 64, 64),
                                   (2500, 320, 128), (513, 192, 256), (129, 128, 64), (1, 256, 32)])
def test_linear_vs_torch(cuda, P, K, N):
    from nfs_b200 import ops
    g = torch.Generator().manual_seed(P + K + N)
    x = (torch.randn(P, K, generator=g)).to(torch.bfloat16).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    ref = x.float() @ w.float().t() + b
    y16, y32 = ops.linear_bf16(x, w, b, act=0, out_f32_cols=N)
    assert float((y32 - ref).abs().max()) <= 2e-4 * float(ref.abs().max())      # fp32 accumulate, other order
    assert float((y16.float() - ref).abs().max()) <= 8e-3 * float(ref.abs().max())   # + one bf16 rounding
    y16r, _ = ops.linear_bf16(x, w, b, act=1)
    assert torch.equal(y16r, torch.relu(y16r)) and float((y16r.float() - torch.relu(ref)).abs().max()) <= 8e-3 * float(ref.abs().max())
    # fused relu-backward mask and partial fp32 output
    m = torch.randn(P, N, generator=g).to(torch.bfloat16).to(cuda)
    y16m, y32m = ops.linear_bf16(x, w, None, act=0, relu_mask_src=m, out_f32_cols=4)
    refm = (x.float() @ w.float().t()) * (m.float() > 0)
    assert float((y16m.float() - refm).abs().max()) <= 8e-3 * float(ref.abs().max())
    assert float((y32m - refm[:, :4]).abs().max()) <= 2e-4 * float(ref.abs().max())
    # head activation: sigmoid on the first three columns only
    _, yh = ops.linear_bf16(x, w, b, act=2, out_bf16=False, out_f32_cols=4)
    refh = torch.cat([torch.sigmoid(ref[:, :3]), ref[:, 3:4]], -1)
    assert float((yh - refh).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("P,M,N", [(64, 128, 64), (1000, 256, 256), (70001, 256, 256), (4097, 256, 64), (300, 128, 256)])
def test_wgrad_vs_torch(cuda, P, M, N):
    from nfs_b200 import ops
    g = torch.Generator().manual_seed(P + M + N)
    u = torch.randn(P, M, generator=g).to(torch.bfloat16).to(cuda)
    v = torch.randn(P, N, generator=g).to(torch.bfloat16).to(cuda)
    ref = u.double().t() @ v.double()
    scale = float(ref.abs().max())
    # (m contiguous) destination [N, M]: the layout of dW[n_out, k_in] with U = X, V = dY
    dw = torch.zeros(N, M, device=cuda)
    db = torch.zeros(N, device=cuda)
    ops.wgrad_bf16(u, v, dw, 1, M, colsum=db, colsum_of_v=True)
    assert float((dw.double().t() - ref).abs().max()) <= 1e-4 * scale
    assert float((db.double() - v.double().sum(0)).abs().max()) <= 1e-4 * float(v.double().sum(0).abs().max() + 1)
    # (n contiguous) destination [M, N], accumulating on top of existing content, bias from U
    dw2 = torch.ones(M, N, device=cuda)
    db2 = torch.zeros(M, device=cuda)
    ops.wgrad_bf16(u, v, dw2, N, 1, colsum=db2, colsum_of_v=False)
    assert float((dw2.double() - 1 - ref).abs().max()) <= 1e-4 * scale
    assert float((db2.double() - u.double().sum(0)).abs().max()) <= 1e-4 * float(u.double().sum(0).abs().max() + 1)
    # column slice of a wider tensor (row pitch > width), no bias
    wide = torch.randn(P, N + 64, generator=g).to(torch.bfloat16).to(cuda)
    dw3 = torch.zeros(M, N, device=cuda)
    ops.wgrad_bf16(u, wide[:, 64:], dw3, N, 1)
    assert float((dw3.double() - u.double().t() @ wide[:, 64:].double()).abs().max()) <= 1e-4 * scale


@pytest.mark.parametrize("P,N,colsum_of_v,n_valid", [(64, 256, True, 256), (1000, 256, False, 200), (70001, 256, True, 256),
                                                     (4097, 64, True, 63), (70001, 64, False, 63), (5, 64, True, 4)])
def test_wgrad_cta_pair_body_vs_single_cta(cuda, monkeypatch, P, N, colsum_of_v, n_valid):
    """The CTA-pair weight-gradient body (wgrad_pair_body.cuh: cta_group::2 MMAs, bias gradient as an N = 16 MMA against
    a block of ones, N = 64 jobs as N = 128 with an out-of-bounds zero box) launched on its own (NFS_WGRAD_PAIR=1)
    against the single-CTA kernel and against fp64: it is the consumer side of nfs_mlp_backward_fused and the body of
    nfs_wgrad_multi_bf16's 256-wide jobs."""
    from helpers import record
    from nfs_b200 import ops
    g = torch.Generator().manual_seed(P + N)
    u = torch.randn(P, 256, generator=g).to(torch.bfloat16).to(cuda)
    v = torch.randn(P, N, generator=g).to(torch.bfloat16).to(cuda)
    res = {}
    for pair in ("0", "1"):
        if pair == "1":
            monkeypatch.setenv("NFS_WGRAD_PAIR", "1")
        else:
            monkeypatch.delenv("NFS_WGRAD_PAIR", raising=False)
        dw = torch.full((N, 256), 0.5, device=cuda)                  # accumulates on top of existing content
        db = torch.zeros(256, device=cuda)
        ops.wgrad_bf16(u, v, dw, 1, 256, colsum=db, colsum_of_v=colsum_of_v, n_valid=n_valid)
        torch.cuda.synchronize()
        res[pair] = (dw.double() - 0.5, db.double())
    ref = v.double().t() @ u.double()
    ref[n_valid:] = 0
    cref = torch.zeros(256, dtype=torch.float64, device=cuda)
    if colsum_of_v:
        cref[:n_valid] = v.double().sum(0)[:n_valid]
    else:
        cref = u.double().sum(0)
    scale, cscale = float(ref.abs().max()) + 1e-30, float(cref.abs().max()) + 1.0
    for pair in ("0", "1"):
        assert float((res[pair][0] - ref).abs().max()) <= 1e-4 * scale, pair
        assert float((res[pair][1] - cref).abs().max()) <= 1e-4 * cscale, pair
    record("wgrad_pair_vs_single", P=P, N=N, dw_rel=float((res["1"][0] - res["0"][0]).abs().max() / scale),
           colsum_rel=float((res["1"][1] - res["0"][1]).abs().max() / cscale))


def _g1_case(cuda, kwargs, P, seed):
    """Oracle (fp32, CPU) vs drop-in module (bf16 tensor cores) from the same seed / state_dict."""
    from models.nerf_model import NeRFMLP
    from oracle import nerf_oracle as O
    torch.manual_seed(seed)
    ref = O.PlainNeRF(**kwargs)
    mod = NeRFMLP(**kwargs)
    mod.load_state_dict(ref.state_dict())          # names / shapes interchange with the reference
    mod = mod.to(cuda)
    g = torch.Generator().manual_seed(seed + 1)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 6
    enc = O.encode(pts, O.frequency_bands((kwargs.get("pos_dim", 63) // 3 - 1) // 2))
    tgt = torch.rand(P, 4, generator=g)
    out_ref = ref(enc)
    loss_ref = ((out_ref - tgt) ** 2).mean()
    g_ref = torch.autograd.grad(loss_ref, list(ref.parameters()))
    out = mod(enc.to(cuda))
    loss = ((out - tgt.to(cuda)) ** 2).mean()
    grads = torch.autograd.grad(loss, list(mod.parameters()))
    return ref, mod, pts, enc, out_ref.detach(), out.detach().cpu(), g_ref, [t.cpu() for t in grads]


class _RoundBF16(torch.autograd.Function):
    """x -> bf16(x) in the forward, g -> bf16(g) in the backward: what a tensor travelling between two GEMMs as a bf16
    operand undergoes in either direction."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _torch_bf16_baseline(ref, enc, tgt):
    """The SAME model evaluated by plain PyTorch with bf16 GEMM operands and fp32 accumulation (weights, activations and
    back-propagated gradients rounded to bf16 where they enter a GEMM; biases, activations' math and reductions in
    fp32) - the numerical floor any bf16-operand implementation shares.  Returns its parameter gradients."""
    import torch.nn.functional as F
    rb = _RoundBF16.apply
    h = rb(enc)
    for lin in ref.layers:
        h = rb(F.relu(F.linear(h, rb(lin.weight), lin.bias)))
    out = torch.cat([torch.sigmoid(F.linear(h, rb(ref.rgb_out.weight), ref.rgb_out.bias)),
                     F.linear(h, rb(ref.sigma_out.weight), ref.sigma_out.bias)], -1)
    return torch.autograd.grad(((out - tgt) ** 2).mean(), list(ref.parameters()))


@pytest.mark.parametrize("kwargs,P", [(dict(), 4096), (dict(pos_dim=75), 1000), (dict(pos_dim=63, hidden_dim=128, n_layers=4), 777),
                                      (dict(pos_dim=27, hidden_dim=64, n_layers=2), 130)])
def test_g1_forward_backward_vs_oracle(cuda, kwargs, P):
    """bf16 operands / fp32 accumulation against the fp32 oracle at random init.
    Stated tolerance: rgb abs <= 2e-2, sigma rel <= 3e-2 + abs 1e-2;
    parameter gradients: relative L2 error <= 8e-2 per tensor (measured: 2.2e-2 at L=10,
    5.2e-2 on the first layer at L=12 - dY travels between layers in bf16, 8 mantissa bits)."""
    from helpers import record
    ref, mod, pts, enc, out_ref, out, g_ref, grads = _g1_case(cuda, kwargs, P, seed=11)
    e_rgb = float((out[:, :3] - out_ref[:, :3]).abs().max())
    e_sig = float(((out[:, 3] - out_ref[:, 3]).abs() / (3e-2 * out_ref[:, 3].abs() + 1e-2)).max())
    rels = [float((a - b).norm() / b.norm().clamp_min(1e-12)) for a, b in zip(grads, g_ref)]
    # what the bound means: plain PyTorch with bf16 GEMM operands (fp32 accumulate) on the same model / batch
    g = torch.Generator().manual_seed(11 + 1)
    _ = torch.rand(P, 3, generator=g)
    tgt = torch.rand(P, 4, generator=g)
    base = _torch_bf16_baseline(ref, enc, tgt)
    rels_base = [float((a - b).norm() / b.norm().clamp_min(1e-12)) for a, b in zip(base, g_ref)]
    record("g1_vs_oracle", kwargs=str(kwargs), rgb_abs=e_rgb, sigma_score=e_sig, grad_rel_l2_max=max(rels),
           torch_bf16_baseline_grad_rel_l2_max=max(rels_base))
    assert e_rgb <= 2e-2 and e_sig <= 1.0, (e_rgb, e_sig)
    assert max(rels) <= 8e-2, rels
    assert max(rels) <= 2.0 * max(rels_base) + 5e-3, (max(rels), max(rels_base))      # no worse than bf16 operands imply
    # fused-encoding entry point == encode-then-forward
    if kwargs.get("pos_dim", 63) % 3 == 0:
        L = (kwargs.get("pos_dim", 63) // 3 - 1) // 2
        from oracle import nerf_oracle as O
        with torch.no_grad():
            fused = mod.forward_points(pts.to(cuda), O.frequency_bands(L)).cpu()
        assert float((fused[:, :3] - out_ref[:, :3]).abs().max()) <= 2e-2


def test_g1_golden(golden, cuda):
    from models.nerf_model import NeRFMLP
    from oracle import nerf_oracle as O
    for c in golden("mlp"):
        if c["kind"] != "g1":
            continue
        torch.manual_seed(c["seed"])
        ref = O.PlainNeRF(**c["kwargs"])
        mod = NeRFMLP(**c["kwargs"])
        mod.load_state_dict(ref.state_dict())
        mod = mod.to(cuda)
        with torch.no_grad():
            out = mod(c["x"].to(cuda)).cpu()
        assert out.shape == c["out"].shape
        assert float((out[:, :3] - c["out"][:, :3]).abs().max()) <= 2e-2
        assert bool(((out[:, 3] - c["out"][:, 3]).abs() <= 3e-2 * c["out"][:, 3].abs() + 1e-2).all())


def test_g1_training_tracks_fp32(cuda):
    """100 Adam steps on a fixed batch: the tensor-core model and the fp32 oracle reach the same loss."""
    from models.nerf_model import NeRFMLP
    from oracle import nerf_oracle as O
    torch.manual_seed(3)
    ref = O.PlainNeRF(hidden_dim=128, n_layers=4)
    mod = NeRFMLP(hidden_dim=128, n_layers=4)
    mod.load_state_dict(ref.state_dict())
    mod = mod.to(cuda)
    g = torch.Generator().manual_seed(4)
    enc = O.encode((torch.rand(2048, 3, generator=g) - 0.5) * 4, O.frequency_bands(10))
    tgt = torch.rand(2048, 4, generator=g)
    o1, o2 = torch.optim.Adam(ref.parameters(), 1e-3), torch.optim.Adam(mod.parameters(), 1e-3)
    enc_c, tgt_c = enc.to(cuda), tgt.to(cuda)
    for _ in range(100):
        o1.zero_grad(); l1 = ((ref(enc) - tgt) ** 2).mean(); l1.backward(); o1.step()
        o2.zero_grad(); l2 = ((mod(enc_c) - tgt_c) ** 2).mean(); l2.backward(); o2.step()
    assert abs(float(l1) - float(l2)) <= 0.05 * float(l1), (float(l1), float(l2))
    with torch.no_grad():
        out = mod(enc_c).cpu()
    assert float((out[:, :3] - ref(enc)[:, :3]).abs().max()) <= 5e-2


@pytest.mark.parametrize("P", [1, 127, 128, 129, 256, 257, 40000, 148 * 256 * 3 + 5])
def test_fused_chain_matches_layerwise(cuda, P, monkeypatch):
    """nfs_mlp_chain (one launch, activations on chip) against the layer-by-layer launches on the same bf16
    operands.  The chain adds the bias on the tensor core (three bf16 terms, ~2^-24 relative) and the layer kernel
    adds the fp32 bias in its epilogue, so pre-activations differ by ~1e-7 relative and a few elements per thousand
    land on the other side of a bf16 rounding boundary: activations agree to 2 bf16 ulps (+ 4e-3 of the largest, for the flips propagated from earlier layers) with
    < 1 % of the elements differing at all (measured 0.08 %) and a relative L2 difference <= 1e-3 (measured 1.4e-4); the fp32 outputs to 1e-2 relative.  The two backward routes get the SAME saved activations,
    so they stay equal up to the order of the fp32 atomics in wgrad."""
    from models.nerf_model import NeRFMLP
    torch.manual_seed(2)
    mod = NeRFMLP().to(cuda)
    plan = mod._get_plan()
    plan.refresh()
    g = torch.Generator().manual_seed(P)
    x16 = torch.randn(P, 64, generator=g).to(torch.bfloat16).to(cuda)
    x16[:, 63] = 0
    out_f, acts_f, save_f = plan.run_forward_fused(x16, keep=True)
    monkeypatch.setenv("NFS_MLP_FUSED", "0")
    out_l, acts_l, _ = plan.run_forward(x16, keep=True)
    assert len(acts_f) == len(acts_l) == 9
    assert torch.equal(acts_f[0], acts_l[0])
    for i, (a, b) in enumerate(zip(acts_f, acts_l)):
        assert a.shape == b.shape
        a, b = a.float(), b.float()
        assert bool(((a - b).abs() <= 2.0 ** -7 * b.abs() + 4e-3 * float(b.abs().max())).all()), "activation %d differs" % i
        assert float((a != b).float().mean()) < 0.01, "activation %d: too many elements differ" % i
        assert float((a - b).norm()) <= 1e-3 * float(b.norm()), "activation %d: relative L2" % i
    assert bool(((out_f - out_l).abs() <= 1e-2 * out_l.abs() + 1e-3).all())
    out_n, acts_n, _ = plan.run_forward_fused(x16, keep=False)     # inference: nothing saved, same arithmetic
    assert torch.equal(out_n, out_f) and len(acts_n) == 1
    g_out = torch.randn(P, 4, generator=g).to(cuda)
    grads_l = plan.run_backward(acts_f, out_f, g_out, save_fwd=None)
    grads_f = plan.run_backward(acts_f, out_f, g_out, save_fwd=save_f)
    for a, b in zip(grads_f, grads_l):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max() + 1e-12) + 1e-7


@pytest.mark.parametrize("P", [1, 129, 5000, 70001])
def test_in_kernel_encoding_matches_two_kernel_route(cuda, P, monkeypatch):
    """nfs_mlp_chain_points (K2 fused into the chain) against posenc_bf16 -> nfs_mlp_chain: identical
    arithmetic, so bit-identical outputs; both within tolerance of the fp32 oracle."""
    from models.nerf_model import NeRFMLP
    from oracle import nerf_oracle as O
    torch.manual_seed(3)
    ref = O.PlainNeRF()
    mod = NeRFMLP()
    mod.load_state_dict(ref.state_dict())
    mod = mod.to(cuda)
    g = torch.Generator().manual_seed(P)
    pts = (torch.rand(P, 3, generator=g) - 0.5) * 8
    bands = O.frequency_bands(10)
    with torch.no_grad():
        fused = mod.forward_points(pts.to(cuda), bands)
        monkeypatch.setenv("NFS_MLP_FUSED_ENC", "0")
        two = mod.forward_points(pts.to(cuda), bands)
        monkeypatch.delenv("NFS_MLP_FUSED_ENC")
        out_ref = ref(O.encode(pts, bands))
    assert torch.equal(fused, two)
    assert float((fused.cpu()[:, :3] - out_ref[:, :3]).abs().max()) <= 2e-2
    assert float(((fused.cpu()[:, 3] - out_ref[:, 3]).abs() / (3e-2 * out_ref[:, 3].abs() + 1e-2)).max()) <= 1.0


@pytest.mark.parametrize("P", [1, 129, 5000, 70001])
def test_in_kernel_encoding_training_forward(cuda, P, monkeypatch):
    """nfs_mlp_chain_points_train (training forward with K2 fused in: the kernel stores the encoded operand for
    layer 0's weight gradient) against posenc_bf16 -> nfs_mlp_chain: bit-identical outputs, the stored operand
    equals the encoding kernel's, parameter gradients equal up to the order of wgrad's fp32 atomics."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import mlp
    from oracle import nerf_oracle as O
    torch.manual_seed(3)
    mod = NeRFMLP().to(cuda)
    g = torch.Generator().manual_seed(P)
    pts = ((torch.rand(P, 3, generator=g) - 0.5) * 8).to(cuda)
    g_out = torch.randn(P, 4, generator=g).to(cuda)
    bands = O.frequency_bands(10)
    res = []
    for enc in ("1", "0"):
        monkeypatch.setenv("NFS_MLP_FUSED_ENC", enc)
        mod.zero_grad()
        out = mod.forward_points(pts, bands)
        out.backward(g_out)
        res.append((out.detach().clone(), [p.grad.clone() for p in mod.parameters()]))
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max() + 1e-12) + 1e-7
    # the operand the kernel stored == the encoding kernel's operand
    plan = mod._get_plan()
    monkeypatch.setenv("NFS_MLP_FUSED_ENC", "1")
    x_two = mlp.encode_operand(pts, bands, plan.k0)
    x_in = torch.zeros(((P + 127) // 128 * 128, plan.k0), device=cuda, dtype=torch.bfloat16)
    plan.run_forward_fused(x_in[:P], True, points=pts, freqs=bands)
    assert torch.equal(x_in[:P], x_two)


def test_chain_stress_repeatable(cuda):
    """The CTA-pair chain hands tiles between two SMs through mbarriers and remote arrives: a protocol race would
    show up as run-to-run differences.  Repeat inference, training forward (activations + sign bits) and the dgrad
    chain on many sizes and demand bit-identical results every time."""
    from models.nerf_model import NeRFMLP
    torch.manual_seed(4)
    mod = NeRFMLP().to(cuda)
    plan = mod._get_plan()
    plan.refresh()
    g = torch.Generator().manual_seed(0)
    for P in (3, 500, 513, 9999, 74 * 512 + 1, 148 * 512 * 2 + 77):
        x16 = torch.randn(P, 64, generator=g).to(torch.bfloat16).to(cuda)
        dy = torch.randn(P, 64, generator=g).to(torch.bfloat16).to(cuda)
        ref_out, _, ref_save = plan.run_forward_fused(x16, keep=True)
        ref_dys = plan.dgrad_chain_fused(dy, ref_save[1], P)
        ref_inf = plan.run_forward_fused(x16, keep=False)[0]
        assert torch.equal(ref_inf, ref_out)
        for _ in range(12):
            out, _, save = plan.run_forward_fused(x16, keep=True)
            assert torch.equal(out, ref_out)
            assert torch.equal(save[0][:, :P], ref_save[0][:, :P]) and torch.equal(save[1][:, :P], ref_save[1][:, :P])
            assert torch.equal(plan.dgrad_chain_fused(dy, save[1], P)[:, :P], ref_dys[:, :P])
            assert torch.equal(plan.run_forward_fused(x16, keep=False)[0], ref_out)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [1, 300, 5000, 40001])
def test_chain_mixed_widths_softmax_head_and_dgrad(cuda, P):
    """Chains whose layers differ in width (NeRFDINOFusion's attention branch, dino_feature_model.py:165-170,185-188):
    forward [192 -> 256 -> 256 -> 128 -> 2-way softmax head] with saved activations and sign bits, then its dgrad chain
    [64 -> 128 -> 256 -> 256] masked by those bits - against bf16-operand / fp32-accumulate torch arithmetic.  The saved
    tensor is as wide as the widest layer; the 128-wide layer fills the first 128 columns of its rows."""
    import ctypes
    from nfs_b200 import _lib
    from nfs_b200._lib import ptr
    from nfs_b200.mlp import bias_terms
    from nfs_b200.ops import _stream
    from helpers import record
    g = torch.Generator().manual_seed(P)
    dims = [(192, 256), (256, 256), (256, 128), (128, 64)]
    Ws = [(torch.randn(n, k, generator=g) * (2.0 / k) ** 0.5).to(torch.bfloat16) for k, n in dims]
    bs = [torch.randn(n, generator=g) * 0.1 for k, n in dims]
    Ws[3][2:] = 0; bs[3][2:] = 0
    x = torch.randn(P, 192, generator=g).to(torch.bfloat16)
    i32 = lambda v: (ctypes.c_int32 * len(v))(*v)
    rows = (P + 127) // 128 * 128

    def stack(mats):
        w = torch.zeros(sum(m.shape[0] for m in mats), 256, dtype=torch.bfloat16)
        r, row0 = 0, []
        for m in mats:
            w[r:r + m.shape[0], :m.shape[1]] = m
            row0.append(r); r += m.shape[0]
        return w.to(cuda), row0

    w, row0 = stack(Ws)
    bt = bias_terms(torch.cat(bs).to(cuda))
    save = torch.full((3, rows, 256), float("nan"), device=cuda, dtype=torch.bfloat16)
    bits = torch.zeros((3, rows, 8), device=cuda, dtype=torch.int32)
    gate = torch.empty((P, 2), device=cuda, dtype=torch.float32)
    xc = x.to(cuda)
    _lib.call("nfs_mlp_chain", ptr(xc), P, 4, i32([k for k, n in dims]), i32([n for k, n in dims]), i32([1, 1, 1, 5]),
              i32(row0), ptr(w), w.shape[0], ptr(bt), None, 0, None, ptr(save), ptr(bits), rows, ptr(gate), 2, _stream())
    # reference: bf16 operands, fp32 accumulation, bf16 activations between layers
    h, acts = x.float(), []
    for l in range(3):
        h = torch.relu(h @ Ws[l].float().T + bs[l]).to(torch.bfloat16).float()
        acts.append(h)
    ref_gate = torch.softmax((h @ Ws[3].float().T + bs[3])[:, :2], dim=-1)
    for l, n in enumerate((256, 256, 128)):
        got = save[l, :P, :n].float().cpu()
        assert torch.isfinite(got).all()
        assert (got - acts[l]).abs().max() <= 0.02 * acts[l].abs().max() + 1e-3, l
    assert torch.isnan(save[2, :P, 128:].float()).all()          # the narrow layer wrote only its own columns
    assert (gate.cpu() - ref_gate).abs().max() < 5e-3
    record("chain_mixed_widths", P=P, gate_err=float((gate.cpu() - ref_gate).abs().max()))

    # dgrad chain: dlog [P,64] -> d a (128) -> d h2 (256) -> d h1 (256), each masked by the sign of the saved layer
    dlog = torch.zeros(P, 64, dtype=torch.bfloat16)
    dlog[:, :2] = torch.randn(P, 2, generator=g).to(torch.bfloat16)
    Wt = [Ws[3].T.contiguous(), Ws[2].T.contiguous(), Ws[1].T.contiguous()]        # [128,64], [256,128], [256,256]
    wt, row0t = stack(Wt)
    dys = torch.full((3, rows, 256), float("nan"), device=cuda, dtype=torch.bfloat16)
    dc = dlog.to(cuda)
    _lib.call("nfs_mlp_chain", ptr(dc), P, 3, i32([64, 128, 256]), i32([128, 256, 256]), i32([4, 4, 4]), i32(row0t),
              ptr(wt), wt.shape[0], None, ptr(bits), rows, i32([2, 1, 0]), ptr(dys), None, rows, None, 0, _stream())
    saved_cpu = [save[l, :P].float().cpu() for l in range(3)]
    d = dlog.float()
    for t, (wmat, src, n) in enumerate(zip(Wt, (2, 1, 0), (128, 256, 256))):
        d = ((d @ wmat.float().T) * (saved_cpu[src][:, :n] > 0)).to(torch.bfloat16).float()
        got = dys[t, :P, :n].float().cpu()
        assert torch.isfinite(got).all()
        assert (got - d).abs().max() <= 0.02 * d.abs().max() + 1e-4, t
        d = got                                                   # follow the kernel's own rounding downstream
