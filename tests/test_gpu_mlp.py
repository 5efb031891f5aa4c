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
