"""Tolerance helpers shared by the parity tests."""
import torch


def rel_err(got, ref, floor=0.0):
    """max elementwise |got-ref| / max(|ref|, floor)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    den = ref.abs().clamp_min(floor if floor > 0 else 1e-30)
    return float(((got - ref).abs() / den).max()) if ref.numel() else 0.0


def per_ray_err(got, ref, floor=0.0, floor_frac=0.0):
    """max over rays of max_s|got-ref| / max(max_s|ref|, floor, floor_frac * max|ref|) - the
    measure for (N,S[,C]) outputs and gradients, whose tiny entries suffer cancellation in
    1-alpha.  `floor` (absolute) / `floor_frac` (fraction of the batch
    maximum) bound the denominator for rays whose whole row is negligible."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    if ref.numel() == 0:
        return 0.0
    n = ref.shape[0]
    d = (got - ref).abs().reshape(n, -1).max(dim=1).values
    lo = max(floor, floor_frac * float(ref.abs().max()), 1e-30)
    s = ref.abs().reshape(n, -1).max(dim=1).values.clamp_min(lo)
    return float((d / s).max())


def record(name, **values):
    """Append measured parity numbers to gpurun_out/parity_metrics.jsonl (when that directory
    exists) so the figures quoted in DESIGN.md / profiles/ come from the test run itself."""
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_metrics.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **values)) + "\n")


def bit_equal(a, b):
    a, b = a.cpu().contiguous(), b.cpu().contiguous()
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == torch.float32:
        return bool(torch.equal(a.view(torch.int32), b.view(torch.int32)))
    return bool(torch.equal(a, b))


def ulp_diff(a, b):
    """max distance in units of last place between two fp32 tensors of equal sign pattern."""
    ai = a.cpu().contiguous().view(torch.int32).long()
    bi = b.cpu().contiguous().view(torch.int32).long()
    return int((ai - bi).abs().max()) if ai.numel() else 0
