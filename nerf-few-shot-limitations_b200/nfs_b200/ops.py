"""Host-side operators over the C ABI (include/nfs_b200.h): argument checking,
output allocation on the input's device, launch on torch's current stream, and
the autograd glue.  Nothing here computes on the CPU; every function raises if the
tensors are not CUDA tensors or the library is missing.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ptr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(name, *tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "%s: expected CUDA tensors (got device %s). nfs_b200 has no CPU fallback; the "
                "reference's CPU path lives in the reference itself." % (name, t.device))


def _f32c(t):
    """fp32 + contiguous (no copy when already so)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# --------------------------------------------------------------------------- K1
class _CompositeFn(torch.autograd.Function):
    """VolumeRenderer.forward (nerf_mlp.py:165-215) / volume_render_radiance
    (volume_renderer.py:4-43) as one kernel per direction.  Nothing but the inputs
    is saved: the backward recomputes alpha / T (include/nfs_b200.h)."""

    @staticmethod
    def forward(ctx, rgb, density, z_vals, rays_d, noise, noise_std, white_bkgd, packed, want_aux):
        n_samples = z_vals.shape[-1]
        n_rays = z_vals.numel() // max(n_samples, 1)
        dev = z_vals.device
        out_rgb = torch.empty((n_rays, 3), device=dev, dtype=torch.float32)
        out_depth = torch.empty((n_rays,), device=dev, dtype=torch.float32) if want_aux else None
        out_w = torch.empty((n_rays, n_samples), device=dev, dtype=torch.float32) if want_aux else None
        with torch.cuda.device(dev):
            _lib.call("nfs_composite_fwd", ptr(rgb), ptr(density), ptr(z_vals), ptr(rays_d), ptr(noise),
                      float(noise_std), n_rays, n_samples, int(bool(white_bkgd)), int(bool(packed)),
                      ptr(out_rgb), ptr(out_depth), ptr(out_w), _stream())
        ctx.save_for_backward(rgb, density, z_vals, rays_d, noise)
        ctx.cfg = (float(noise_std), int(bool(white_bkgd)), int(bool(packed)), n_rays, n_samples)
        ctx.set_materialize_grads(False)
        if want_aux:
            return out_rgb, out_depth, out_w
        return out_rgb, None, None

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_w):
        rgb, density, z_vals, rays_d, noise = ctx.saved_tensors
        noise_std, white, packed, n_rays, n_samples = ctx.cfg
        dev = z_vals.device
        if g_rgb is None:
            g_rgb = torch.zeros((n_rays, 3), device=dev, dtype=torch.float32)
        g_rgb, g_depth, g_w = _f32c(g_rgb), _f32c(g_depth), _f32c(g_w)
        d_rgb = torch.empty_like(rgb)
        d_density = None if packed else torch.empty_like(density)
        with torch.cuda.device(dev):
            _lib.call("nfs_composite_bwd", ptr(rgb), ptr(density), ptr(z_vals), ptr(rays_d), ptr(noise),
                      noise_std, ptr(g_rgb), ptr(g_depth), ptr(g_w), n_rays, n_samples, white, packed,
                      ptr(d_rgb), ptr(d_density), _stream())
        return d_rgb, d_density, None, None, None, None, None, None, None


def _no_geometry_grads(who, z_vals, rays_d):
    """The reference's autograd also differentiates through dists = diff(z_vals) * ||rays_d|| and depth = sum w z
    (nerf_mlp.py:181-185,208); the CUDA backward returns gradients for rgb / density only.  No caller of the reference
    asks for the others - if one does (learned poses, differentiable resampling) fail loudly instead of returning zeros."""
    if torch.is_grad_enabled() and (z_vals.requires_grad or rays_d.requires_grad):
        raise RuntimeError("%s: gradients w.r.t. z_vals / rays_d are not implemented by the CUDA compositing backward "
                           "(detach them, or see INTEGRATION.md 'Differentiability')" % who)


def composite(rgb, density, z_vals, rays_d, noise=None, noise_std=0.0, white_bkgd=False):
    """rgb (...,S,3), density (...,S,1)|(...,S), z_vals (...,S), rays_d (...,3)
    -> rgb_map (...,3), depth (...), weights (...,S)."""
    _need_cuda("composite", rgb, density, z_vals, rays_d, noise)
    _no_geometry_grads("composite", z_vals, rays_d)
    lead = z_vals.shape[:-1]
    S = z_vals.shape[-1]
    if density.dim() == z_vals.dim() + 1:
        if density.shape[-1] != 1:
            raise RuntimeError("composite: density must be (...,S,1) or (...,S)")
        density = density.reshape(*lead, S)
    if rgb.shape[:-1] != z_vals.shape or rgb.shape[-1] != 3 or density.shape != z_vals.shape \
            or rays_d.shape != (*lead, 3):
        raise RuntimeError("composite: shape mismatch rgb %s density %s z_vals %s rays_d %s" % (
            tuple(rgb.shape), tuple(density.shape), tuple(z_vals.shape), tuple(rays_d.shape)))
    if S == 0:
        raise RuntimeError("composite: n_samples must be positive")
    if S == 1:
        # Reference quirk kept: with one sample `dists[..., :1]` of the empty difference is
        # itself empty (nerf_mlp.py:181-182), so nothing is composited: rgb = 0 (1 on white),
        # depth = 0, weights has shape (..., 0) and every gradient is zero.
        zero = rgb.sum(dim=-2) * 0.0 + (density.sum(dim=-1) * 0.0)[..., None]
        return zero + (1.0 if white_bkgd else 0.0), zero[..., 0], density[..., :0] * 0.0
    o_rgb, o_depth, o_w = _CompositeFn.apply(_f32c(rgb), _f32c(density), _f32c(z_vals), _f32c(rays_d),
                                             _f32c(noise), noise_std, white_bkgd, False, True)
    return o_rgb.reshape(*lead, 3), o_depth.reshape(lead), o_w.reshape(*lead, S)


def composite_packed(rgb_sigma, z_vals, rays_d, noise=None, noise_std=0.0, white_bkgd=False, want_aux=False):
    """rgb_sigma (...,S,4), z_vals (...,S), rays_d (...,3) -> rgb_map (...,3)
    [, depth (...), weights (...,S) when want_aux]."""
    _need_cuda("composite_packed", rgb_sigma, z_vals, rays_d, noise)
    _no_geometry_grads("composite_packed", z_vals, rays_d)
    lead = z_vals.shape[:-1]
    if rgb_sigma.shape[:-1] != z_vals.shape or rgb_sigma.shape[-1] != 4 or rays_d.shape != (*lead, 3):
        raise RuntimeError("composite_packed: shape mismatch rgb_sigma %s z_vals %s rays_d %s" % (
            tuple(rgb_sigma.shape), tuple(z_vals.shape), tuple(rays_d.shape)))
    if z_vals.shape[-1] == 0:
        raise RuntimeError("composite_packed: n_samples must be positive")
    if z_vals.shape[-1] == 1:   # same reference quirk as in composite(): empty dists, rgb_map = 0
        return rgb_sigma.sum(dim=-2)[..., :3] * 0.0
    o_rgb, o_depth, o_w = _CompositeFn.apply(_f32c(rgb_sigma), None, _f32c(z_vals), _f32c(rays_d), _f32c(noise),
                                             noise_std, white_bkgd, True, want_aux)
    if want_aux:
        return o_rgb.reshape(*lead, 3), o_depth.reshape(lead), o_w.reshape(*lead, z_vals.shape[-1])
    return o_rgb.reshape(*lead, 3)


class _CompositeLossFn(torch.autograd.Function):
    """Compositing with the loss in its epilogue (nfs_composite_loss_fwd): the only differentiable
    output is the loss; the kernel has already written d loss / d rgb_map (and d depth), so the backward
    is nfs_composite_bwd on those, scaled by the incoming gradient of the loss."""

    @staticmethod
    def forward(ctx, rgb, density, z_vals, rays_d, target, target_depth, rgb_weight, depth_weight, white_bkgd,
                packed, want_weights, dy_slot=None):
        n_samples = z_vals.shape[-1]
        n_rays = z_vals.numel() // n_samples
        dev = z_vals.device
        out_rgb = torch.empty((n_rays, 3), device=dev, dtype=torch.float32)
        out_depth = torch.empty((n_rays,), device=dev, dtype=torch.float32)
        out_w = torch.empty((n_rays, n_samples), device=dev, dtype=torch.float32) if want_weights else None
        g_rgb = torch.empty((n_rays, 3), device=dev, dtype=torch.float32)
        g_depth = torch.empty((n_rays,), device=dev, dtype=torch.float32) if target_depth is not None else None
        sums = torch.zeros((32, 2), device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            _lib.call("nfs_composite_loss_fwd", ptr(rgb), ptr(density), ptr(z_vals), ptr(rays_d), None, 0.0,
                      ptr(target), ptr(target_depth), float(rgb_weight), float(depth_weight), n_rays, n_samples,
                      int(bool(white_bkgd)), int(bool(packed)), ptr(out_rgb), ptr(out_depth), ptr(out_w),
                      ptr(g_rgb), ptr(g_depth), ptr(sums), _stream())
        terms = sums.sum(0)                                   # [sum sq err, sum abs depth err] (fp64)
        mse = (terms[0] / (3.0 * n_rays)).float()
        l1 = (terms[1] / float(n_rays)).float() if target_depth is not None else None
        loss = rgb_weight * mse if l1 is None else rgb_weight * mse + depth_weight * l1
        ctx.save_for_backward(rgb, density, z_vals, rays_d, g_rgb, g_depth)
        ctx.cfg = (int(bool(white_bkgd)), int(bool(packed)), n_rays, n_samples)
        ctx.dy_slot = dy_slot if packed else None
        outs = (loss, mse, l1, out_rgb, out_depth, out_w)
        ctx.mark_non_differentiable(*[o for o in outs[1:] if o is not None])
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        rgb, density, z_vals, rays_d, g_rgb, g_depth = ctx.saved_tensors
        white, packed, n_rays, n_samples = ctx.cfg
        none = (None,) * 10
        if g_loss is None:
            return (None, None) + none
        g_rgb = g_rgb * g_loss
        if g_depth is not None:
            g_depth = g_depth * g_loss
        if ctx.dy_slot is not None:
            # `rgb` is the packed output of an MLP call that runs inside a mlp.StepSession: write the gradient of the
            # head's pre-activations straight into the session's bf16 operand (nfs_composite_bwd_dy) - no fp32
            # d(rgb_sigma) tensor, no nfs_act_grad_bf16 launch.  The MLP's backward only needs to be triggered: it
            # receives a stride-0 placeholder.
            sess, r0 = ctx.dy_slot
            P = n_rays * n_samples
            if sess.defer_composite_backward(r0, rgb, z_vals, rays_d, g_rgb, g_depth, n_rays, n_samples, white):
                # merged route: this launch happens inside nfs_render_fused_bwd at the session's flush, together with
                # the MLP backward - the step's whole backward is one C call
                sess.dy_written.add(r0)
                return (_placeholder(rgb), None) + none
            with torch.cuda.device(z_vals.device):
                _lib.call("nfs_composite_bwd_dy", ptr(rgb), ptr(z_vals), ptr(rays_d), ptr(g_rgb), ptr(g_depth), None, n_rays,
                          n_samples, white, ptr(sess.dy[r0:r0 + P]), int(sess.dy.stride(0)), _stream())
            sess.dy_written.add(r0)
            return (_placeholder(rgb), None) + none
        d_rgb = torch.empty_like(rgb)
        d_density = None if packed else torch.empty_like(density)
        with torch.cuda.device(z_vals.device):
            _lib.call("nfs_composite_bwd", ptr(rgb), ptr(density), ptr(z_vals), ptr(rays_d), None, 0.0, ptr(g_rgb),
                      ptr(g_depth), None, n_rays, n_samples, white, packed, ptr(d_rgb), ptr(d_density), _stream())
        return (d_rgb, d_density) + none


_PLACEHOLDERS = {}


def _placeholder(like):
    """A zero 'gradient' of like's shape that owns one element (stride 0 everywhere)."""
    key = (str(like.device), like.dtype)
    z = _PLACEHOLDERS.get(key)
    if z is None:
        z = _PLACEHOLDERS[key] = torch.zeros((), device=like.device, dtype=like.dtype)
    return z.expand(like.shape)


def composite_loss(rgb, density, z_vals, rays_d, target_rgb, target_depth=None, rgb_weight=1.0, depth_weight=0.1,
                   white_bkgd=False, want_weights=False, dy_slot=None):
    """VolumeRenderer.forward (nerf_mlp.py:165-215) followed by the rgb / depth terms of NeRFLoss
    (nerf_mlp.py:225-258; train.py:36-44 is the rgb term alone) in ONE kernel.

    rgb (N,S,3) + density (N,S,1)|(N,S), or rgb = packed (N,S,4) [r,g,b,sigma] with density None;
    z_vals (N,S); rays_d (N,3); target_rgb (N,3); target_depth (N)|None.
    Returns a dict: 'total' = rgb_weight * mse [+ depth_weight * l1] (the only differentiable entry),
    'rgb' = mse, ['depth' = l1,] and the detached renderings 'rgb_map', 'depth_map' [, 'weights'].
    The regularisation term of NeRFLoss (mean(weights^2)) is not fused: callers that use it composite with
    ops.composite and apply models.nerf_mlp.NeRFLoss."""
    packed = density is None
    _need_cuda("composite_loss", rgb, density, z_vals, rays_d, target_rgb, target_depth)
    _no_geometry_grads("composite_loss", z_vals, rays_d)
    if z_vals.dim() != 2 or z_vals.shape[1] < 2:
        raise RuntimeError("composite_loss: z_vals must be (N,S) with S >= 2")
    N, S = z_vals.shape
    if not packed and density.dim() == 3:
        density = density.reshape(N, S)
    if rgb.shape != (N, S, 4 if packed else 3) or (not packed and density.shape != (N, S)) or rays_d.shape != (N, 3) \
            or target_rgb.shape != (N, 3) or (target_depth is not None and target_depth.shape != (N,)):
        raise RuntimeError("composite_loss: shape mismatch")
    if N == 0:
        raise RuntimeError("composite_loss: empty batch (the mean of no pixels is undefined)")
    loss, mse, l1, o_rgb, o_depth, o_w = _CompositeLossFn.apply(
        _f32c(rgb), _f32c(density), _f32c(z_vals), _f32c(rays_d), _f32c(target_rgb), _f32c(target_depth),
        float(rgb_weight), float(depth_weight), white_bkgd, packed, want_weights, dy_slot)
    out = {"total": loss, "rgb": mse, "rgb_map": o_rgb, "depth_map": o_depth}
    if l1 is not None:
        out["depth"] = l1
    if o_w is not None:
        out["weights"] = o_w
    return out


# --------------------------------------------------------------------------- K6
def generate_rays(H, W, focal, c2w, pix_idx=None, image=None):
    """get_rays (ray_sampler.py:4-30 / ray_utils.py:4-37) for the pixels pix_idx (int64, row-major p = j*W + i;
    None = every pixel in order) plus the gather of their colours from image (H,W,3)|(H*W,3)|None - the batch
    assembly of train.py:272-278 - in one kernel.  Returns rays_o, rays_d (n,3) [, target (n,3)].
    Bit-exact with the reference's CPU arithmetic."""
    _need_cuda("generate_rays", c2w, pix_idx, image)
    if c2w.dim() != 2 or c2w.shape[0] < 3 or c2w.shape[1] != 4:
        raise RuntimeError("generate_rays: c2w must be (3,4) or (4,4)")
    c2w = _f32c(c2w)
    dev = c2w.device
    if pix_idx is None:
        n = H * W
    else:
        if pix_idx.dtype != torch.int64:
            pix_idx = pix_idx.long()
        pix_idx = pix_idx.contiguous().reshape(-1)
        n = pix_idx.numel()
    if image is not None:
        if image.numel() != H * W * 3:
            raise RuntimeError("generate_rays: image must hold H*W*3 values")
        image = _f32c(image)
    rays_o = torch.empty((n, 3), device=dev, dtype=torch.float32)
    rays_d = torch.empty((n, 3), device=dev, dtype=torch.float32)
    target = torch.empty((n, 3), device=dev, dtype=torch.float32) if image is not None else None
    if n == 0:
        return (rays_o, rays_d) if image is None else (rays_o, rays_d, target)
    with torch.cuda.device(dev):
        _lib.call("nfs_rays_generate", int(H), int(W), float(focal), ptr(c2w), 4, ptr(pix_idx), n, ptr(image),
                  ptr(rays_o), ptr(rays_d), ptr(target), _stream())
    return (rays_o, rays_d) if image is None else (rays_o, rays_d, target)


# --------------------------------------------------------------------------- K2
def posenc(x, freqs, include_input=True):
    """x (...,D) -> (..., D*(2L+include_input)); freqs: L fp32 values (any device)."""
    _need_cuda("posenc", x)
    x = _f32c(x)
    D = x.shape[-1]
    L = int(freqs.numel())
    P = x.numel() // max(D, 1)
    W = D * (2 * L + (1 if include_input else 0))
    out = torch.empty((*x.shape[:-1], W), device=x.device, dtype=torch.float32)
    if P == 0 or W == 0:
        return out
    fr = freqs.detach().to(device=x.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(x.device):
        _lib.call("nfs_posenc_fwd", ptr(x), ptr(fr), P, D, L, int(bool(include_input)), ptr(out), _stream())
    return out


# --------------------------------------------------------------------------- K4
def stratified_tables(near, far, n_samples, lindisp=False):
    """The S-entry tables of ray_utils.py:57-76 / ray_sampler.py:49-56, evaluated with
    the reference's own torch CPU arithmetic (torch.linspace on CPU is a vectorised,
    width-dependent formula - SURVEY.md section 8c - so it is never re-derived in a kernel).
    Returns (z_base, lower, upper), each (S,) fp32 on the CPU."""
    t_vals = torch.linspace(0.0, 1.0, steps=n_samples)
    if lindisp:
        z = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
    else:
        z = near * (1.0 - t_vals) + far * t_vals
    mids = 0.5 * (z[1:] + z[:-1])
    upper = torch.cat([mids, z[-1:]], -1)
    lower = torch.cat([z[:1], mids], -1)
    return z, lower, upper


_table_cache = {}


def _tables_on(device, near, far, n_samples, lindisp):
    key = (str(device), float(near), float(far), int(n_samples), bool(lindisp))
    hit = _table_cache.get(key)
    if hit is None:
        hit = tuple(t.to(device) for t in stratified_tables(near, far, n_samples, lindisp))
        if len(_table_cache) > 64:
            _table_cache.clear()
        _table_cache[key] = hit
    return hit


def importance_table(device, n_importance):
    """u = linspace(0, 1, N_importance) of ray_utils.py:113 (perturb=False), evaluated by torch on the
    CPU like the reference does and cached on `device`."""
    key = ("u", str(device), int(n_importance))
    hit = _table_cache.get(key)
    if hit is None:
        hit = _table_cache[key] = torch.linspace(0., 1., n_importance).to(device)
    return hit


def sample_stratified(rays_o, rays_d, near, far, n_samples, t_rand=None, lindisp=False, want_pts=True):
    """rays (...,3) -> pts (...,S,3), z (...,S).  t_rand (...,S) = the uniform draws of
    ray_utils.py:78 (None = perturb=False)."""
    _need_cuda("sample_stratified", rays_o, rays_d, t_rand)
    lead = rays_o.shape[:-1]
    if rays_o.shape[-1] != 3 or rays_d.shape != rays_o.shape:
        raise RuntimeError("sample_stratified: rays must be (...,3) with equal shapes")
    if n_samples <= 0:
        raise RuntimeError("sample_stratified: N_samples must be positive")
    rays_o, rays_d, t_rand = _f32c(rays_o), _f32c(rays_d), _f32c(t_rand)
    n_rays = rays_o.numel() // 3
    if t_rand is not None and t_rand.shape != (*lead, n_samples):
        raise RuntimeError("sample_stratified: t_rand must be rays.shape[:-1] + (S,)")
    dev = rays_o.device
    z_base, lower, upper = _tables_on(dev, near, far, n_samples, lindisp)
    z = torch.empty((*lead, n_samples), device=dev, dtype=torch.float32)
    pts = torch.empty((*lead, n_samples, 3), device=dev, dtype=torch.float32) if want_pts else None
    if n_rays:
        with torch.cuda.device(dev):
            _lib.call("nfs_sample_stratified", ptr(rays_o), ptr(rays_d), ptr(z_base), ptr(lower), ptr(upper),
                      ptr(t_rand), n_rays, n_samples, ptr(z), ptr(pts), _stream())
    return pts, z


def sample_hierarchical(rays_o, rays_d, z_vals, weights, n_importance, u=None, cdf=None, debug=False,
                        want_pts=True):
    """Inverse-CDF resampling (ray_utils.py:101-143).  z_vals (N,M+1), weights (N,M);
    u (N,Ni) uniform draws or a (Ni,) table broadcast to every ray (perturb=False).
    cdf (N,M+1): use this cdf instead of the kernel's own (kernel-level parity).
    Returns pts (N,M+1+Ni,3), z (N,M+1+Ni) [, dict(cdf, idx, samples) when debug]."""
    _need_cuda("sample_hierarchical", rays_o, rays_d, z_vals, weights, u, cdf)
    if z_vals.dim() != 2 or weights.dim() != 2:
        raise RuntimeError("sample_hierarchical: z_vals and weights must be 2-D")
    N, M1 = z_vals.shape
    M = weights.shape[-1]
    if M1 != M + 1 or weights.shape[0] != N:
        # the reference raises here too (expand of (N,S) to last dim S+1, ray_utils.py:127-129)
        raise RuntimeError(
            "sample_hierarchical: z_vals (N,%d) needs weights (N,%d), got %s - the reference's gather "
            "only works for weights.shape[-1] == z_vals.shape[-1] - 1" % (M1, M1 - 1, tuple(weights.shape)))
    if M <= 0:
        raise RuntimeError("sample_hierarchical: need at least one bin")
    Ni = int(n_importance)
    dev = z_vals.device
    # The resampling is evaluated without autograd (the reference's is differentiable w.r.t. weights / z_vals through the
    # interpolation, ray_utils.py:132-135; NeRF pipelines detach there and so does every caller in the reference):
    # inputs are detached, INTEGRATION.md 'Differentiability'.
    rays_o, rays_d, z_vals, weights, cdf = (t.detach() if t is not None else None for t in (rays_o, rays_d, z_vals, weights, cdf))
    rays_o, rays_d, z_vals, weights, cdf = _f32c(rays_o), _f32c(rays_d), _f32c(z_vals), _f32c(weights), _f32c(cdf)
    if os.environ.get("NFS_DEBUG_CHECKS", "0") != "0" and z_vals.numel() and not bool((z_vals[:, 1:] >= z_vals[:, :-1]).all()):
        # the kernel merges the new samples into z_vals by rank and needs z_vals ascending (every sampler of the
        # reference produces ascending depths); the reference's torch.sort would also accept unsorted input
        raise RuntimeError("sample_hierarchical: z_vals must be ascending along the last dimension")
    if u is None:
        raise RuntimeError("sample_hierarchical: u is required (draws or the linspace table)")
    u = _f32c(u)
    if u.dim() == 1:
        u_stride = 0
        if u.numel() != Ni:
            raise RuntimeError("sample_hierarchical: u table must have N_importance entries")
    else:
        if u.shape != (N, Ni):
            raise RuntimeError("sample_hierarchical: u must be (N, N_importance)")
        u_stride = Ni
    total = M1 + Ni
    z_out = torch.empty((N, total), device=dev, dtype=torch.float32)
    pts = torch.empty((N, total, 3), device=dev, dtype=torch.float32) if want_pts else None
    dbg = None
    if debug:
        dbg = dict(cdf=torch.empty((N, M1), device=dev, dtype=torch.float32),
                   idx=torch.empty((N, Ni), device=dev, dtype=torch.int64),
                   samples=torch.empty((N, Ni), device=dev, dtype=torch.float32))
    if N:
        with torch.cuda.device(dev):
            _lib.call("nfs_sample_hierarchical", ptr(rays_o), ptr(rays_d), ptr(z_vals), ptr(weights), ptr(u),
                      u_stride, ptr(cdf), N, M, Ni, ptr(z_out), ptr(pts),
                      ptr(dbg["cdf"]) if debug else None, ptr(dbg["idx"]) if debug else None,
                      ptr(dbg["samples"]) if debug else None, _stream())
    if debug:
        return pts, z_out, dbg
    return pts, z_out


# --------------------------------------------------------------------------- K5
def project_gather(points_3d, pose, focal, H, W, features=None, want_projection=True, pose_inv=None):
    """World points (...,3) -> (points_2d (...,2), depths (...), valid (...) bool[, sampled (...,C)]):
    project_points_to_image (ray_utils.py:176-210) and, when `features` (1|B=1,Hp,Wp,C) is given, the
    bilinear feature lookup of sample_features_at_points (dino_feature_model.py:114-148) in the same
    kernel.  The 4x4 inverse is torch.inverse on the device, like the reference (ray_utils.py:192); callers
    that reuse a view pass `pose_inv` (torch.inverse synchronises, which a CUDA-graph capture cannot)."""
    _need_cuda("project_gather", points_3d, pose, features)
    lead = points_3d.shape[:-1]
    if points_3d.shape[-1] != 3 or pose.shape != (4, 4):
        raise RuntimeError("project_gather: points must be (...,3) and pose (4,4)")
    pts = _f32c(points_3d).reshape(-1, 3)
    P = pts.shape[0]
    dev = pts.device
    pose_inv = torch.inverse(_f32c(pose)).contiguous() if pose_inv is None else _f32c(pose_inv)
    feat = None
    Hp = Wp = C = 0
    if features is not None:
        if features.dim() == 4:
            if features.shape[0] != 1:
                raise RuntimeError("project_gather: one feature map per call (batch 1), got %s" % (tuple(features.shape),))
            features = features[0]
        if features.dim() != 3:
            raise RuntimeError("project_gather: features must be (1,Hp,Wp,C) or (Hp,Wp,C)")
        feat = _f32c(features)
        Hp, Wp, C = feat.shape
    p2d = torch.empty((P, 2), device=dev, dtype=torch.float32) if want_projection else None
    depth = torch.empty((P,), device=dev, dtype=torch.float32) if want_projection else None
    valid = torch.empty((P,), device=dev, dtype=torch.uint8)
    sampled = torch.empty((P, C), device=dev, dtype=torch.float32) if feat is not None else None
    if P:
        with torch.cuda.device(dev):
            _lib.call("nfs_project_gather", ptr(pts), ptr(pose_inv), float(focal), int(H), int(W), ptr(feat), Hp, Wp, C,
                      P, ptr(p2d), ptr(depth), ptr(valid), ptr(sampled), _stream())
    out = (p2d.reshape(*lead, 2) if want_projection else None, depth.reshape(lead) if want_projection else None,
           valid.bool().reshape(lead))
    if feat is not None:
        out = out + (sampled.reshape(*lead, C),)
    return out


def sample_features(features, points_2d):
    """F.grid_sample(bilinear, zeros, align_corners=False) of a (1,Hp,Wp,C) map at normalised points (N,2)
    (dino_feature_model.py:114-148) -> (N,C): the lookup half of project_gather for callers that already
    hold points_2d."""
    _need_cuda("sample_features", features, points_2d)
    if features.dim() != 4 or features.shape[0] != 1 or points_2d.dim() != 2 or points_2d.shape[-1] != 2:
        raise RuntimeError("sample_features: features (1,Hp,Wp,C), points_2d (N,2)")
    N = points_2d.shape[0]
    pts = _f32c(points_2d)
    feat = _f32c(features[0])
    Hp, Wp, C = feat.shape
    sampled = torch.empty((N, C), device=pts.device, dtype=torch.float32)
    if N:
        with torch.cuda.device(pts.device):
            _lib.call("nfs_project_gather", ptr(pts), None, 1.0, 2, 2, ptr(feat), Hp, Wp, C, N, None, None, None,
                      ptr(sampled), _stream())
    return sampled


# --------------------------------------------------------------------------- K3 building blocks
def _bf16c(t):
    if t is None:
        return None
    if t.dtype != torch.bfloat16:
        raise RuntimeError("expected a bfloat16 tensor, got %s" % t.dtype)
    return t.contiguous()


def linear_bf16(x, w, bias=None, act=0, relu_mask_src=None, out_bf16=True, out_f32_cols=0, out=None):
    """Y = act(X W^T + b) on tcgen05 (nfs_linear_bf16).  x [P,K] bf16, w [N,K] bf16 (K % 64 == 0,
    N % 32 == 0), bias fp32 [N].  `out`: a bf16 [P,N] column block (stride(1) == 1, any row pitch) of a
    wider tensor that receives the result instead of a fresh tensor.
    Returns (y_bf16 [P,N] | None, y_f32 [P,out_f32_cols] | None)."""
    _need_cuda("linear_bf16", x, w, bias, relu_mask_src, out)
    x, w, relu_mask_src = _bf16c(x), _bf16c(w), _bf16c(relu_mask_src)
    P, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise RuntimeError("linear_bf16: x [P,%d] does not match w %s" % (K, tuple(w.shape)))
    if relu_mask_src is not None and relu_mask_src.shape != (P, N):
        raise RuntimeError("linear_bf16: relu_mask_src must be [P,N]")
    bias = _f32c(bias)
    y16, pitch = None, 0
    if out is not None:
        if out.dtype != torch.bfloat16 or out.shape != (P, N) or out.stride(1) != 1:
            raise RuntimeError("linear_bf16: out must be a bf16 [P,N] block with contiguous columns")
        y16, pitch = out, out.stride(0)
    elif out_bf16:
        y16 = torch.empty((P, N), device=x.device, dtype=torch.bfloat16)
    y32 = torch.empty((P, out_f32_cols), device=x.device, dtype=torch.float32) if out_f32_cols else None
    if P:
        with torch.cuda.device(x.device):
            _lib.call("nfs_linear_bf16", ptr(x), ptr(w), ptr(bias), ptr(relu_mask_src), P, K, N, int(act),
                      int(out_f32_cols), ptr(y16), int(pitch), ptr(y32), _stream())
    return y16, y32


def wgrad_bf16(u, v, dw, ld_m, ld_n, colsum=None, colsum_of_v=True, m_valid=0, n_valid=0):
    """dw[m*ld_m + n*ld_n] += sum_p u[p,m] v[p,n]  (nfs_wgrad_bf16).  u [P,M], v [P,N] bf16 (row
    slices of wider tensors are fine: the row pitch is taken from the strides); dw fp32."""
    _need_cuda("wgrad_bf16", u, v, dw, colsum)
    if u.dtype != torch.bfloat16 or v.dtype != torch.bfloat16 or u.stride(1) != 1 or v.stride(1) != 1:
        raise RuntimeError("wgrad_bf16: operands must be bf16 with contiguous columns")
    if dw.dtype != torch.float32 or (colsum is not None and colsum.dtype != torch.float32):
        raise RuntimeError("wgrad_bf16: destinations must be fp32")
    P, M = u.shape
    N = v.shape[1]
    if v.shape[0] != P:
        raise RuntimeError("wgrad_bf16: operands disagree on the number of points")
    if P:
        with torch.cuda.device(u.device):
            _lib.call("nfs_wgrad_bf16", ptr(u), u.stride(0), ptr(v), v.stride(0), P, M, N, int(m_valid), int(n_valid),
                      ptr(dw), int(ld_m), int(ld_n), ptr(colsum), int(bool(colsum_of_v)), _stream())
    return dw


def wgrad_job_array(jobs, who="wgrad_multi"):
    """list of dicts with the keyword arguments of wgrad_bf16 (u, v, dw, ld_m, ld_n, colsum, colsum_of_v, m_valid,
    n_valid) -> (ctypes array of nfs_wgrad_job, number of non-empty jobs, device, indices of the kept jobs)."""
    arr = (_lib.WgradJob * max(len(jobs), 1))()
    n, dev, kept = 0, None, []
    for i, jb in enumerate(jobs):
        u, v, dw, colsum = jb["u"], jb["v"], jb["dw"], jb.get("colsum")
        _need_cuda(who, u, v, dw, colsum)
        if u.dtype != torch.bfloat16 or v.dtype != torch.bfloat16 or u.stride(1) != 1 or v.stride(1) != 1:
            raise RuntimeError(who + ": operands must be bf16 with contiguous columns")
        if dw.dtype != torch.float32 or (colsum is not None and colsum.dtype != torch.float32):
            raise RuntimeError(who + ": destinations must be fp32")
        if v.shape[0] != u.shape[0]:
            raise RuntimeError(who + ": operands disagree on the number of points")
        if u.shape[0] == 0:
            continue
        dev = u.device
        a = arr[n]
        a.u_bf16, a.u_pitch, a.v_bf16, a.v_pitch = u.data_ptr(), u.stride(0), v.data_ptr(), v.stride(0)
        a.n_points, a.m_dim, a.n_dim = u.shape[0], u.shape[1], v.shape[1]
        a.m_valid, a.n_valid = int(jb.get("m_valid", 0)), int(jb.get("n_valid", 0))
        a.dw, a.ld_m, a.ld_n = dw.data_ptr(), int(jb["ld_m"]), int(jb["ld_n"])
        a.colsum = colsum.data_ptr() if colsum is not None else None
        a.colsum_of_v = int(bool(jb.get("colsum_of_v", True)))
        kept.append(i)
        n += 1
    return arr, n, dev, kept


def wgrad_multi(jobs):
    """Several wgrad_bf16 calls in one launch (nfs_wgrad_multi_bf16).  jobs: list of dicts with the keyword
    arguments of wgrad_bf16 (u, v, dw, ld_m, ld_n, colsum, colsum_of_v, m_valid, n_valid)."""
    arr, n, dev, _ = wgrad_job_array(jobs)
    if n:
        with torch.cuda.device(dev):
            _lib.call("nfs_wgrad_multi_bf16", ctypes.byref(arr), n, _stream())


def render_fused(plan, freqs, rays_o, rays_d, near, far, n_coarse, n_importance=0, t_rand=None, u=None, white_bkgd=False,
                 want_weights=False):
    """nfs_render_fused_fwd: stratified depths -> sampler + encoding + MLP (one kernel) -> compositing
    [-> inverse-CDF resampling -> MLP -> compositing] behind one C call (inference; plan = mlp.G1Plan of a
    nerf_model.NeRFMLP whose first layer takes the 10-octave encoding).  t_rand (N,Sc)|None and u (N,Ni)|(Ni,) as
    sample_stratified / sample_hierarchical.  Returns a dict like pipeline.render_rays."""
    rays_o, rays_d, t_rand = _f32c(rays_o), _f32c(rays_d), _f32c(t_rand)
    _need_cuda("render_fused", rays_o, rays_d, t_rand, u)
    N = rays_o.shape[0]
    dev = rays_o.device
    Sc, Ni = int(n_coarse), int(n_importance)
    Sf = Sc + Ni
    z_base, lower, upper = _tables_on(dev, near, far, Sc, False)
    f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    z_c, raw_c, w_c, rgb_c, dep_c = f(N, Sc), f(N, Sc, 4), f(N, Sc), f(N, 3), f(N)
    out = {"rgb": rgb_c, "depth": dep_c, "weights": w_c, "z_vals": z_c}
    z_f = raw_f = w_f = rgb_f = dep_f = bins = None
    u_stride = 0
    if Ni > 0:
        if u is None:
            raise RuntimeError("render_fused: u is required for the fine pass (draws or the linspace table)")
        u = _f32c(u)
        u_stride = 0 if u.dim() == 1 else Ni
        z_f, raw_f, rgb_f, dep_f, bins = f(N, Sf), f(N, Sf, 4), f(N, 3), f(N), f(N, Sc - 1)
        w_f = f(N, Sf) if want_weights else None
        out = {"rgb": rgb_f, "depth": dep_f, "weights": w_f, "z_vals": z_f, "rgb_coarse": rgb_c, "depth_coarse": dep_c,
               "weights_coarse": w_c, "z_coarse": z_c}
    if N:
        model = plan.chain_model(freqs)
        with torch.cuda.device(dev):
            _lib.call("nfs_render_fused_fwd", ctypes.byref(model), ptr(rays_o), ptr(rays_d), N, Sc, ptr(z_base), ptr(lower),
                      ptr(upper), ptr(t_rand), Ni, ptr(u), u_stride, int(bool(white_bkgd)), ptr(z_c), ptr(raw_c), ptr(w_c),
                      ptr(bins), ptr(rgb_c), ptr(dep_c), ptr(z_f), ptr(raw_f), ptr(w_f), ptr(rgb_f), ptr(dep_f), _stream())
    return out
