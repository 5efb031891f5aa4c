"""The render path end to end: stratified sampling -> (fused encoding +) MLP -> compositing
[-> inverse-CDF resampling -> MLP -> compositing], i.e. what NeRFDINOTrainer.render_rays
(/root/reference/src/training/train.py:188-242) does for the baseline model, with the
hierarchical pass of utils.ray_utils.hierarchical_sampling (ray_utils.py:86-143) wired in the
way BASELINE.json config 3 / 5 describe (64 coarse + 128 importance samples = 192 fine).

Every stage is one of the library's CUDA kernels; tensors between stages stay in HBM in the
layout the next kernel wants (the MLP emits packed [r,g,b,sigma] rows, which is exactly the
packed operand of the compositing kernel, so nothing is sliced or copied in between).
"""
import os

import torch

from . import ops


def render_rays(model, freq_bands, rays_o, rays_d, near, far, n_coarse=64, n_importance=0, perturb=True,
                white_bkgd=False, t_rand=None, u=None, target=None):
    """model: models.nerf_model.NeRFMLP (or any module with forward_points(points, freq_bands)
    -> (...,4) [rgb|sigma]).  rays (N,3).  Returns a dict with rgb/depth/weights/z of the last
    pass ('rgb', ...) and of the coarse pass ('rgb_coarse', ...) when n_importance > 0.
    t_rand (N,n_coarse) / u (N,n_importance) override the draws (parity tests); otherwise they come
    from the global CUDA generator in the order the reference draws them.
    target (N,3): the rgb MSE of every pass (train.py:36-44) is evaluated in the compositing kernel's epilogue
    (ops.composite_loss); out['loss'] = sum over the passes is then the only differentiable entry and the
    renderings are detached."""
    N = rays_o.shape[0]
    dev = rays_o.device
    if perturb and t_rand is None and u is None and n_importance > 0 and (N * n_coarse) % 4 == 0:
        # both draws of the step from one generator launch (the coarse jitter first, as the reference draws them)
        r = torch.rand(N * (n_coarse + n_importance), device=dev)
        t_rand, u = r[:N * n_coarse].view(N, n_coarse), r[N * n_coarse:].view(N, n_importance)
    elif perturb and t_rand is None:
        t_rand = torch.rand(N, n_coarse, device=dev)
    plan = model._get_plan() if hasattr(model, "_get_plan") else None
    in_kernel = plan is not None and hasattr(model, "forward_rays") and (plan.refresh() or True) \
        and getattr(plan, "can_encode_in_kernel", lambda f: False)(freq_bands) and os.environ.get("NFS_RENDER_FUSED", "1") != "0"
    if in_kernel and target is None and not torch.is_grad_enabled():
        # inference: the whole path behind one C call (nfs_render_fused_fwd); positions / encodings never reach HBM
        if n_importance > 0 and u is None:
            u = torch.rand(N, n_importance, device=dev) if perturb else ops.importance_table(dev, n_importance)
        return ops.render_fused(plan, freq_bands, rays_o, rays_d, near, far, n_coarse, n_importance,
                                t_rand=t_rand if perturb else None, u=u, white_bkgd=white_bkgd, want_weights=True)
    sess = _active_session(model) if (in_kernel and target is not None and torch.is_grad_enabled()) else None
    if sess is not None and sess.merged and os.environ.get("NFS_RENDER_FUSED_TRAIN", "1") != "0":
        # training step on the merged route: the whole forward behind one C call (nfs_render_fused_fwd_train), the whole
        # backward behind another (nfs_render_fused_bwd, issued by the session's flush)
        if n_importance > 0 and u is None:
            u = torch.rand(N, n_importance, device=dev) if perturb else ops.importance_table(dev, n_importance)
        return _render_train(sess, plan, freq_bands, rays_o, rays_d, target, near, far, n_coarse, n_importance,
                             t_rand if perturb else None, u, white_bkgd)
    if in_kernel:
        # sampler fused into the chain kernel: only the depths are sampled here
        _, z = ops.sample_stratified(rays_o, rays_d, near, far, n_coarse, t_rand=t_rand if perturb else None, want_pts=False)
        raw = model.forward_rays(rays_o, rays_d, z, freq_bands)
    else:
        pts, z = ops.sample_stratified(rays_o, rays_d, near, far, n_coarse, t_rand=t_rand if perturb else None)
        raw = model.forward_points(pts.reshape(-1, 3), freq_bands).reshape(N, n_coarse, 4)
    loss = None
    if target is not None:
        c = ops.composite_loss(raw, None, z, rays_d, target, None, 1.0, 0.0, white_bkgd, want_weights=n_importance > 0,
                               dy_slot=_dy_slot(model))
        rgb, depth, weights, loss = c["rgb_map"], c["depth_map"], c.get("weights"), c["total"]
    else:
        rgb, depth, weights = ops.composite_packed(raw, z, rays_d, white_bkgd=white_bkgd, want_aux=True)
    out = {"rgb": rgb, "depth": depth, "weights": weights, "z_vals": z}
    if n_importance > 0:
        out.update(rgb_coarse=rgb, depth_coarse=depth, weights_coarse=weights, z_coarse=z)
        with torch.no_grad():
            w_in = weights.detach()[:, :-1].contiguous()      # M = S-1 bins between the S coarse depths
            if u is None:
                u = torch.rand(N, n_importance, device=dev) if perturb else ops.importance_table(dev, n_importance)
            pts_f, z_f = ops.sample_hierarchical(rays_o, rays_d, z, w_in, n_importance, u=u, want_pts=not in_kernel)
        S = n_coarse + n_importance
        if in_kernel:
            raw_f = model.forward_rays(rays_o, rays_d, z_f, freq_bands)
        else:
            raw_f = model.forward_points(pts_f.reshape(-1, 3), freq_bands).reshape(N, S, 4)
        if target is not None:
            c = ops.composite_loss(raw_f, None, z_f, rays_d, target, None, 1.0, 0.0, white_bkgd, dy_slot=_dy_slot(model))
            rgb_f, depth_f, w_f, loss = c["rgb_map"], c["depth_map"], None, loss + c["total"]
        else:
            rgb_f, depth_f, w_f = ops.composite_packed(raw_f, z_f, rays_d, white_bkgd=white_bkgd, want_aux=True)
        out.update(rgb=rgb_f, depth=depth_f, weights=w_f, z_vals=z_f)
    if loss is not None:
        out["loss"] = loss
    return out


def _active_session(model):
    """The mlp.StepSession that pipeline.train_step / GraphedTrainStep opened on the model's plan, or None."""
    get = getattr(model, "_get_plan", None)
    return getattr(get(), "_session", None) if get is not None else None


class _RenderTrainFn(torch.autograd.Function):
    """Forward = nfs_render_fused_fwd_train (one C call: sampler + encoding + MLP + compositing + loss for the coarse and
    the fine pass, activations into the session's arenas); backward hands both compositing backwards to the session,
    whose flush() issues nfs_render_fused_bwd (one C call: compositing backward + dgrad chain + all weight gradients).
    The only differentiable output is the loss; parameter gradients go straight into the optimizer's flat buffer."""

    @staticmethod
    def forward(ctx, sess, plan, freqs, rays_o, rays_d, target, cfg, t_rand, u, *params):
        import ctypes
        from . import _lib
        from ._lib import ptr
        from .ops import _f32c, _stream, _tables_on
        near, far, Sc, Ni, white = cfg
        rays_o, rays_d, target, t_rand = _f32c(rays_o), _f32c(rays_d), _f32c(target), _f32c(t_rand)
        N, dev, Sf = rays_o.shape[0], rays_o.device, Sc + Ni
        z_base, lower, upper = _tables_on(dev, near, far, Sc, False)
        f = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
        z_c, raw_c, w_c, rgb_c, dep_c, g_c = f(N, Sc), f(N, Sc, 4), f(N, Sc), f(N, 3), f(N), f(N, 3)
        z_f = raw_f = rgb_f = dep_f = g_f = bins = None
        u_stride = 0
        r0 = [sess.take(N * Sc), 0]
        if Ni > 0:
            u = _f32c(u)
            u_stride = 0 if u.dim() == 1 else Ni
            z_f, raw_f, rgb_f, dep_f, g_f, bins = f(N, Sf), f(N, Sf, 4), f(N, 3), f(N), f(N, 3), f(N, Sc - 1)
            r0[1] = sess.take(N * Sf)
        sums = torch.empty(128, device=dev, dtype=torch.float64)
        loss3 = f(3)
        st = _lib.ChainTrain(sess.x16.data_ptr(), sess.save.data_ptr(), sess.bits.data_ptr(), sess.rows_cap,
                             (ctypes.c_int64 * 2)(*r0))
        model = plan.chain_model(freqs)
        with torch.cuda.device(dev):
            _lib.call("nfs_render_fused_fwd_train", ctypes.byref(model), ctypes.byref(st), ptr(rays_o), ptr(rays_d), N, Sc,
                      ptr(z_base), ptr(lower), ptr(upper), ptr(t_rand), Ni, ptr(u), u_stride, int(bool(white)), ptr(target),
                      1.0, ptr(z_c), ptr(raw_c), ptr(w_c), ptr(bins), ptr(rgb_c), ptr(dep_c), ptr(g_c), ptr(z_f), ptr(raw_f),
                      ptr(rgb_f), ptr(dep_f), ptr(g_f), ptr(sums), ptr(loss3), _stream())
        ctx.sess, ctx.r0, ctx.cfg, ctx.n_params = sess, r0, (N, Sc, Sf, Ni, white), len(params)
        ctx.save_for_backward(*[t for t in (rays_d, raw_c, z_c, g_c, raw_f, z_f, g_f) if t is not None])
        outs = (loss3[0], rgb_c, dep_c, w_c, z_c) + ((rgb_f, dep_f, z_f) if Ni > 0 else ())
        ctx.mark_non_differentiable(*outs[1:])
        ctx.set_materialize_grads(False)
        return outs

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        N, Sc, Sf, Ni, white = ctx.cfg
        sess = ctx.sess
        none = (None,) * (9 + ctx.n_params)
        saved = ctx.saved_tensors
        rays_d, raw_c, z_c, g_c = saved[:4]
        passes = [(ctx.r0[0], raw_c, z_c, g_c, Sc)]
        if Ni > 0:
            raw_f, z_f, g_f = saved[4:7]
            passes.append((ctx.r0[1], raw_f, z_f, g_f, Sf))
        for r0, raw, z, g, S in passes:
            sess.pending -= 1
            if g_loss is None:
                continue
            if not sess.defer_composite_backward(r0, raw, z, rays_d, g * g_loss, None, N, S, int(bool(white))):
                raise RuntimeError("nfs_b200: the fused training render needs the session's merged backward route")
            sess.dy_written.add(r0)
        return none


def _render_train(sess, plan, freqs, rays_o, rays_d, target, near, far, n_coarse, n_importance, t_rand, u, white_bkgd):
    outs = _RenderTrainFn.apply(sess, plan, freqs, rays_o, rays_d, target, (near, far, int(n_coarse), int(n_importance),
                                bool(white_bkgd)), t_rand, u, *plan.params())
    loss, rgb_c, dep_c, w_c, z_c = outs[:5]
    out = {"rgb": rgb_c, "depth": dep_c, "weights": w_c, "z_vals": z_c, "loss": loss}
    if n_importance > 0:
        rgb_f, dep_f, z_f = outs[5:8]
        out.update(rgb_coarse=rgb_c, depth_coarse=dep_c, weights_coarse=w_c, z_coarse=z_c, rgb=rgb_f, depth=dep_f,
                   weights=None, z_vals=z_f)
    return out


def _dy_slot(model):
    """(StepSession, first row) of the model's last forward_points call when it ran inside a step session and autograd
    is recording: the compositing backward then writes the MLP head's bf16 gradient operand in place."""
    import os
    if not torch.is_grad_enabled() or os.environ.get("NFS_K1_BWD_DY", "1") == "0":
        return None
    get = getattr(model, "_get_plan", None)
    return getattr(get(), "_last_slot", None) if get is not None else None


def train_step(model, optimizer, freq_bands, rays_o, rays_d, target, near, far, n_coarse=64, n_importance=128,
               perturb=True, loss_scale=1.0, allreduce=None, fused_loss=True):
    """One optimisation step of BASELINE config 3: render (coarse + fine), MSE on both passes
    (train.py:36-44 uses rgb MSE only; evaluated in the compositing epilogue unless fused_loss=False),
    backward through the compositing and MLP kernels, optional gradient all-reduce, fused Adam.
    Returns the (detached) loss tensor - no host sync."""
    optimizer.zero_grad()
    if getattr(allreduce, "in_graph", False):
        allreduce.wait_readers()             # peers have finished reading the last step's gradient buffer
    sess = _session_for(model, optimizer)
    if sess is not None:
        sess.begin(rays_o.shape[0] * (n_coarse + (n_coarse + n_importance if n_importance > 0 else 0)))
    try:
        if fused_loss:
            loss = render_rays(model, freq_bands, rays_o, rays_d, near, far, n_coarse, n_importance, perturb,
                               target=target)["loss"]
        else:
            out = render_rays(model, freq_bands, rays_o, rays_d, near, far, n_coarse, n_importance, perturb)
            loss = torch.mean((out["rgb"] - target) ** 2)
            if n_importance > 0:
                loss = loss + torch.mean((out["rgb_coarse"] - target) ** 2)
        (loss * loss_scale if loss_scale != 1.0 else loss).backward()
    except BaseException:
        if sess is not None:
            sess.abort()
        raise
    if hasattr(optimizer, "gather_grads"):
        if sess is not None:
            sess.flush()                     # all weight gradients of the step, straight into optimizer.grad
            g = optimizer.grad
        else:
            g = optimizer.gather_grads()
        if getattr(allreduce, "in_graph", False):
            allreduce.fused_step()           # sum over the ranks fused into the Adam kernel (peer memory)
        else:
            if allreduce is not None:
                allreduce(g)
            optimizer.step(grad_scale=1.0, gathered=True)
    else:
        optimizer.step()
    return loss.detach()


def _session_for(model, optimizer):
    """The model's mlp.StepSession (merged weight-gradient launches writing into the optimizer's flat gradient)
    when model is the plain NeRFMLP on the fused chain and optimizer is the FusedAdam that owns all of its
    parameters; else None (autograd accumulates per-call gradients as usual)."""
    from . import mlp
    get = getattr(model, "_get_plan", None)
    if get is None or not hasattr(optimizer, "gather_grads"):
        return None
    plan = get()
    if not isinstance(plan, mlp.G1Plan):
        return None
    plan.refresh()
    if not mlp.StepSession.supported(plan, optimizer):
        return None
    sess = getattr(plan, "_step_session", None)
    if sess is None or sess.opt is not optimizer:
        sess = plan._step_session = mlp.StepSession(plan, optimizer)
    return sess


@torch.no_grad()
def render_image(model, freq_bands, rays_o, rays_d, near, far, n_coarse=64, n_importance=128, chunk=65536,
                 white_bkgd=False):
    """Full-frame render (BASELINE config 5): rays (R,3) in chunks, eval mode (perturb=False)."""
    outs = []
    for i in range(0, rays_o.shape[0], chunk):
        o = render_rays(model, freq_bands, rays_o[i:i + chunk], rays_d[i:i + chunk], near, far, n_coarse,
                        n_importance, perturb=False, white_bkgd=white_bkgd)
        outs.append(o["rgb"])
    return torch.cat(outs, 0) if outs else rays_o.new_zeros((0, 3))


def render_rays_conditioned(model, rays_o, rays_d, near, far, n_samples, pose, focal, H, W, feature_map,
                            perturb=True, white_bkgd=False, t_rand=None, pose_inv=None):
    """NeRFDINOTrainer.render_rays with use_dino (train.py:188-242, BASELINE config 4): stratified samples ->
    projection onto the source view + bilinear feature lookup (one kernel) -> NeRFWithDINO -> compositing.
    model: models.nerf_mlp.NeRFWithDINO; feature_map (1,Hp,Wp,C) precomputed for the view; pose (4,4)."""
    N = rays_o.shape[0]
    if perturb and t_rand is None:
        t_rand = torch.rand(N, n_samples, device=rays_o.device)
    pts, z = ops.sample_stratified(rays_o, rays_d, near, far, n_samples, t_rand=t_rand if perturb else None)
    pts_flat = pts.reshape(-1, 3)
    dirs = rays_d.unsqueeze(1).expand(-1, n_samples, -1).reshape(-1, 3)          # train.py:225
    plan = model._get_plan() if hasattr(model, "_get_plan") else None
    if hasattr(plan, "chain_a_bwd") and plan.D > 0 and os.environ.get("NFS_G3_OPERAND", "1") != "0":
        # projection + feature lookup + encoding as the producer of the first layer's operand (nfs_g3_operand)
        from . import mlp_g3
        if pose_inv is None:
            pose_inv = torch.inverse(pose)                                         # ray_utils.py:192
        rgb, density = mlp_g3.g3_forward_from_map(plan, pts_flat, dirs, feature_map, pose_inv, focal, H, W)
    else:
        _, _, _, feats = ops.project_gather(pts_flat, pose, focal, H, W, features=feature_map, want_projection=False,
                                            pose_inv=pose_inv)
        rgb, density = model(pts_flat, dirs, feats)
    rgb_map, depth, weights = ops.composite(rgb.reshape(N, n_samples, 3), density.reshape(N, n_samples, 1), z, rays_d,
                                            white_bkgd=white_bkgd)
    return {"rgb": rgb_map, "depth": depth, "weights": weights, "z_vals": z}


class GraphedStep:
    """An optimisation step captured once into CUDA graphs and replayed: every launch of
    `loss_closure()` (render + loss), its backward, the gradient flattening and the fused Adam become one
    cudaGraphLaunch, so the step is paced by the GPU instead of by Python / launch latency.

    Everything the step reads must be static: the closure reads fixed device buffers (the caller copies new
    inputs into them before calling), the Adam step count and learning rate live on the device
    (nfs_adam_step_dev), uniform draws come from torch's graph-safe generator.  With `allreduce` (data
    parallel) the step is two graphs with the gradient all-reduce between them:
    [render + backward + gradient flattening] -> NCCL all-reduce of the flat buffer -> [Adam]."""

    def __init__(self, optimizer, loss_closure, loss_scale=1.0, allreduce=None, warmup=3):
        from . import mlp
        if not hasattr(optimizer, "gather_grads"):
            raise RuntimeError("GraphedStep needs nfs_b200.optim.FusedAdam (device-resident optimizer state)")
        self.opt, self.closure = optimizer, loss_closure
        self.loss_scale, self.allreduce = float(loss_scale), allreduce
        o = optimizer
        dev = o.flat.device
        snap = [t.clone() for t in (o.flat, o.exp_avg, o.exp_avg_sq, o._step_dev, o._state)]
        rng = torch.cuda.get_rng_state(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        self.fused_exchange = bool(getattr(allreduce, "in_graph", False))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):            # first launches: function attributes, caches, allocator
                self._forward_backward()
                if allreduce is not None and not self.fused_exchange:
                    allreduce(o.grad)
                self._update()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # The all-reduce stays OUTSIDE the graphs: capturing the NCCL collective together with the step hung
        # on a 2-GPU box (round 1), so data parallel runs [graph] -> eager all-reduce -> [graph].
        # With the peer-memory exchange (dist.PeerExchange) the reduction is part of the Adam kernel and the whole
        # step is one graph; an NCCL all-reduce stays between two graphs.
        self.allreduce_in_graph = self.fused_exchange
        self.g_step = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_step):
            self.loss = self._forward_backward()
            if allreduce is None or self.fused_exchange:
                self._update()
        self.g_update = None
        if allreduce is not None and not self.fused_exchange:
            self.g_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_update, pool=self.g_step.pool()):
                self._update()
        for dst, src in zip((o.flat, o.exp_avg, o.exp_avg_sq, o._step_dev, o._state), snap):
            dst.copy_(src)
        torch.cuda.set_rng_state(rng, dev)
        mlp.bump_weight_epoch()

    def _forward_backward(self):
        self.opt.zero_grad()
        if self.fused_exchange:
            self.allreduce.wait_readers()
        loss = self.closure()
        (loss * self.loss_scale if self.loss_scale != 1.0 else loss).backward()
        self.opt.gather_grads()
        return loss.detach()

    def _update(self):
        if self.fused_exchange:
            self.allreduce.fused_step()
        else:
            self.opt.step(gathered=True)

    def replay(self):
        """One optimisation step on whatever the static input buffers hold; returns the loss tensor of this
        step (a static buffer, no host sync)."""
        from . import mlp
        self.opt.sync_lr()           # a scheduler may have moved optimizer.lr since the capture: the graph reads the device copy
        self.g_step.replay()
        if self.g_update is not None:               # two-graph form: eager all-reduce between the graphs
            self.allreduce(self.opt.grad)
            self.g_update.replay()
        mlp.bump_weight_epoch()      # the graph changed the fp32 masters: eager callers must repack
        return self.loss


class GraphedTrainStep(GraphedStep):
    """train_step (BASELINE config 3: coarse + fine render of the G1 model, MSE on both passes) as a
    GraphedStep; rays and targets are copied into fixed buffers before each replay."""

    def __init__(self, model, optimizer, freq_bands, n_rays, near, far, n_coarse=64, n_importance=128,
                 perturb=True, loss_scale=1.0, allreduce=None, warmup=3):
        from . import mlp
        dev = optimizer.flat.device
        self.model, self.bands = model, mlp.freqs_on(dev, freq_bands)
        self.cfg = (near, far, n_coarse, n_importance, perturb)
        self.rays_o = torch.zeros(n_rays, 3, device=dev)
        self.rays_d = torch.zeros(n_rays, 3, device=dev)
        self.rays_d[:, 2] = -1.0
        self.target = torch.zeros(n_rays, 3, device=dev)
        super().__init__(optimizer, self._loss, loss_scale=loss_scale, allreduce=allreduce, warmup=warmup)
        # the graphs hold raw pointers into the session's arenas and the plan's stacked operands: keep those
        # tensors alive even if another caller later makes the session / plan allocate bigger ones
        sess = _session_for(model, optimizer)
        plan = model._get_plan() if hasattr(model, "_get_plan") else None
        self._captured = [getattr(o, k, None) for o, keys in (
            (sess, ("x16", "save", "bits", "dy", "dys", "head_w", "head_b", "w0_t", "quad_flags", "scatter")),
            (plan, ("w_stack", "wt_stack", "b_stack", "_table"))) if o is not None for k in keys]

    def _forward_backward(self):
        sess = _session_for(self.model, self.opt)
        if sess is None:
            return super()._forward_backward()
        near, far, n_coarse, n_importance, perturb = self.cfg
        self.opt.zero_grad()
        if self.fused_exchange:
            self.allreduce.wait_readers()
        sess.begin(self.rays_o.shape[0] * (n_coarse + (n_coarse + n_importance if n_importance > 0 else 0)))
        try:
            loss = self.closure()
            (loss * self.loss_scale if self.loss_scale != 1.0 else loss).backward()
        except BaseException:
            sess.abort()
            raise
        sess.flush()
        return loss.detach()

    def _loss(self):
        near, far, n_coarse, n_importance, perturb = self.cfg
        return render_rays(self.model, self.bands, self.rays_o, self.rays_d, near, far, n_coarse, n_importance, perturb,
                           target=self.target)["loss"]

    def __call__(self, rays_o, rays_d, target):
        self.rays_o.copy_(rays_o, non_blocking=True)
        self.rays_d.copy_(rays_d, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        return self.replay()
