#!/usr/bin/env python
"""Benchmark of the NeRF render hot path on B200 (contract: see DESIGN.md "Measurement").

Headline workload at every N: BASELINE.json configs[2], the baseline NeRF training step - 4096 rays per GPU from a
synthetic 800x800 Blender-lego-shaped view, 64 stratified coarse samples + 128 importance samples (192 fine
evaluations) per ray through nerf_model.NeRFMLP, compositing and rgb MSE on both passes, backward, fused Adam
(reference: src/training/train.py:188-292).  One "step" = one optimisation step.  Rays shard across GPUs
(weak scaling: 4096 rays per GPU); the only exchange is the sum of the flat fp32 weight gradient (1.9 MB).

  python bench.py [--gpus N] [--steps K] [--warmup W]          the CUDA path
  python bench.py --impl reference ...                          the reference's CPU path (oracle port of it:
                                                                the same ATen op sequence) on the host cores
`extras` carries BASELINE configs[1] (compositing micro-bench, HBM roofline), configs[3] (DINO-NeRF step) and
configs[4] (800x800 render), each with its own CPU baseline and clock record.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "nerf-few-shot-limitations_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

# ---- headline workload (BASELINE.json configs[2])
TRAIN_RAYS = 4096
N_COARSE, N_IMPORTANCE = 64, 128
EVALS_PER_RAY = N_COARSE + (N_COARSE + N_IMPORTANCE)        # 64 coarse + 192 fine network evaluations
TRAIN_FLOP_PER_POINT = 2823168.0                            # fwd + dgrad (no layer-0 dgrad) + wgrad of G1
FWD_FLOP_PER_POINT = 951808.0
WORKLOAD = "nerf_train_step_4096rays_64c+192f_G1_bf16"
# ---- cfg 2 (compositing micro-bench)
N_RAYS = 1 << 20
N_SAMPLES = 64
FWD_BYTES = 24 * N_SAMPLES + 28          # read rgb 12S, density 4S, z 4S, rays_d 12; write rgb 12, depth 4, weights 4S
BWD_BYTES = 36 * N_SAMPLES + 28          # read 20S + rays_d 12 + g_rgb 12 + g_depth 4; write d_rgb 12S + d_density 4S


def _peaks_file():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    p = _peaks_file()
    if "hbm_gbs" in p:
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bf16_peak():
    """Sustained dense bf16 peak (the step is timed inside a long loop under the power cap)."""
    p = _peaks_file()
    if "bf16_tflops_sustained" in p:
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


def lego_rays(n_rays, H=800, W=800, seed=0):
    """Synthetic Blender-lego-shaped rays: one 800x800 pinhole view,
    camera_angle_x = 0.6911112, camera on the r = 4.0311 sphere at phi = -30 deg looking at the
    origin, directions NOT normalised (models/ray_sampler.py:18-30 convention); n_rays pixels
    drawn with replacement.  Input synthesis only (host side, outside every timed region)."""
    g = torch.Generator().manual_seed(seed)
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    th = math.radians(float(torch.rand((), generator=g) * 360.0 - 180.0))
    ph = math.radians(-30.0)
    t = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0311], [0, 0, 0, 1]], dtype=torch.float32)
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0],
                       [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0],
                       [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    c2w = flip @ rt @ rp @ t
    pick = torch.randint(0, H * W, (n_rays,), generator=g)
    i = (pick % W).float()
    j = (pick // W).float()
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[:, None, :] * c2w[:3, :3], -1)
    return c2w[:3, 3].expand(rays_d.shape).contiguous(), rays_d.contiguous()


def make_inputs(n_rays, n_samples, seed, device):
    """Synthetic lego-shaped compositing batch (SURVEY.md section 8d cfg 2): rgb~U[0,1), density~10*N(0,1),
    z = one stratified draw in [2,6], unnormalised rays_d from an 800x800 Blender camera."""
    g = torch.Generator().manual_seed(seed)
    _, rays_d = lego_rays(n_rays, seed=seed)
    rgb = torch.rand(n_rays, n_samples, 3, generator=g)
    density = torch.randn(n_rays, n_samples, 1, generator=g) * 10.0
    t = torch.linspace(0.0, 1.0, n_samples)
    zb = 2.0 * (1 - t) + 6.0 * t
    mids = 0.5 * (zb[1:] + zb[:-1])
    lower, upper = torch.cat([zb[:1], mids]), torch.cat([mids, zb[-1:]])
    z = lower + (upper - lower) * torch.rand(n_rays, n_samples, generator=g)
    target = torch.rand(n_rays, 3, generator=g)
    depth_t = 2.0 + 4.0 * torch.rand(n_rays, generator=g)
    outs = dict(rgb=rgb, density=density, z=z, rays_d=rays_d, target=target, depth_t=depth_t)
    if device is not None:
        outs = {k: v.to(device) for k, v in outs.items()}
    return outs


class ClockSampler:
    """SM clock / throttle reasons sampled every 50 ms by an NVML thread while the timed region
    runs (nvidia-smi -lms to a file loses its buffered output when it is terminated)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.samples, self.reasons, self.mx, self.err = [], set(), None, None
        self._stop = threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:      # no NVML: the clocks record says so instead of inventing numbers
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append((sm, util))
                for nm, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0}
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
        if self.err:
            out["error"] = self.err
        sm = sorted(s for s, _ in self.samples)
        if sm:
            out.update(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], samples=len(sm))
        out["reasons"] = sorted(self.reasons)
        return out


class Timer:
    """CUDA-event timing of `fn` under a clock record: keeps the GPU under the same load for `settle` seconds before
    and `tail` seconds after the timed region so that the 50 ms NVML samples describe it; max over ranks."""

    def __init__(self, dev, rank, dist, local_rank, quick):
        self.dev, self.rank, self.dist, self.local_rank, self.quick = dev, rank, dist, local_rank, quick

    def sync_all(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def _iters_for(self, fn, seconds):
        """Number of calls of fn that fill `seconds` - the SAME number on every rank (fn may contain a cross-rank
        exchange, so the ranks must make identical call sequences: never loop on the local clock)."""
        if seconds <= 0 or self.quick:
            return 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        per_call = self.max_over_ranks((time.perf_counter() - t0) / 3 * 1e3) * 1e-3
        return max(1, min(2000, int(seconds / max(per_call, 1e-6))))

    def run(self, fn, steps, warmup, settle=0.3, tail=0.2, clocks=True):
        """-> (ms per step [max over ranks], clocks dict | None)"""
        sampler = ClockSampler(self.local_rank) if (clocks and self.rank == 0 and not self.quick) else None
        for _ in range(self._iters_for(fn, settle)):
            fn()
        for _ in range(max(warmup, 3)):
            fn()
        self.sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        self.sync_all()
        ms = self.max_over_ranks(a.elapsed_time(b) / steps)
        if clocks and not self.quick and tail > 0:         # same condition and count on every rank
            for _ in range(max(1, min(2000, int(tail / max(ms * 1e-3, 1e-6))))):
                fn()
            torch.cuda.synchronize()
        return ms, (sampler.stop() if sampler is not None else None)


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_train_step_rate(n_rays, reps, threads, warm=1):
    """The reference's training step on the host cores (oracle port = the same ATen op sequence as
    NeRFDINOTrainer.render_rays + loss + backward + Adam, train.py:188-292, with the hierarchical pass of
    ray_utils.py:86-143): stratified sampling -> encoding -> NeRFMLP -> compositing -> inverse-CDF resampling ->
    encoding -> NeRFMLP -> compositing -> MSE on both passes -> autograd backward -> torch.optim.Adam.
    A step works on `n_rays` rays x (64 + 192) evaluations; returns (rays/s, seconds per step)."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = O.PlainNeRF()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    bands = O.frequency_bands(10)
    ro, rd = lego_rays(n_rays, seed=0)
    target = torch.rand(n_rays, 3)

    def step():
        opt.zero_grad()
        pts, z = O.stratified(ro, rd, 2.0, 6.0, N_COARSE, t_rand=torch.rand(n_rays, N_COARSE))
        raw = model(O.encode(pts.reshape(-1, 3), bands)).reshape(n_rays, N_COARSE, 4)
        rgb_c, _, w = O.render(raw[..., :3], raw[..., 3:], z, rd)
        with torch.no_grad():
            h = O.hierarchical(ro, rd, z, w.detach()[:, :-1], torch.rand(n_rays, N_IMPORTANCE))
        S = N_COARSE + N_IMPORTANCE
        raw_f = model(O.encode(h["pts"].reshape(-1, 3), bands)).reshape(n_rays, S, 4)
        rgb_f, _, _ = O.render(raw_f[..., :3], raw_f[..., 3:], h["z"], rd)
        loss = torch.mean((rgb_c - target) ** 2) + torch.mean((rgb_f - target) ** 2)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    sec = (time.perf_counter() - t0) / reps
    return n_rays / sec, sec


def cpu_composite_rate(n_rays, reps, threads):
    """cfg 2 on the host cores: VolumeRenderer.forward + autograd backward (oracle port), rays/s."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    d = make_inputs(n_rays, N_SAMPLES, seed=0, device=None)
    rgb, den = d["rgb"].requires_grad_(), d["density"].requires_grad_()
    g_rgb = torch.randn(n_rays, 3) / n_rays
    g_depth = torch.randn(n_rays) / n_rays
    total = 0.0
    for i in range(reps + 1):
        t0 = time.perf_counter()
        o = O.render(rgb, den, d["z"], d["rays_d"])
        torch.autograd.grad([o[0], o[1]], [rgb, den], [g_rgb, g_depth])
        if i:                      # first pass warms the allocator / thread pool
            total += time.perf_counter() - t0
    return n_rays / (total / reps), total / reps


def cpu_render_rate(n_rays, reps, threads):
    """cfg 5 on the host cores: eval-mode coarse + fine render of n_rays rays (no_grad), rays/s."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = O.PlainNeRF().eval()
    bands = O.frequency_bands(10)
    ro, rd = lego_rays(n_rays, seed=7)
    u = torch.linspace(0.0, 1.0, N_IMPORTANCE)
    total = 0.0
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            pts, z = O.stratified(ro, rd, 2.0, 6.0, N_COARSE)
            raw = model(O.encode(pts.reshape(-1, 3), bands)).reshape(n_rays, N_COARSE, 4)
            _, _, w = O.render(raw[..., :3], raw[..., 3:], z.contiguous(), rd)
            h = O.hierarchical(ro, rd, z.contiguous(), w[:, :-1], u)
            raw_f = model(O.encode(h["pts"].reshape(-1, 3), bands)).reshape(n_rays, N_COARSE + N_IMPORTANCE, 4)
            O.render(raw_f[..., :3], raw_f[..., 3:], h["z"], rd)
            if i:
                total += time.perf_counter() - t0
    return n_rays / (total / reps), total / reps


def cpu_dino_rate(n_rays, reps, threads):
    """cfg 4 on the host cores: stratified samples -> projection + grid_sample feature lookup -> NeRFWithDINO ->
    compositing -> MSE -> backward -> Adam (oracle port), rays/s."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    model = O.ConditionedNeRF(pos_freq=12, dir_freq=4, dino_dim=64)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    ro, rd = lego_rays(n_rays, H=128, W=128, seed=200)
    tgt = torch.rand(n_rays, 3)
    fmap = torch.randn(1, 9, 9, 64)
    pose = torch.eye(4)
    pose[2, 3] = 4.0
    focal = 0.5 * 128 / math.tan(0.5 * 0.6911112)
    total = 0.0
    for i in range(reps + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        pts, z = O.stratified(ro, rd, 2.0, 6.0, 64, t_rand=torch.rand(n_rays, 64))
        flat = pts.reshape(-1, 3)
        with torch.no_grad():
            p2d, _, _ = O.project_points(flat, pose, focal, 128, 128)
            feats = O.sample_features(fmap, p2d)
        dirs = rd.unsqueeze(1).expand(-1, 64, -1).reshape(-1, 3)
        rgb, den = model(flat, dirs, feats)
        out, _, _ = O.render(rgb.reshape(n_rays, 64, 3), den.reshape(n_rays, 64, 1), z, rd)
        torch.mean((out - tgt) ** 2).backward()
        opt.step()
        if i:
            total += time.perf_counter() - t0
    return n_rays / (total / reps), total / reps


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the headline step (oracle port - the reference is a
    script tree that can neither be pip-installed nor travels to the GPU box).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 256                                        # 1/16 of the step's rays: 65 536 network evaluations
    rate, sec = cpu_train_step_rate(sample, max(1, args.steps), threads, warm=max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": "rays/sec (train fwd+bwd)", "value": rate, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": TRAIN_RAYS, "coarse": N_COARSE, "fine": N_COARSE + N_IMPORTANCE,
                   "note": "each step = a %d-ray sample of the 4096-ray step (x256 evaluations): sampler, encoding, "
                           "NeRFMLP, compositing, hierarchical resampling, MSE, autograd backward, Adam; ATen CPU fp32"
                           % sample},
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": "%d of 4096 rays x (64 + 192) evaluations per step, %.2f s per step" % (sample, sec)},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arms
def kernel_breakdown(step_eager, reps=3):
    """Per-kernel device time of the eagerly launched step (kineto / CUPTI), us per step, largest first."""
    from torch.profiler import profile, ProfilerActivity
    for _ in range(2):
        step_eager()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            step_eager()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0:
            rows.append({"kernel": e.key[:90], "us_per_step": t / reps, "launches_per_step": e.count / reps})
    rows.sort(key=lambda r: -r["us_per_step"])
    return rows


def bench_train_step(T, world, args):
    """Headline: the graph-replayed training step.  -> dict of measurements."""
    from models.nerf_model import NeRFMLP
    from nfs_b200 import _lib, pipeline
    from nfs_b200 import dist as nd
    from nfs_b200.optim import FusedAdam
    dev, rank = T.dev, T.rank
    torch.manual_seed(0)
    model = NeRFMLP().to(dev).train()
    opt = FusedAdam(model.parameters(), lr=5e-4)
    bands = 2.0 ** torch.linspace(0.0, 9.0, 10)
    n = TRAIN_RAYS
    ro, rd = lego_rays(n, seed=100 + rank)
    target = torch.rand(n, 3, generator=torch.Generator().manual_seed(100 + rank))
    allreduce = nd.make_allreduce(opt) if world > 1 else None
    l0 = _lib.launch_count()
    step = pipeline.GraphedTrainStep(model, opt, bands, n, 2.0, 6.0, N_COARSE, N_IMPORTANCE, perturb=True,
                                     loss_scale=1.0 / world, allreduce=allreduce, warmup=3)
    launches_per_step = (_lib.launch_count() - l0) // 4             # 3 eager warm-up steps + 1 capture
    step(ro.to(dev), rd.to(dev), target.to(dev))                    # inputs resident in the step's static buffers
    ms, clk = T.run(step.replay, args.steps, args.warmup, settle=0.5)
    loss_end = float(step.loss.item())

    # end to end through the public call with HOST buffers (pinned): H2D of the step's rays / targets and the D2H
    # read of its loss are inside the timed region
    h_ro, h_rd, h_tg = ro.pin_memory(), rd.pin_memory(), target.pin_memory()
    h_loss = torch.empty(1, pin_memory=True)

    def e2e_step():
        loss = step(h_ro, h_rd, h_tg)
        h_loss.copy_(loss.reshape(1), non_blocking=True)

    e2e_ms, _ = T.run(e2e_step, max(3, args.steps), 3, settle=0.0, tail=0.0, clocks=False)
    out = {"ms": ms, "clocks": clk, "e2e_ms": e2e_ms, "h2d": 3 * n * 3 * 4, "d2h": 4, "launches_per_step": int(launches_per_step),
           "loss": loss_end, "n_params": int(opt.flat.numel()),
           "allreduce": getattr(allreduce, "describe", "none (1 GPU)") if world > 1 else "none (1 GPU)",
           "graphs_per_step": 1 if step.g_update is None else 2}
    if rank == 0 and not args.quick and not args.no_breakdown:
        try:
            d_ro, d_rd, d_tg = ro.to(dev), rd.to(dev), target.to(dev)
            rows = kernel_breakdown(lambda: pipeline.train_step(model, opt, bands, d_ro, d_rd, d_tg, 2.0, 6.0, N_COARSE,
                                                                N_IMPORTANCE))
            out["kernels"] = rows[:10]
            out["kernels_total_us"] = sum(r["us_per_step"] for r in rows)
        except Exception as e:
            out["kernels_error"] = repr(e)[:200]
    out["_keep"] = (model, opt, bands)
    return out


def bench_composite(T, world, args):
    """cfg 2: compositing forward + backward of 2^20 rays x 64 samples per GPU through the C ABI (HBM roofline)."""
    import ctypes
    from nfs_b200 import _lib
    from nfs_b200._lib import ptr
    from models.nerf_mlp import VolumeRenderer
    dev, rank = T.dev, T.rank
    d = make_inputs(N_RAYS, N_SAMPLES, seed=rank, device=dev)
    rgb, den, z, rays_d = d["rgb"], d["density"].reshape(N_RAYS, N_SAMPLES), d["z"], d["rays_d"]
    out_rgb = torch.empty(N_RAYS, 3, device=dev)
    out_depth = torch.empty(N_RAYS, device=dev)
    out_w = torch.empty(N_RAYS, N_SAMPLES, device=dev)
    d_rgb, d_den = torch.empty_like(rgb), torch.empty_like(den)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd():
        _lib.call("nfs_composite_fwd", ptr(rgb), ptr(den), ptr(z), ptr(rays_d), None, 0.0, N_RAYS, N_SAMPLES,
                  0, 0, ptr(out_rgb), ptr(out_depth), ptr(out_w), stream)

    fwd()       # upstream gradients of loss = mse(rgb, target) + 0.1 * mean|depth - d*| at the first forward
    g_rgb = (2.0 / (3 * N_RAYS)) * (out_rgb - d["target"])
    g_depth = (0.1 / N_RAYS) * torch.sign(out_depth - d["depth_t"])

    def bwd():
        _lib.call("nfs_composite_bwd", ptr(rgb), ptr(den), ptr(z), ptr(rays_d), None, 0.0, ptr(g_rgb),
                  ptr(g_depth), None, N_RAYS, N_SAMPLES, 0, 0, ptr(d_rgb), ptr(d_den), stream)

    steps = 5 if args.quick else 20
    fwd_ms, _ = T.run(fwd, steps, 3, settle=0.0, tail=0.0, clocks=False)
    bwd_ms, _ = T.run(bwd, steps, 3, settle=0.0, tail=0.0, clocks=False)
    both_ms, clk = T.run(lambda: (fwd(), bwd()), steps, 3)
    peak, peak_src = peaks()
    res = {"workload": "composite_fwd_bwd_1Mrays_x64_fp32", "rays_per_s": world * N_RAYS / (both_ms * 1e-3),
           "ms_per_step": both_ms, "clocks": clk,
           "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                        "fwd": {"ms": fwd_ms, "gbs": FWD_BYTES * N_RAYS / (fwd_ms * 1e-3) / 1e9,
                                "frac": FWD_BYTES * N_RAYS / (fwd_ms * 1e-3) / 1e9 / peak},
                        "bwd": {"ms": bwd_ms, "gbs": BWD_BYTES * N_RAYS / (bwd_ms * 1e-3) / 1e9,
                                "frac": BWD_BYTES * N_RAYS / (bwd_ms * 1e-3) / 1e9 / peak},
                        "step_gbs": (FWD_BYTES + BWD_BYTES) * N_RAYS / (both_ms * 1e-3) / 1e9,
                        "step_frac": (FWD_BYTES + BWD_BYTES) * N_RAYS / (both_ms * 1e-3) / 1e9 / peak}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            res["roofline"]["traffic"] = json.load(open(traffic_file))
        except Exception:
            pass
    if not args.quick:
        # end to end through VolumeRenderer + autograd with pinned host buffers.  The D2H leg returns the
        # renderings and the loss (16.8 MB); the 1.07 GB of input gradients stay on the device, as they would
        # for the MLP that consumes them.
        vr = VolumeRenderer().eval()
        host = {k: d[k].cpu().pin_memory() for k in ("rgb", "density", "z", "rays_d", "target", "depth_t")}
        h_out = torch.empty(N_RAYS, 4, pin_memory=True)
        h_loss = torch.empty(1, pin_memory=True)

        def e2e_step():
            r = host["rgb"].to(dev, non_blocking=True).requires_grad_()
            s = host["density"].to(dev, non_blocking=True).requires_grad_()
            zz = host["z"].to(dev, non_blocking=True)
            dd = host["rays_d"].to(dev, non_blocking=True)
            tg = host["target"].to(dev, non_blocking=True)
            dt = host["depth_t"].to(dev, non_blocking=True)
            o_rgb, o_depth, _ = vr(r, s, zz, dd)
            loss = ((o_rgb - tg) ** 2).mean() + 0.1 * (o_depth - dt).abs().mean()
            loss.backward()
            h_out[:, :3].copy_(o_rgb.detach(), non_blocking=True)
            h_out[:, 3].copy_(o_depth.detach(), non_blocking=True)
            h_loss.copy_(loss.detach().reshape(1), non_blocking=True)

        e_ms, _ = T.run(e2e_step, 5, 3, settle=0.0, tail=0.0, clocks=False)
        res["e2e"] = {"value": world * N_RAYS / (e_ms * 1e-3), "unit": "rays/s",
                      "h2d_bytes_per_step": sum(host[k].numel() * 4 for k in host), "d2h_bytes_per_step": h_out.numel() * 4 + 4,
                      "note": "gradients (1.07 GB) stay on the device"}
    return res


def bench_dino(T, world, args):
    """cfg 4: DINO-NeRF (experiments/dino_nerf.yaml): NeRFWithDINO, pos_freq 12, 64-d feature map (random values
    stand in for the frozen Dinov2 + projection head, which is per-view preprocessing), batch 512 rays x 64
    samples: projection + bilinear feature lookup + MLP + compositing, forward + backward + Adam, graph replay."""
    from models.nerf_mlp import NeRFWithDINO
    from nfs_b200 import dist as nd
    from nfs_b200 import pipeline
    from nfs_b200.optim import FusedAdam
    dev, rank = T.dev, T.rank
    torch.manual_seed(1)
    g3 = NeRFWithDINO(pos_freq=12, dir_freq=4, dino_dim=64).to(dev).train()
    opt3 = FusedAdam(g3.parameters(), lr=5e-4)
    nb = 512
    ro4, rd4 = lego_rays(nb, H=128, W=128, seed=200 + rank)
    ro4, rd4 = ro4.to(dev), rd4.to(dev)
    tgt4 = torch.rand(nb, 3, device=dev)
    fmap = torch.randn(1, 9, 9, 64, device=dev)
    pose4 = torch.eye(4, device=dev)
    pose4[2, 3] = 4.0
    focal4 = 0.5 * 128 / math.tan(0.5 * 0.6911112)
    pose4_inv = torch.inverse(pose4)             # per view, outside the captured step

    def loss4():
        o = pipeline.render_rays_conditioned(g3, ro4, rd4, 2.0, 6.0, 64, pose4, focal4, 128, 128, fmap, perturb=True,
                                             pose_inv=pose4_inv)
        return torch.mean((o["rgb"] - tgt4) ** 2)

    allreduce = nd.make_allreduce(opt3) if world > 1 else None
    step4 = pipeline.GraphedStep(opt3, loss4, loss_scale=1.0 / world, allreduce=allreduce)
    ms4, clk = T.run(step4.replay, 5 if args.quick else 30, 3)
    flop4 = 64 * 5.51e6                          # ~5.51 MFLOP per point for G3 training
    peak, _ = bf16_peak()
    return {"workload": "dino_nerf_step_512rays_x64_G3", "rays_per_s": world * nb / (ms4 * 1e-3), "ms_per_step": ms4,
            "rays_per_gpu": nb, "points_per_step": nb * 64, "tflops_per_gpu": nb * flop4 / (ms4 * 1e-3) / 1e12,
            "frac_of_bf16_sustained_peak": nb * flop4 / (ms4 * 1e-3) / 1e12 / peak, "launch": "CUDA graph replay",
            "clocks": clk}


def bench_render(T, world, args, model, bands):
    """cfg 5: one 800x800 frame (640 000 rays, 64 + 192 evaluations per ray, eval mode) split contiguously over the GPUs
    (strong scaling)."""
    from nfs_b200 import dist as nd
    from nfs_b200 import pipeline
    dev, rank = T.dev, T.rank
    total = 640000
    lo, hi = nd.shard_range(total, rank, world)
    ro2, rd2 = lego_rays(hi - lo, seed=7)
    ro2, rd2 = ro2.to(dev), rd2.to(dev)
    model.eval()
    ms, clk = T.run(lambda: pipeline.render_image(model, bands, ro2, rd2, 2.0, 6.0, N_COARSE, N_IMPORTANCE, chunk=65536),
                    1 if args.quick else 3, 1, settle=0.0 if args.quick else 0.3)
    model.train()
    peak, _ = bf16_peak()
    tf = (hi - lo) * EVALS_PER_RAY * FWD_FLOP_PER_POINT / (ms * 1e-3) / 1e12
    return {"workload": "render_800x800_64c+192f_G1", "rays_per_s": total / (ms * 1e-3), "ms_per_frame": ms,
            "rays_per_gpu": hi - lo, "scaling": "strong", "tflops_per_gpu": tf, "frac_of_bf16_sustained_peak": tf / peak,
            "clocks": clk}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg 2 / cfg 4 / cfg 5 measurements")
    ap.add_argument("--no-breakdown", action="store_true", help="skip the per-kernel (kineto) breakdown of the step")
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid (ncu): no clock-settle loops, no e2e leg, no CPU baselines, no breakdown")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    T = Timer(dev, rank, dist, local_rank, args.quick)

    tr = bench_train_step(T, world, args)
    model, opt, bands = tr.pop("_keep")
    extras = None
    if not args.no_extras:
        extras = {}
        for name, fn in (("composite_cfg2", lambda: bench_composite(T, world, args)),
                         ("dino_nerf_cfg4", lambda: bench_dino(T, world, args)),
                         ("render_cfg5", lambda: bench_render(T, world, args, model, bands))):
            try:
                extras[name] = fn()
            except Exception as e:                  # the headline line must survive a failure of an extra
                extras[name] = {"error": repr(e)[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms = tr["ms"]
    peak, peak_src = bf16_peak()
    flops_step = TRAIN_RAYS * EVALS_PER_RAY * TRAIN_FLOP_PER_POINT          # per GPU
    achieved = flops_step / (ms * 1e-3) / 1e12
    burst = _peaks_file().get("bf16_tflops")
    line = {
        "metric": "rays/sec (train fwd+bwd)", "value": world * TRAIN_RAYS / (ms * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu": TRAIN_RAYS, "coarse": N_COARSE, "fine": N_COARSE + N_IMPORTANCE,
                   "model": "nerf_model.NeRFMLP 63->256x8->4 (random init), fp32 masters, bf16 operands, fp32 accumulate",
                   "step": "sampler + encoding + MLP + compositing, coarse and fine, MSE on both, backward, fused Adam; "
                           "one CUDA graph replay per step",
                   "l2": "the step's working set (6.6 GB of saved activations / gradients of 1 048 576 points) >> 126 MB "
                         "L2, no flush needed",
                   "gradient_exchange": tr["allreduce"], "graphs_per_step": tr["graphs_per_step"],
                   "n_params": tr["n_params"], "loss_after_run": tr["loss"]},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "flop_per_step": flops_step, "flop_per_point": TRAIN_FLOP_PER_POINT,
                     "points_per_step": TRAIN_RAYS * EVALS_PER_RAY,
                     "frac_of_burst_peak": (achieved / float(burst)) if burst else None,
                     "scope": "whole step (all kernels of one graph replay) against the algorithmic MLP flops; "
                              "per-kernel times under `kernels`"},
        "e2e": {"value": world * TRAIN_RAYS / (tr["e2e_ms"] * 1e-3), "unit": "rays/s", "ms_per_step": tr["e2e_ms"],
                "h2d_bytes_per_step": tr["h2d"], "d2h_bytes_per_step": tr["d2h"],
                "api": "nfs_b200.pipeline.GraphedTrainStep.__call__(rays_o, rays_d, target) with pinned host tensors; "
                       "the loss is read back to pinned host memory every step"},
        "gpu_launches": tr["launches_per_step"] * args.steps,
        "launches_per_step": tr["launches_per_step"],
        "clocks": tr["clocks"],
    }
    if "kernels" in tr:
        line["kernels"] = tr["kernels"]
        line["kernels_total_us"] = tr["kernels_total_us"]
    traffic_file = os.path.join(ROOT, "profiles", "traffic_step.json")
    if os.path.exists(traffic_file):
        try:
            line["roofline"]["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_step")
        except Exception:
            pass
    if not args.no_cpu_baseline and not args.quick:
        threads = os.cpu_count() or 1
        rate, sec = cpu_train_step_rate(256, 8, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": "8 steps over 256 of the 4096 rays x (64 + 192) evaluations, %.2f s per step" % sec}
        if extras is not None:
            for name, fn, sample in (("composite_cfg2", lambda: cpu_composite_rate(1 << 17, 4, threads), "4 passes over 131072 of 1048576 rays x 64, fwd + autograd bwd"),
                                     ("dino_nerf_cfg4", lambda: cpu_dino_rate(128, 3, threads), "3 steps over 128 of 512 rays x 64 samples"),
                                     ("render_cfg5", lambda: cpu_render_rate(1024, 2, threads), "2 passes over 1024 of 640000 rays x (64 + 192)")):
                if isinstance(extras.get(name), dict) and "error" not in extras[name]:
                    try:
                        r, s = fn()
                        extras[name]["cpu_baseline"] = {"value": r, "unit": "rays/s", "cores": threads, "kind": "port",
                                                        "sample": sample + ", %.2f s each" % s}
                    except Exception as e:
                        extras[name]["cpu_baseline"] = {"error": repr(e)[:200]}
    line["extras"] = extras
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
