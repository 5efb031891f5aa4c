"""Fused flat Adam / AdamW (nfs_adam_step) for the small NeRF MLPs, plus gradient flattening.

Replaces optim.Adam(...).step() of train.py:114-118,286 and optim.AdamW of
train_multiscale.py:61 with one kernel over one flat fp32 buffer; the nn.Parameters become
views of that buffer, so names, shapes and state_dict() are unchanged.  torch.optim works on
the drop-in modules too - this class only removes ~40 tiny launches per step and gives the
data-parallel driver a single buffer to all-reduce.
"""
import torch

from . import _lib, mlp
from ._lib import ptr
from .ops import _stream


class FusedAdam:
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam: parameters must live on a CUDA device (no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.decoupled = lr, betas, eps, weight_decay, decoupled
        self.sizes = [p.numel() for p in self.params]
        n = sum(self.sizes)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p, k in zip(self.params, self.sizes):
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)       # parameter is now a view of the flat buffer
                off += k
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.grad = torch.zeros_like(self.flat)
        # step count and learning rate live on the device so that step() can be replayed from a CUDA graph
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._state = torch.tensor([1.0, 1.0, float(lr)], device=dev, dtype=torch.float32)
        self._lr_on_device = float(lr)
        self._steps_host = 0

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """Flatten the parameters' .grad into self.grad (one concatenation kernel)."""
        parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        torch.cat(parts, out=self.grad)
        return self.grad

    @property
    def step_count(self):
        """Number of steps taken (reads the device counter: host sync; not for the hot loop)."""
        return int(self._step_dev.item())

    def sync_lr(self):
        """Push a changed learning rate (MultiStepLR: `optimizer.lr = ...`) to its device copy.  Called by the eager
        step() and by GraphedStep.replay() - a captured nfs_adam_step_dev reads the rate from the device, so the
        fill must happen outside the graph, before the replay."""
        if float(self.lr) != self._lr_on_device:
            self._state[2:3].fill_(float(self.lr))
            self._lr_on_device = float(self.lr)

    def set_lr(self, lr):
        self.lr = float(lr)
        self.sync_lr()

    def step(self, grad_scale=1.0, gathered=False):
        if not gathered:
            self.gather_grads()
        self.sync_lr()
        with torch.cuda.device(self.flat.device):
            _lib.call("nfs_adam_step_dev", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                      self.flat.numel(), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                      float(self.weight_decay), ptr(self._step_dev), ptr(self._state), float(grad_scale),
                      int(bool(self.decoupled)), _stream())
        mlp.bump_weight_epoch()          # cached bf16 operand copies are stale now

    def state_dict(self, torch_format=False):
        """Flat form (default), or - torch_format=True - the layout of torch.optim.Adam.state_dict(): per-parameter
        {'step', 'exp_avg', 'exp_avg_sq'} (slices of the flat moments) + one param_group, interchangeable with the
        reference's optimizer (train.py:114-118, 378)."""
        if not torch_format:
            return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                    "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}
        step, state, off = self.step_count, {}, 0
        for i, (p, k) in enumerate(zip(self.params, self.sizes)):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
            off += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": bool(self.decoupled), "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts both forms of state_dict() (and therefore torch.optim.Adam's)."""
        if "param_groups" in sd:
            st = sd["state"]
            off, step = 0, 0
            for i, (p, k) in enumerate(zip(self.params, self.sizes)):
                if i in st:
                    self.exp_avg[off:off + k].copy_(st[i]["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off:off + k].copy_(st[i]["exp_avg_sq"].reshape(-1))
                    step = int(float(st[i]["step"]))
                else:
                    self.exp_avg[off:off + k].zero_()
                    self.exp_avg_sq[off:off + k].zero_()
                off += k
            self._step_dev.fill_(step)
            self.set_lr(sd["param_groups"][0].get("lr", self.lr))
            return
        self._step_dev.fill_(int(sd["step"]))
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.set_lr(sd.get("lr", self.lr))
