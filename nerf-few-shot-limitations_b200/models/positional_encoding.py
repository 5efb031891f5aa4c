"""Drop-in for /root/reference/src/models/positional_encoding.py (R3 in SURVEY.md section 8a)."""
try:
    from . import _bootstrap  # noqa: F401
except ImportError:
    import _bootstrap  # noqa: F401

import torch
import torch.nn as nn

from nfs_b200 import ops as _ops


class PositionalEncoding(nn.Module):
    """[x, sin(x f_0), cos(x f_0), ..., sin(x f_{L-1}), cos(x f_{L-1})] in one kernel.

    Constructor and attributes follow positional_encoding.py:9-18: `freq_bands` is a
    plain tensor attribute (not a buffer, so it is absent from state_dict, exactly like
    the reference), powers of two when log_sampling else linspace(1, 2^(L-1), L).
    Inputs never need gradients in any reference caller, so forward is not differentiable
    with respect to x (it raises if x.requires_grad).
    """

    def __init__(self, num_freqs=10, include_input=True, log_sampling=True):
        super().__init__()
        self.include_input = include_input
        self.num_freqs = num_freqs
        if log_sampling:
            self.freq_bands = 2. ** torch.linspace(0., num_freqs - 1, steps=num_freqs)
        else:
            self.freq_bands = torch.linspace(2. ** 0., 2. ** (num_freqs - 1), steps=num_freqs)
        self._dev_freqs = {}

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("PositionalEncoding: gradients w.r.t. coordinates are not supported "
                               "(no reference caller differentiates through the encoding)")
        key = str(x.device)
        fr = self._dev_freqs.get(key)
        if fr is None:
            fr = self.freq_bands.to(device=x.device, dtype=torch.float32)
            self._dev_freqs = {key: fr}
        return _ops.posenc(x, fr, self.include_input)
