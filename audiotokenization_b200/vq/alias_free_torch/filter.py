"""Kaiser-windowed sinc low-pass taps (host mirror of vq/alias_free_torch/filter.py).

Only parameter holders live here: the filtering itself is fused into the
anti-aliased activation kernel (``bc_snake_fwd`` with ``antialias=1``).
"""
from __future__ import annotations

import math

import torch
from torch import nn


def kaiser_sinc_filter1d(cutoff, half_width, kernel_size):
    """[1,1,kernel_size] float32 taps; same closed form as the reference (filter.py:28-57)."""
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    att = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if att > 50.0:
        beta = 0.1102 * (att - 8.7)
    elif att >= 21.0:
        beta = 0.5842 * (att - 21) ** 0.4 + 0.07886 * (att - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half_size, half_size) + 0.5
    else:
        time = torch.arange(kernel_size) - half_size
    if cutoff == 0:
        return torch.zeros(1, 1, kernel_size)
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    filt = filt / filt.sum()
    return filt.view(1, 1, kernel_size)


class LowPassFilter1d(nn.Module):
    """Parameter holder for the down-sampling FIR (``filter`` buffer, filter.py:60-84)."""

    def __init__(self, cutoff=0.5, half_width=0.6, stride: int = 1, padding: bool = True,
                 padding_mode: str = "replicate", kernel_size: int = 12):
        super().__init__()
        if cutoff < -0.0:
            raise ValueError("Minimum cutoff must be larger than zero.")
        if cutoff > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")
        self.kernel_size = kernel_size
        self.even = kernel_size % 2 == 0
        self.pad_left = kernel_size // 2 - int(self.even)
        self.pad_right = kernel_size // 2
        self.stride = stride
        self.padding = padding
        self.padding_mode = padding_mode
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))
