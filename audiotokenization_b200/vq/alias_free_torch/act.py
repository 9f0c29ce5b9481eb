"""Activation1d (host mirror of vq/alias_free_torch/act.py:7-32).

``antialias=False``: plain SnakeBeta.  ``antialias=True``: 2x up FIR -> SnakeBeta ->
2x down FIR, executed as ONE fused kernel (one HBM read + one write per element).
"""
from __future__ import annotations

import torch
from torch import nn

from ... import ops
from .resample import DownSample1d, UpSample1d


class Activation1d(nn.Module):
    def __init__(self, activation, antialias: bool = False, up_ratio: int = 2, down_ratio: int = 2,
                 up_kernel_size: int = 12, down_kernel_size: int = 12):
        super().__init__()
        self.antialias = antialias
        self.up_ratio = up_ratio
        self.down_ratio = down_ratio
        self.act = activation
        if antialias:
            if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
                raise NotImplementedError("the fused anti-aliased activation supports ratio 2 / 12 taps "
                                          "(the only setting the codec uses)")
            self.upsample = UpSample1d(up_ratio, up_kernel_size)
            self.downsample = DownSample1d(down_ratio, down_kernel_size)

    def _fir(self):
        up, down = self.upsample.filter, self.downsample.lowpass.filter
        key = (up._version, down._version, up.data_ptr(), down.data_ptr())
        cache = getattr(self, "_fir_cache", None)
        if cache is None or cache[0] != key:
            if not torch.equal(up, down):
                raise NotImplementedError("up- and down-sampling FIRs differ; the fused kernel assumes the "
                                          "reference's identical 12-tap filter")
            cache = (key, up.reshape(-1).float().contiguous())
            self._fir_cache = cache
        return cache[1]

    def forward_cl(self, x_cl):
        a, ib = self.act.device_params()
        if not self.antialias:
            return ops.snake(x_cl, a, ib)
        return ops.snake(x_cl, a, ib, antialias=True, fir=self._fir())

    @torch.no_grad()
    def forward(self, x):  # [B, C, T]
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))
