"""SnakeBeta activation (host mirror of vq/activations.py:62-119 of the reference).

Holds the reference's parameters (``alpha``, ``beta``; log-scale by default in the
codec) and evaluates on the GPU through ``bc_snake_fwd``; no PyTorch arithmetic.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class SnakeBeta(nn.Module):
    """y = x + 1/(b + 1e-9) * sin^2(x a);  a = exp(alpha), b = exp(beta) when ``alpha_logscale``."""

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = nn.Parameter(init.clone() * alpha, requires_grad=False)
        self.beta = nn.Parameter(init.clone() * alpha, requires_grad=False)
        self.no_div_by_zero = 0.000000001
        self._cache = None

    def device_params(self):
        """(a, 1/(b+eps)) as float32 device tensors, cached until alpha/beta change."""
        key = (self.alpha._version, self.beta._version, self.alpha.data_ptr(), self.beta.data_ptr(), self.alpha.device)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                a = self.alpha.detach().float()
                b = self.beta.detach().float()
                if self.alpha_logscale:
                    a, b = torch.exp(a), torch.exp(b)
                ib = 1.0 / (b + self.no_div_by_zero)
            self._cache = (key, a.contiguous(), ib.contiguous())
        return self._cache[1], self._cache[2]

    def forward_cl(self, x_cl):
        a, ib = self.device_params()
        return ops.snake(x_cl, a, ib)

    @torch.no_grad()
    def forward(self, x):  # [B, C, T]
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))
