"""ResidualVQ (host mirror of vq/residual_vq.py:6-53)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .factorized_vector_quantize import FactorizedVectorQuantize


class ResidualVQ(nn.Module):
    def __init__(self, *, num_quantizers, codebook_size, **kwargs):
        super().__init__()
        VQ = FactorizedVectorQuantize
        if type(codebook_size) == int:
            codebook_size = [codebook_size] * num_quantizers
        self.layers = nn.ModuleList([VQ(codebook_size=size, **kwargs) for size in codebook_size])
        self.num_quantizers = num_quantizers

    def forward_cl(self, z_cl, want_margin=False, want_zq=True):
        """z_cl [B,T,C] -> (z_q [B,T,C] | None, idx int32 [n_q,B,T], margins | None).  ``want_zq=False`` (the
        encode -> indices path of extract_indices.py, which drops the quantised latents) skips every dequantisation
        the residual loop does not need: all of them for one quantizer, the last one otherwise."""
        B, T, C = z_cl.shape
        residual = z_cl.clone() if len(self.layers) > 1 else None
        z_q = None
        all_idx, margins = [], []
        for i, layer in enumerate(self.layers):
            idx, margin, _ = layer.encode_cl(z_cl if residual is None else residual, want_margin=want_margin)
            all_idx.append(idx)
            margins.append(margin)
            if not want_zq and i == len(self.layers) - 1:
                z_q = None
                break
            # quantized_out += q ; residual -= q   (residual_vq.py:27-33) in one kernel
            z_q = layer.dequant_cl(idx, z_q=z_q, residual=residual)
        return z_q, torch.stack(all_idx), (torch.stack(margins) if want_margin else None)

    @torch.no_grad()
    def forward(self, x):
        """x [B,C,T] -> (quantized [B,C,T], indices int64 [n_q,B,T], losses [n_q] zeros)."""
        z_q, idx, _ = self.forward_cl(ops.to_channels_last(x))
        losses = torch.zeros(len(self.layers), device=x.device)
        return z_q.permute(0, 2, 1), idx.long(), losses

    @torch.no_grad()
    def vq2emb(self, vq, proj=True):
        """vq [B,T,n_q] -> [B,T,C] (channel-last, like the reference)."""
        out = None
        for i, layer in enumerate(self.layers):
            idx = vq[:, :, i].to(torch.int32).contiguous()
            if out is None:
                out = layer.dequant_cl(idx, proj=proj, check_range=True)
            else:
                out = layer.dequant_cl(idx, z_q=out, proj=proj, check_range=True)
        return out

    def get_emb(self):
        return [layer.get_emb() for layer in self.layers]
