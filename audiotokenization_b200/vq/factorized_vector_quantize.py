"""FactorizedVectorQuantize (host mirror of vq/factorized_vector_quantize.py:10-109).

Inference semantics of the reference (``self.training`` False): project to the
low-dimensional code space, L2-normalise, pick the nearest (cosine) code, look up the
RAW codebook row and project back.  One fused kernel each way (``bc_vq_encode`` /
``bc_vq_dequant``); the N x K distance matrix of the reference is never materialised.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .. import ops


class _LinearWN(nn.Module):
    """weight_norm(nn.Linear): weight_g [out,1], weight_v [out,in], bias [out]."""

    def __init__(self, in_features, out_features):
        super().__init__()
        bound = 1.0 / math.sqrt(in_features)
        v = (torch.rand(out_features, in_features) * 2 - 1) * bound
        self.weight_g = nn.Parameter(v.norm(dim=1, keepdim=True), requires_grad=False)
        self.weight_v = nn.Parameter(v, requires_grad=False)
        self.bias = nn.Parameter((torch.rand(out_features) * 2 - 1) * bound, requires_grad=False)

    def folded(self):
        key = (self.weight_g._version, self.weight_v._version, self.bias._version, self.weight_v.data_ptr(),
               self.weight_v.device)
        c = getattr(self, "_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                v, g = self.weight_v.detach().double(), self.weight_g.detach().double()
                w = (v * (g / v.norm(dim=1, keepdim=True))).float().contiguous()
                c = (key, w, self.bias.detach().float().contiguous())
            self._cache = c
        return c[1], c[2]

    @property
    def weight(self):
        return self.folded()[0]


class FactorizedVectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, codebook_dim, commitment, **kwargs):
        super().__init__()
        self.dim = dim
        self.codebook_size = codebook_size
        self.codebook_dim = codebook_dim
        self.commitment = commitment
        if dim != self.codebook_dim:
            self.in_proj = _LinearWN(dim, self.codebook_dim)
            self.out_proj = _LinearWN(self.codebook_dim, dim)
        else:
            self.in_proj = nn.Identity()
            self.out_proj = nn.Identity()
        self._codebook = nn.Embedding(codebook_size, self.codebook_dim)
        self._codebook.weight.requires_grad_(False)

    @property
    def codebook(self):
        return self._codebook

    # ---- cached device-side parameter views ---------------------------------
    def _codebooks(self):
        w = self._codebook.weight
        key = (w._version, w.data_ptr(), w.device)
        c = getattr(self, "_cb_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                raw = w.detach().float().contiguous()
                # F.normalize(codebook): row / max(||row||, 1e-12)   (factorized_vector_quantize.py:99)
                norm = raw / raw.norm(dim=1, keepdim=True).clamp_min(1e-12)
            c = (key, raw, norm.contiguous())
            self._cb_cache = c
        return c[1], c[2]

    def _proj(self, which):
        m = getattr(self, which)
        return (None, None) if isinstance(m, nn.Identity) else m.folded()

    # ---- channels-last core ---------------------------------------------------
    def encode_cl(self, z_cl, want_margin=False, want_ze=False):
        """z_cl [B,T,C] -> (idx int32 [B,T], margin | None, z_e [B,T,D] | None)."""
        B, T, C = z_cl.shape
        if self.training:
            raise NotImplementedError("FactorizedVectorQuantize: only eval-mode (inference) is implemented")
        w_in, b_in = self._proj("in_proj")
        _, cbn = self._codebooks()
        idx, margin, z_e = ops.vq_encode(z_cl.reshape(B * T, C), w_in, b_in, cbn, want_margin, want_ze)
        return (idx.view(B, T), None if margin is None else margin.view(B, T),
                None if z_e is None else z_e.view(B, T, -1))

    def dequant_cl(self, idx, *, z_q=None, residual=None, proj=True, check_range=False):
        """idx [B,T] -> z_q [B,T,C] (or [B,T,D] when ``proj`` is False)."""
        B, T = idx.shape
        cb, _ = self._codebooks()
        w_out, b_out = self._proj("out_proj") if proj else (None, None)
        C = self.dim if (proj and w_out is not None) else self.codebook_dim
        out = ops.vq_dequant(idx.reshape(-1), cb, w_out, b_out, C,
                             z_q=None if z_q is None else z_q.view(B * T, C),
                             residual=None if residual is None else residual.view(B * T, C),
                             check_range=check_range)
        return out.view(B, T, C)

    # ---- reference API ----------------------------------------------------------
    @torch.no_grad()
    def forward(self, z):
        """z [B,D,T] -> (z_q [B,D,T], indices int64 [B,T], commit_loss [B] zeros)."""
        z_cl = ops.to_channels_last(z)
        idx, _, _ = self.encode_cl(z_cl)
        z_q = self.dequant_cl(idx)
        commit_loss = torch.zeros(z.shape[0], device=z.device)
        return z_q.permute(0, 2, 1), idx.long(), commit_loss

    @torch.no_grad()
    def decode_latents(self, latents):
        """latents [B,D_code,T] (already projected) -> (z_q [B,D_code,T] raw codes, indices [B,T])."""
        lat_cl = ops.to_channels_last(latents)
        B, T, D = lat_cl.shape
        _, cbn = self._codebooks()
        idx, _, _ = ops.vq_encode(lat_cl.reshape(B * T, D), None, None, cbn)
        idx = idx.view(B, T)
        return self.decode_code(idx.long()), idx.long()

    @torch.no_grad()
    def vq2emb(self, vq, proj=True):
        """vq int [B,T] -> [B,T,C] channel-last embedding (factorized_vector_quantize.py:78-82)."""
        return self.dequant_cl(vq.to(torch.int32), proj=proj, check_range=True)

    def get_emb(self):
        return self.codebook.weight

    @torch.no_grad()
    def embed_code(self, embed_id):
        return self.dequant_cl(embed_id.to(torch.int32), proj=False, check_range=True)

    def decode_code(self, embed_id):
        return self.embed_code(embed_id).transpose(1, 2)
