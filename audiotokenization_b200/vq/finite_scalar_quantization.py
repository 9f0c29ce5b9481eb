"""FSQ (host mirror of vq/vector_quantize_pytorch_lucidrains/finite_scalar_quantization.py:55-259), the quantizer
``BigCodecDecoder(fsq=True)`` selects (vq/codec_decoder.py:41-47,87-89).

Inference semantics of the reference for the configuration the decoder builds -- ``FSQ(levels, channel_first=True,
dim=in_channels)``: one codebook, projections ``nn.Linear(dim, len(levels))`` / ``nn.Linear(len(levels), dim)``
(state-dict keys ``project_in.weight/bias``, ``project_out.weight/bias``; every buffer is non-persistent), float32
quantisation.  ``forward`` = project -> bound (tanh) -> round -> mixed-radix index -> project back, one fused kernel
each way (``bc_fsq_encode`` / ``bc_vq_dequant`` over the implicit codebook).  Options the decoder never sets raise.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from .. import ops


class FSQ(nn.Module):
    def __init__(self, levels: List[int], dim: Optional[int] = None, num_codebooks=1, keep_num_codebooks_dim=None,
                 scale=None, allowed_dtypes=(torch.float32, torch.float64), channel_first: bool = False,
                 projection_has_bias: bool = True, return_indices=True, force_quantization_f32=True,
                 preserve_symmetry: bool = False, noise_approx_prob=0.0):
        super().__init__()
        if num_codebooks != 1 or preserve_symmetry or noise_approx_prob != 0.0 or not return_indices or keep_num_codebooks_dim:
            raise NotImplementedError("FSQ: only the configuration BigCodecDecoder builds is implemented (one codebook, "
                                      "plain tanh bound, indices returned)")
        if not (1 <= len(levels) <= 8):
            raise NotImplementedError("FSQ: 1..8 levels")
        self.levels = [int(v) for v in levels]
        _levels = torch.tensor(self.levels, dtype=torch.int32)
        self.register_buffer("_levels", _levels, persistent=False)
        _basis = torch.cumprod(torch.tensor([1] + self.levels[:-1]), dim=0, dtype=torch.int32)
        self.register_buffer("_basis", _basis, persistent=False)
        self.scale = scale
        self.codebook_dim = len(levels)
        self.num_codebooks = 1
        self.effective_codebook_dim = self.codebook_dim
        self.keep_num_codebooks_dim = False
        self.dim = self.codebook_dim if dim is None else dim
        self.channel_first = channel_first
        self.has_projections = self.dim != self.effective_codebook_dim
        self.project_in = nn.Linear(self.dim, self.codebook_dim, bias=projection_has_bias) if self.has_projections else nn.Identity()
        self.project_out = nn.Linear(self.codebook_dim, self.dim, bias=projection_has_bias) if self.has_projections else nn.Identity()
        for p in self.parameters():
            p.requires_grad_(False)
        self.return_indices = True
        self.codebook_size = int(_levels.prod().item())
        self.register_buffer("implicit_codebook", self._indices_to_codes(torch.arange(self.codebook_size)), persistent=False)

    # ---- the reference's small helpers (host tensors / tiny device tensors; not on the hot path) ----
    def _scale_and_shift_inverse(self, zhat):
        half_width = self._levels // 2
        return (zhat - half_width) / half_width

    def indices_to_level_indices(self, indices):
        indices = indices.unsqueeze(-1)
        return (indices // self._basis) % self._levels

    def _indices_to_codes(self, indices):
        return self._scale_and_shift_inverse(self.indices_to_level_indices(indices))

    def _kernel_params(self):
        """[5, d] float32: half_l | offset | shift | half_width | basis, with the reference's own expressions
        (finite_scalar_quantization.py:111-116,142-148,170-175)."""
        dev = self._levels.device
        c = getattr(self, "_prm_cache", None)
        if c is None or c[0] != dev:
            eps = 1e-3
            half_l = (self._levels - 1) * (1 + eps) / 2
            offset = torch.where(self._levels % 2 == 0, 0.5, 0.0)
            shift = (offset / half_l).atanh()
            half_width = (self._levels // 2).float()
            prm = torch.stack([half_l.float(), offset.float(), shift.float(), half_width, self._basis.float()]).contiguous()
            c = (dev, prm)
            self._prm_cache = c
        return c[1]

    def _proj(self, which):
        m = getattr(self, which)
        if isinstance(m, nn.Identity):
            return None, None
        return m.weight.detach().float().contiguous(), (None if m.bias is None else m.bias.detach().float().contiguous())

    # ---- channels-last cores --------------------------------------------------------------------------------------
    def encode_cl(self, z_cl, want_codes=False, want_boundary=False):
        """z_cl [B,T,C] -> (idx int32 [B,T], codes [B,T,d] | None, boundary [B,T] | None)."""
        # (training mode differs from eval only through noise_approx_prob, which the constructor pins to 0)
        B, T, C = z_cl.shape
        w_in, b_in = self._proj("project_in")
        if w_in is not None and b_in is None:
            b_in = torch.zeros(self.codebook_dim, device=z_cl.device)
        idx, codes, boundary = ops.fsq_encode(z_cl.reshape(B * T, C), w_in, b_in, self._kernel_params(), want_codes, want_boundary)
        return (idx.view(B, T), None if codes is None else codes.view(B, T, -1),
                None if boundary is None else boundary.view(B, T))

    def dequant_cl(self, idx, check_range=False):
        """idx int32 [B,T] -> [B,T,dim]: implicit codebook row + project_out."""
        B, T = idx.shape
        w_out, b_out = self._proj("project_out")
        if w_out is not None and b_out is None:
            b_out = torch.zeros(self.dim, device=idx.device)
        C = self.dim if w_out is not None else self.codebook_dim
        out = ops.vq_dequant(idx.reshape(-1), self.implicit_codebook.float().contiguous(), w_out, b_out, C, check_range=check_range)
        return out.view(B, T, C)

    def forward_cl(self, z_cl, want_margin=False, want_zq=True):
        """Driver-level core with ResidualVQ.forward_cl's return convention: (z_q | None, idx int32 [1,B,T], boundary | None)."""
        idx, _, boundary = self.encode_cl(z_cl, want_boundary=want_margin)
        z_q = self.dequant_cl(idx) if want_zq else None
        return z_q, idx.unsqueeze(0), (boundary.unsqueeze(0) if want_margin else None)

    # ---- reference API ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, z):
        """channel_first: z [B,dim,T] -> (out [B,dim,T], indices int32 [B,T]); else z [B,T,dim] -> ([B,T,dim], [B,T])."""
        if z.dim() != 3:
            raise NotImplementedError("FSQ: [batch, dim, time] / [batch, time, dim] inputs only")
        z_cl = ops.to_channels_last(z) if self.channel_first else z.contiguous()
        if z_cl.shape[-1] != self.dim:
            raise ValueError(f"expected dimension of {self.dim} but found dimension of {z_cl.shape[-1]}")
        idx, _, _ = self.encode_cl(z_cl)
        out = self.dequant_cl(idx)
        return (out.permute(0, 2, 1) if self.channel_first else out), idx

    @torch.no_grad()
    def indices_to_codes(self, indices):
        """Inverse of the index computation + project_out (finite_scalar_quantization.py:182-201)."""
        if indices.dim() != 2:
            raise NotImplementedError("FSQ.indices_to_codes: [batch, time] indices only")
        out = self.dequant_cl(indices.to(torch.int32).contiguous(), check_range=True)
        return out.permute(0, 2, 1) if self.channel_first else out
