"""BigCodecDecoder (host mirror of vq/codec_decoder.py:15-142).

Owns the quantizer and the latent -> waveform stack.  ``forward(x, vq=True)`` quantises,
``forward(x, vq=False)`` decodes -- the identity test ``vq is True`` is the reference's
(vq/codec_decoder.py:86).
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import ops
from . import activations
from .alias_free_torch import Activation1d
from .module import DecoderBlock, ResLSTM, WNConv1d, module_scope
from .residual_vq import ResidualVQ
from .finite_scalar_quantization import FSQ


class BigCodecDecoder(nn.Module):
    def __init__(self, in_channels=1024, upsample_initial_channel=1536, ngf=48, use_rnn=True,
                 rnn_bidirectional=False, rnn_num_layers=2, up_ratios=(5, 5, 2, 2, 2), dilations=(1, 3, 9),
                 causal=False, antialias=False, fsq=False, fsq_levels=[4, 4, 4, 8], vq_num_quantizers=1,
                 vq_commit_weight=0.25, vq_weight_init=False, vq_full_commit_loss=False, codebook_size=8192,
                 codebook_dim=8):
        super().__init__()
        self.hop_length = np.prod(up_ratios)
        self.ngf = ngf
        self.up_ratios = up_ratios
        self.fsq = fsq
        if fsq:   # vq/codec_decoder.py:41-47
            self.quantizer = FSQ(levels=fsq_levels, channel_first=True, dim=in_channels)
            assert codebook_size == np.prod(fsq_levels), "codebook_size must be equal to the product of fsq_levels"
        else:
            self.quantizer = ResidualVQ(num_quantizers=vq_num_quantizers, dim=in_channels, codebook_size=codebook_size,
                                        codebook_dim=codebook_dim, threshold_ema_dead_code=2,
                                        commitment=vq_commit_weight, weight_init=vq_weight_init,
                                        full_commit_loss=vq_full_commit_loss)
        channels = upsample_initial_channel
        layers = [WNConv1d(in_channels, channels, kernel_size=7, padding=3, causal=causal)]
        if use_rnn:
            layers += [ResLSTM(channels, num_layers=rnn_num_layers, bidirectional=rnn_bidirectional)]
        output_dim = channels
        for i, stride in enumerate(up_ratios):
            input_dim = channels // 2 ** i
            output_dim = channels // 2 ** (i + 1)
            layers += [DecoderBlock(input_dim, output_dim, stride, dilations, causal=causal, antialias=antialias)]
        layers += [
            Activation1d(activation=activations.SnakeBeta(output_dim, alpha_logscale=True), antialias=antialias),
            WNConv1d(output_dim, 1, kernel_size=7, padding=3, causal=causal),
            nn.Tanh(),
        ]
        self.model = nn.Sequential(*layers)
        self.precision = None     # arithmetic mode of this module's own calls (None: the enclosing precision_scope)
        self.eval()

    # ---- channels-last cores ----------------------------------------------------
    def decode_cl(self, z_cl):
        """z_cl [B,T',C] -> waveform [B,T,1]."""
        mods = list(self.model)
        with module_scope(self):
            h = mods[0].forward_cl(z_cl)
            for m in mods[1:-3]:
                h = m.forward_cl(h)
            act, conv = mods[-3], mods[-2]
            if act.antialias:
                return conv.forward_cl(act.forward_cl(h), tanh=True)
            return conv.forward_cl(h, act=act.act, tanh=True)

    # ---- reference API ------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, vq=True):
        if vq is True:
            if self.fsq:   # vq/codec_decoder.py:87-89: (x, int32 indices [B,T], zeros [B])
                x, q = self.quantizer(x)
                commit_loss = torch.zeros(x.shape[0], device=x.device)
            else:
                x, q, commit_loss = self.quantizer(x)
            return x, q, commit_loss
        return self.decode_cl(ops.to_channels_last(x)).permute(0, 2, 1)

    def vq2emb(self, vq):
        self.quantizer = self.quantizer.eval()
        return self.quantizer.vq2emb(vq)

    def get_emb(self):
        self.quantizer = self.quantizer.eval()
        return self.quantizer.get_emb()

    def inference_vq(self, vq):
        return self.forward(vq[None, :, :], vq=False)

    def inference(self, x):
        return self.forward(x, vq=False), None

    def remove_weight_norm(self):
        """No-op: the kernels always consume the folded weight."""

    def apply_weight_norm(self):
        """No-op: parameters are always stored as weight_g / weight_v."""

    def reset_parameters(self):
        """Nothing to do (see BigCodecEncoder.reset_parameters)."""
