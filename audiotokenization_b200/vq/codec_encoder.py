"""BigCodecEncoder (host mirror of vq/codec_encoder.py:14-90).

waveform [B,1,T] -> conv stem -> EncoderBlocks -> [ResLSTM] -> SnakeBeta -> conv ->
latents [B,out_channels,T'].  Same constructor, attributes and state-dict keys as the
reference; the arithmetic runs in the sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .. import ops
from . import activations
from .alias_free_torch import Activation1d
from .module import EncoderBlock, ResLSTM, WNConv1d, _act_conv, module_scope


class BigCodecEncoder(nn.Module):
    def __init__(self, ngf=48, use_rnn=True, rnn_bidirectional=False, causal=False, antialias=False,
                 rnn_num_layers=2, up_ratios=(2, 2, 2, 5, 5), dilations=(1, 3, 9), out_channels=1024):
        super().__init__()
        self.hop_length = np.prod(up_ratios)
        self.ngf = ngf
        self.up_ratios = up_ratios
        if causal:
            assert not rnn_bidirectional
        d_model = ngf
        block = [WNConv1d(1, d_model, kernel_size=7, padding=3, causal=causal)]
        for stride in up_ratios:
            d_model *= 2
            block += [EncoderBlock(d_model, stride=stride, dilations=dilations, causal=causal, antialias=antialias)]
        if use_rnn:
            block += [ResLSTM(d_model, num_layers=rnn_num_layers, bidirectional=rnn_bidirectional)]
        block += [
            Activation1d(activation=activations.SnakeBeta(d_model, alpha_logscale=True), antialias=antialias),
            WNConv1d(d_model, out_channels, kernel_size=3, padding=1, causal=causal),
        ]
        self.block = nn.Sequential(*block)
        self.enc_dim = d_model
        self.precision = None     # arithmetic mode of this module's own calls (None: the enclosing precision_scope)
        self.eval()

    def _split(self):
        mods = list(self.block)
        n_front = len(mods) - 2 - (1 if isinstance(mods[-3], ResLSTM) else 0)
        return mods[:n_front], mods[n_front:-2], mods[-2], mods[-1]

    def front_cl(self, x_cl):
        """Conv stem + EncoderBlocks: x_cl [B,T,1] -> frame-rate features [B,T',enc_dim].  This part is
        independent per utterance AND per time tile, so it is run in small micro-batches."""
        front, _, _, _ = self._split()
        with module_scope(self):
            h = front[0].forward_cl(x_cl)
            for m in front[1:]:
                h = m.forward_cl(h)
        return h

    # The deep end of the conv stack has few 128-frame tiles per utterance (94 at 256 channels, 19 after the last
    # strided conv for a 30 s clip): at a micro-batch of 8 that is 5.08 tiles per SM, i.e. a sixth round for 8 % of
    # the SMs.  So the front end is scheduled in two stages: `front_shallow_cl` on small micro-batches (bounded
    # activation memory), `front_deep_cl` on several of their outputs at once (full last round).
    def _front_stages(self):
        front, _, _, _ = self._split()
        stem, blocks = front[0], front[1:]
        k = max(len(blocks) - 2, 0)          # blocks[:k] whole; blocks[k]: units shallow, strided conv deep; the rest deep
        return stem, blocks[:k], blocks[k], blocks[k + 1:]

    def front_shallow_cl(self, x_cl, out=None):
        """Stem + EncoderBlocks up to the ResidualUnits of the second-to-last block; ``out`` = destination slice."""
        stem, whole, pivot, _ = self._front_stages()
        with module_scope(self):
            h = stem.forward_cl(x_cl)
            for m in whole:
                h = m.forward_cl(h)
            return pivot.units_cl(h, out=out)

    def front_deep_cl(self, h):
        """The pivot block's strided conv + the last block -> frame-rate features."""
        _, _, pivot, rest = self._front_stages()
        with module_scope(self):
            h = pivot.down_cl(h)
            for m in rest:
                h = m.forward_cl(h)
        return h

    def back_cl(self, h):
        """[ResLSTM] + SnakeBeta + final conv on frame-rate features.  The LSTM is sequential in time, so
        it is run over as many utterances at once as possible (its cost per step is almost flat in B)."""
        _, rnn, act, conv = self._split()
        with module_scope(self):
            for m in rnn:
                h = m.forward_cl(h)
            return _act_conv(act, conv, h)

    def forward_cl(self, x_cl):
        """x_cl [B,T,1] -> latents [B,T',out_channels] (channels-last)."""
        return self.back_cl(self.front_cl(x_cl))

    @torch.no_grad()
    def forward(self, x):
        """x float32 [B,1,T] -> [B,out_channels,T'] (a channels-last-strided view; values as the reference)."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"BigCodecEncoder expects [B,1,T], got {tuple(x.shape)}")
        return self.forward_cl(ops.to_channels_last(x)).permute(0, 2, 1)

    def inference(self, x):
        return self.forward(x)

    # weight-norm is folded once at load time; these exist for API compatibility
    def remove_weight_norm(self):
        """No-op: the kernels always consume the folded weight (g*v/||v||)."""

    def apply_weight_norm(self):
        """No-op: parameters are always stored as weight_g / weight_v."""

    def reset_parameters(self):
        """The reference re-initialises a derived attribute here (SURVEY.md 8c hazard 1); nothing to do."""
