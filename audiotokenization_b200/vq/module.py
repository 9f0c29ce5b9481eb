"""Conv / residual / LSTM building blocks (host mirror of vq/module.py:11-167).

Same class names, constructor arguments, parameter names and shapes as the
reference, so a reference ``state_dict`` loads unchanged (old-style weight-norm
triples ``weight_g / weight_v / bias``).  ``forward`` takes and returns the
reference's ``[B, C, T]`` tensors; internally everything runs channels-last
through the C ABI (``forward_cl``), with SnakeBeta fused into the consuming conv
and bias / residual / tanh fused into its epilogue.

Inference only: parameters do not require grad and no autograd graph is built.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import ops
from . import activations
from .alias_free_torch import Activation1d

def _os_env_flag(name: str, default: bool) -> bool:
    import os
    v = os.environ.get(name)
    return default if v is None else v not in ("0", "false", "False", "")


_PRECISION = ["fp32"]
LSTM_TENSOR_CORE = [True]   # tensor-core modes: run the LSTM recurrence on tcgen05 when H allows
LSTM_WAVEFRONT = [_os_env_flag("BC_LSTM_WAVEFRONT", True)]     # ... and, for batches whose two layers fit the chip side by side, as a two-layer wave front
LSTM_WAVEFRONT_CHUNK = [128]   # steps per chunk of the wave front


def _cabi_sm_count() -> int:
    from .. import _cabi
    c = _cabi.__dict__.setdefault("_SM_COUNT", {})
    dev = torch.cuda.current_device()
    if dev not in c:
        c[dev] = _cabi.device_info(dev)["sm_count"]
    return c[dev]
FUSE_RESUNIT = [True]   # tensor-core modes: run a whole ResidualUnit as one kernel when the geometry allows
STREAM = [True]         # tensor-core modes: wide layers run on the persistent streamed-weight kernel
STREAM_PAIR = [True]    # ... in its CTA-pair form (tcgen05 cta_group::2) where the geometry has one
CONVTR_STREAM = [True]  # ... and transposed convs as ONE launch of it (all phases as channel blocks)
import os as _os
STREAM_MIN_CIN = [int(_os.environ.get("BC_STREAM_MIN_CIN", "32"))]    # ... when the layer has at least this many input channels
STREAM_RU_MIN_C = [int(_os.environ.get("BC_STREAM_RU_MIN_C", "128"))]  # ... ResidualUnits: narrower ones keep their weights resident (ru_persist)


def _stream_tile(c_in, c_out, k, stride, dilation, precision, fused=False):
    """n_tile of the streamed-weight kernel when policy and geometry select it, else None."""
    if precision == "fp32" or not STREAM[0]:
        return None
    if c_in < (STREAM_RU_MIN_C[0] if fused else STREAM_MIN_CIN[0]):
        return None
    return ops.stream_plan(c_in, c_out, k, stride, dilation, precision, fused)


def set_precision(mode: str) -> None:
    """Arithmetic mode of the dense contractions: 'fp32' (CUDA cores, exact float32),
    'bf16' (tcgen05 single pass) or 'bf16x3' (tcgen05 split-precision)."""
    if mode not in ops.PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}")
    _PRECISION[0] = mode


def get_precision() -> str:
    return _PRECISION[0]


class precision_scope:
    """``with precision_scope("bf16x3"): ...`` -- set the arithmetic mode for a block and restore it after."""

    def __init__(self, mode: str):
        if mode not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}")
        self.mode = mode

    def __enter__(self):
        self.prev = _PRECISION[0]
        _PRECISION[0] = self.mode
        return self

    def __exit__(self, *exc):
        _PRECISION[0] = self.prev
        return False


class module_scope:
    """Arithmetic mode of a top-level module call: ``BigCodecEncoder`` / ``BigCodecDecoder`` carry their own
    ``precision`` attribute (set by ``BigCodecModel``), so the reference's call pattern
    ``lm.model['CodecEnc'](x)`` -- calling the sub-module directly -- runs in the mode the model was built with
    instead of silently falling back to the process-wide default.  ``None`` keeps the enclosing scope's mode."""

    def __init__(self, module):
        self.mode = getattr(module, "precision", None)
        if self.mode is not None and self.mode not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}")

    def __enter__(self):
        self.prev = _PRECISION[0]
        if self.mode is not None:
            _PRECISION[0] = self.mode
        return self

    def __exit__(self, *exc):
        _PRECISION[0] = self.prev
        return False


def _fold(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """weight_norm(dim=0): w = g * v / ||v|| with the norm over all dims but 0 (float64 fold, once per load)."""
    v64, g64 = v.detach().double(), g.detach().double()
    norm = v64.flatten(1).norm(dim=1).view([-1] + [1] * (v64.dim() - 1))
    return v64 * (g64 / norm)


class _WNParams(nn.Module):
    """weight_g / weight_v / bias holder with a packed-weight cache."""

    def _key(self):
        return (self.weight_g._version, self.weight_v._version, self.bias._version, self.weight_v.data_ptr(),
                self.weight_v.device)

    def _cached(self, build):
        key = self._key()
        c = getattr(self, "_pack_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                c = (key,) + tuple(build())
            self._pack_cache = c
        return c[1:]

    @property
    def weight(self):
        """Folded weight in the reference's layout (derived, like old-style weight_norm's ``.weight``)."""
        return _fold(self.weight_g, self.weight_v).float()


class _Conv1dWN(_WNParams):
    """weight_norm(nn.Conv1d) (vq/module.py:59-65).  ``left_pad`` overrides symmetric padding (causal)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, left_pad=None):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.left_pad = padding if left_pad is None else left_pad
        self.total_pad = 2 * padding if left_pad is None else left_pad
        bound = 1.0 / math.sqrt(in_channels * kernel_size)
        v = (torch.rand(out_channels, in_channels, kernel_size) * 2 - 1) * bound
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1), requires_grad=False)
        self.weight_v = nn.Parameter(v, requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(out_channels), requires_grad=False)

    def out_length(self, t_in: int) -> int:
        return (t_in + self.total_pad - self.dilation * (self.kernel_size - 1) - 1) // self.stride + 1

    def packed(self):
        def build():
            w = _fold(self.weight_g, self.weight_v)                 # [out, in, k]
            return (w.permute(2, 1, 0).contiguous().float(),        # [k, in, out]
                    self.bias.detach().float().contiguous())
        return self._cached(build)

    def packed_for(self, precision: str):
        """(weight image, bias, effective precision): the bf16 tensor-core image when the geometry has a
        tcgen05 tiling, else the fp32 [k,in,out] array (edge convs stay on the CUDA-core kernel)."""
        w, b = self.packed()
        plan = ops.tc_plan(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.dilation,
                           precision)
        if plan is None:
            return w, b, "fp32"
        cache = self.__dict__.setdefault("_tc_cache", {})
        key = (precision, self._key())
        if cache.get("key") != key:
            cache.clear()
            cache.update(key=key, w=ops.pack_tc_weight(w, plan, precision))
        return cache["w"], b, precision

    def stream_image(self, precision: str, n_tile: int, pair: bool = False):
        cache = self.__dict__.setdefault("_stream_cache", {})
        key = (precision, n_tile, pair, self._key())
        if cache.get("key") != key:
            pack = ops.pack_stream_weight_pair if pair else ops.pack_stream_weight
            cache.clear()
            cache.update(key=key, w=pack(self.packed()[0], n_tile, precision))
        return cache["w"]

    def forward_cl(self, x_cl, act: Optional[activations.SnakeBeta] = None, res=None, tanh=False):
        a = ib = None
        if act is not None:
            a, ib = act.device_params()
        precision = get_precision()
        nt = _stream_tile(self.in_channels, self.out_channels, self.kernel_size, self.stride, self.dilation, precision)
        if nt is not None:
            pair = STREAM_PAIR[0] and ops.stream_pair_ok(self.in_channels, self.out_channels, self.kernel_size, self.stride,
                                                         self.dilation, precision)
            return ops.conv1d_stream(x_cl, self.stream_image(precision, nt, pair), self.packed()[1], k=self.kernel_size,
                                     c_out=self.out_channels, stride=self.stride, dilation=self.dilation,
                                     pad_left=self.left_pad, t_out=self.out_length(x_cl.shape[1]), snake_a=a,
                                     snake_ib=ib, res=res, tanh=tanh, precision=precision, pair=pair)
        w, b, prec = self.packed_for(precision)
        return ops.conv1d(x_cl, w, b, stride=self.stride, dilation=self.dilation, pad_left=self.left_pad,
                          t_out=self.out_length(x_cl.shape[1]), snake_a=a, snake_ib=ib, res=res, tanh=tanh,
                          precision=prec)

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class _ConvTranspose1dWN(_WNParams):
    """weight_norm(nn.ConvTranspose1d) (vq/module.py:67-72): weight [in, out, k], norm per INPUT channel."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0, causal_trim=0):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.stride, self.padding, self.output_padding, self.causal_trim = stride, padding, output_padding, causal_trim
        bound = 1.0 / math.sqrt(out_channels * kernel_size)
        v = (torch.rand(in_channels, out_channels, kernel_size) * 2 - 1) * bound
        self.weight_g = nn.Parameter(v.flatten(1).norm(dim=1).view(-1, 1, 1), requires_grad=False)
        self.weight_v = nn.Parameter(v, requires_grad=False)
        self.bias = nn.Parameter((torch.rand(out_channels) * 2 - 1) * bound, requires_grad=False)
        s, k = stride, kernel_size
        t_out_minus = -s - 2 * padding + k + output_padding - causal_trim   # T_out = T_in*s + this
        if s == 1 or k != 2 * s or t_out_minus != 0:
            raise NotImplementedError(
                f"ConvTranspose1d(k={k}, stride={s}, padding={padding}, output_padding={output_padding}): only the "
                "codec's k = 2*stride up-sampling geometry (T_out = T_in*stride) is implemented")

    def packed(self):
        def build():
            w = _fold(self.weight_g, self.weight_v)                 # [in, out, k]
            s, p = self.stride, self.padding
            phases = []
            for ph in range(s):
                j0 = (ph + p) % s
                phases.append(torch.stack([w[:, :, j0 + s], w[:, :, j0]]))   # tap0 -> x[m+q-1], tap1 -> x[m+q]
            return (torch.stack(phases).contiguous().float(),       # [s, 2, in, out]
                    self.bias.detach().float().contiguous())
        return self._cached(build)

    def packed_for(self, precision: str):
        w, b = self.packed()
        plan = ops.tc_plan(self.in_channels, self.out_channels, 2, 1, 1, precision)
        if plan is None:
            return w, b, "fp32"
        cache = self.__dict__.setdefault("_tc_cache", {})
        key = (precision, self._key())
        if cache.get("key") != key:
            cache.clear()
            cache.update(key=key, w=torch.stack([ops.pack_tc_weight(w[ph], plan, precision)
                                                 for ph in range(self.stride)]).contiguous())
        return cache["w"], b, precision

    def stream_image(self, precision: str):
        """(image, tiled bias, three_tap) for the single-launch streamed-weight form, or None when the geometry has no
        plan (narrow / irregular layers keep the per-phase path)."""
        nt = _stream_tile(self.in_channels, self.stride * self.out_channels, 2, 1, 1, precision)
        if nt is None:
            return None
        if self.out_channels % nt != 0:       # n-tiles straddle phases: three-tap zero-padded form needs its own plan
            nt = _stream_tile(self.in_channels, self.stride * self.out_channels, 3, 1, 1, precision)
            if nt is None:
                return None
        three = self.out_channels % nt != 0
        pair = STREAM_PAIR[0] and ops.stream_pair_ok(self.in_channels, self.stride * self.out_channels, 3 if three else 2, 1, 1, precision)
        cache = self.__dict__.setdefault("_stream_cache", {})
        key = (precision, nt, pair, self._key())
        if cache.get("key") != key:
            w, b = self.packed()
            img, three = ops.pack_convtr_stream_weight(w, self.stride, self.padding, nt, precision, pair)
            cache.clear()
            cache.update(key=key, w=img, b=b.repeat(self.stride).contiguous(), three=three, pair=pair)
        return cache["w"], cache["b"], cache["three"], cache["pair"]

    def forward_cl(self, x_cl, act: Optional[activations.SnakeBeta] = None):
        precision = get_precision()
        a = ib = None
        if act is not None:
            a, ib = act.device_params()
        st = self.stream_image(precision) if CONVTR_STREAM[0] else None
        if st is not None:
            return ops.conv_transpose1d_stream(x_cl, st[0], st[1], stride=self.stride, padding=self.padding,
                                               c_out=self.out_channels, three_tap=st[2], snake_a=a, snake_ib=ib,
                                               precision=precision, pair=st[3])
        w, b, prec = self.packed_for(precision)
        return ops.conv_transpose1d(x_cl, w, b, stride=self.stride, padding=self.padding, snake_a=a, snake_ib=ib,
                                    precision=prec, c_out=self.out_channels)

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class CausalConv1d(nn.Module):
    """Left-padded conv (vq/module.py:11-48); parameters live under ``.conv`` like the reference."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=0, stride=1, dilation=1, groups=1, bias=True,
                 padding_mode="zeros", device=None, dtype=None):
        super().__init__()
        if groups != 1 or not bias or padding_mode != "zeros":
            raise NotImplementedError("CausalConv1d: only groups=1, bias=True, zero padding are on the hot path")
        self.padding = (kernel_size - stride) * dilation
        self.conv = _Conv1dWN(in_channels, out_channels, kernel_size, stride=stride, padding=0, dilation=dilation,
                              left_pad=self.padding)

    def forward_cl(self, x_cl, act=None, res=None, tanh=False):
        return self.conv.forward_cl(x_cl, act=act, res=res, tanh=tanh)

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class CausalConvTranspose1d(nn.Module):
    """ConvTranspose1d without padding, last ``stride`` samples dropped (vq/module.py:50-57)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, bias=True, device=None, dtype=None):
        super().__init__()
        if not bias:
            raise NotImplementedError("CausalConvTranspose1d: bias=False is not on the hot path")
        self.stride = stride
        self.conv = _ConvTranspose1dWN(in_channels, out_channels, kernel_size, stride=stride, padding=0,
                                       output_padding=0, causal_trim=stride)

    def forward_cl(self, x_cl, act=None):
        return self.conv.forward_cl(x_cl, act=act)

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


def WNConv1d(*args, causal=False, **kwargs):
    """Factory with the reference's signature (vq/module.py:59-65)."""
    if causal:
        return CausalConv1d(*args, **kwargs)
    return _Conv1dWN(*args, **kwargs)


def WNConvTranspose1d(*args, causal=False, **kwargs):
    """Factory with the reference's signature (vq/module.py:67-72)."""
    if causal:
        return CausalConvTranspose1d(*args, **kwargs)
    return _ConvTranspose1dWN(*args, **kwargs)


def _act_conv(act: Activation1d, conv, x_cl, **kw):
    """Activation1d followed by a conv: fused prologue when not anti-aliased."""
    if act.antialias:
        return conv.forward_cl(act.forward_cl(x_cl), **kw)
    return conv.forward_cl(x_cl, act=act.act, **kw)


class ResidualUnit(nn.Module):
    """x + conv1(snake(conv7_dilated(snake(x))))  (vq/module.py:74-89)."""

    def __init__(self, dim: int = 16, dilation: int = 1, causal: bool = False, antialias: bool = False):
        super().__init__()
        pad = 0 if causal else ((7 - 1) * dilation) // 2
        self.block = nn.Sequential(
            Activation1d(activation=activations.SnakeBeta(dim, alpha_logscale=True), antialias=antialias),
            WNConv1d(dim, dim, kernel_size=7, dilation=dilation, padding=pad, causal=causal),
            Activation1d(activation=activations.SnakeBeta(dim, alpha_logscale=True), antialias=antialias),
            WNConv1d(dim, dim, kernel_size=1),
        )

    def _fused_plan(self, precision):
        """Geometry of the one-kernel path, or None (fp32 mode, anti-aliased activations, C > 128, ...)."""
        if precision == "fp32" or not FUSE_RESUNIT[0] or self.block[0].antialias:
            return None
        conv7 = self.block[1].conv if isinstance(self.block[1], CausalConv1d) else self.block[1]
        conv1 = self.block[3]
        C = conv7.in_channels
        plan = ops.resunit_plan(C, conv7.kernel_size, conv7.dilation, precision)
        if plan is None:
            return None
        return conv7, conv1, plan[0], (C, C // 16, 1), plan[1]

    def _stream_forward(self, x_cl, prec, out=None):
        """Whole unit on the streamed-weight kernel (wide layers), or None."""
        if not FUSE_RESUNIT[0] or self.block[0].antialias:
            return None
        conv7 = self.block[1].conv if isinstance(self.block[1], CausalConv1d) else self.block[1]
        conv1 = self.block[3]
        C = conv7.in_channels
        nt = _stream_tile(C, C, conv7.kernel_size, 1, conv7.dilation, prec, fused=True)
        if nt is None:
            return None
        sa1, sib1 = self.block[0].act.device_params()
        sa2, sib2 = self.block[2].act.device_params()
        pair = STREAM_PAIR[0] and ops.stream_pair_ok(C, C, conv7.kernel_size, 1, conv7.dilation, prec, fused=True)
        return ops.resunit_stream(x_cl, conv7.stream_image(prec, nt, pair), conv7.packed()[1], sa1, sib1,
                                  conv1.stream_image(prec, nt, pair), conv1.packed()[1], sa2, sib2, k=conv7.kernel_size,
                                  dilation=conv7.dilation, pad_left=conv7.left_pad, precision=prec, out=out, pair=pair)

    def forward_cl(self, x_cl, out=None):
        """``out``: optional contiguous [B,T,C] destination (a slice of a larger batch buffer)."""
        prec = get_precision()
        y = self._stream_forward(x_cl, prec, out)
        if y is not None:
            return y
        y = self._forward_narrow(x_cl, prec)
        if out is None:
            return y
        out.copy_(y)
        return out

    def _forward_narrow(self, x_cl, prec):
        fused = self._fused_plan(prec)
        if fused is None:
            h = _act_conv(self.block[0], self.block[1], x_cl)
            return _act_conv(self.block[2], self.block[3], h, res=x_cl)
        conv7, conv1, plan7, plan1, kind = fused
        cache = self.__dict__.setdefault("_ru_cache", {})
        key = (prec, plan7, kind, conv7._key(), conv1._key())
        if cache.get("key") != key:
            cache.clear()
            if kind in (3, 4):    # CTA-pair kernel: per-rank images (half of every B operand per SM)
                cache.update(key=key, w7=ops.pack_pair_weights(conv7.packed()[0], stacked=(kind == 4)),
                             w1=ops.pack_pair_weights(conv1.packed()[0], stacked=False))
            else:
                cache.update(key=key, w7=ops.pack_tc_weight(conv7.packed()[0], plan7, prec),
                             w1=ops.pack_tc_weight(conv1.packed()[0], plan1, prec))
        sa1, sib1 = self.block[0].act.device_params()
        sa2, sib2 = self.block[2].act.device_params()
        return ops.resunit(x_cl, cache["w7"], conv7.packed()[1], sa1, sib1, cache["w1"], conv1.packed()[1], sa2, sib2,
                           k=conv7.kernel_size, dilation=conv7.dilation, pad_left=conv7.left_pad, precision=prec)

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class EncoderBlock(nn.Module):
    """3 ResidualUnits -> snake -> strided conv (vq/module.py:91-113)."""

    def __init__(self, dim: int = 16, stride: int = 1, dilations=(1, 3, 9), causal: bool = False,
                 antialias: bool = False):
        super().__init__()
        runits = [ResidualUnit(dim // 2, dilation=d, causal=causal, antialias=antialias) for d in dilations]
        pad = 0 if causal else (stride // 2 + stride % 2 if stride != 1 else 0)
        self.block = nn.Sequential(
            *runits,
            Activation1d(activation=activations.SnakeBeta(dim // 2, alpha_logscale=True), antialias=antialias),
            WNConv1d(dim // 2, dim, kernel_size=2 * stride if stride != 1 else 1, stride=stride, padding=pad,
                     causal=causal),
        )

    def units_cl(self, x_cl, out=None):
        """The ResidualUnits (time resolution of the block's input); the last one may write into ``out``."""
        units = list(self.block)[: len(self.block) - 2]
        for i, ru in enumerate(units):
            x_cl = ru.forward_cl(x_cl, out=out if i == len(units) - 1 else None)
        return x_cl

    def down_cl(self, x_cl):
        """SnakeBeta + strided conv."""
        n = len(self.block)
        return _act_conv(self.block[n - 2], self.block[n - 1], x_cl)

    def forward_cl(self, x_cl):
        return self.down_cl(self.units_cl(x_cl))

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class DecoderBlock(nn.Module):
    """snake -> transposed conv -> 3 ResidualUnits (vq/module.py:115-141)."""

    def __init__(self, input_dim: int = 16, output_dim: int = 8, stride: int = 1, dilations=(1, 3, 9),
                 causal: bool = False, antialias: bool = False):
        super().__init__()
        if causal:
            tconv_kwargs = {}
        else:
            tconv_kwargs = {"padding": stride // 2 + stride % 2 if stride != 1 else 0,
                            "output_padding": stride % 2 if stride != 1 else 0}
        self.block = nn.Sequential(
            Activation1d(activation=activations.SnakeBeta(input_dim, alpha_logscale=True), antialias=antialias),
            WNConvTranspose1d(input_dim, output_dim, kernel_size=2 * stride if stride != 1 else 1, stride=stride,
                              causal=causal, **tconv_kwargs),
        )
        self.block.extend([ResidualUnit(output_dim, dilation=d, causal=causal, antialias=antialias)
                           for d in dilations])

    def forward_cl(self, x_cl):
        act, up = self.block[0], self.block[1]
        if act.antialias:
            h = up.forward_cl(act.forward_cl(x_cl))
        else:
            h = up.forward_cl(x_cl, act=act.act)
        for ru in list(self.block)[2:]:
            h = ru.forward_cl(h)
        return h

    @torch.no_grad()
    def forward(self, x):
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))


class _LSTMParams(nn.Module):
    """Parameter holder with nn.LSTM's names (weight_ih_l{k}, weight_hh_l{k}, bias_ih_l{k}, bias_hh_l{k})."""

    def __init__(self, input_size, hidden_size, num_layers):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        bound = 1.0 / math.sqrt(hidden_size)
        for l in range(num_layers):
            in_f = input_size if l == 0 else hidden_size
            for name, shape in (("weight_ih", (4 * hidden_size, in_f)), ("weight_hh", (4 * hidden_size, hidden_size)),
                                ("bias_ih", (4 * hidden_size,)), ("bias_hh", (4 * hidden_size,))):
                self.register_parameter(f"{name}_l{l}", nn.Parameter((torch.rand(shape) * 2 - 1) * bound,
                                                                     requires_grad=False))

    def packed(self, layer: int):
        ps = [getattr(self, f"{n}_l{layer}") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        key = tuple(p._version for p in ps) + (ps[0].data_ptr(), ps[0].device)
        cache = self.__dict__.setdefault("_pack_cache", {})
        c = cache.get(layer)
        if c is None or c[0] != key:
            with torch.no_grad():
                w_ih, w_hh, b_ih, b_hh = [p.detach().float() for p in ps]
                H = self.hidden_size
                w_in = w_ih.t().contiguous().unsqueeze(0)                       # [1, in, 4H] (K = 1 conv)
                bias = (b_ih + b_hh).contiguous()
                # packed[cta][j][u*4+g] = w_hh[g*H + cta*4 + u][j]   (bc_lstm_pack_whh layout)
                w_rec = w_hh.view(4, H // 4, 4, H).permute(1, 3, 2, 0).contiguous()
            c = (key, w_in, bias, w_rec)
            cache[layer] = c
        return c[1:]

    def recurrent_image_for(self, layer: int, precision: str):
        """W_hh as the bf16 image of the tensor-core recurrence kernel."""
        w_hh = getattr(self, f"weight_hh_l{layer}")
        cache = self.__dict__.setdefault("_rec_cache", {})
        key = (layer, precision)
        ver = (w_hh._version, w_hh.data_ptr())
        if key not in cache or cache[key][0] != ver:
            with torch.no_grad():
                cache[key] = (ver, ops.pack_lstm_tc_weight(w_hh.detach(), precision))
        return cache[key][1]

    def input_proj_for(self, layer: int, precision: str):
        """Input-projection weight in the form bc_conv1d_fwd wants for ``precision`` (+ effective precision)."""
        w_in, bias, _ = self.packed(layer)
        plan = ops.tc_plan(w_in.shape[1], w_in.shape[2], 1, 1, 1, precision)
        if plan is None:
            return w_in, bias, "fp32"
        cache = self.__dict__.setdefault("_tc_cache", {})
        key = (layer, precision, w_in.data_ptr())
        if key not in cache:
            cache[key] = ops.pack_tc_weight(w_in, plan, precision)
        return cache[key], bias, precision


    def input_proj_stream(self, layer: int, precision: str, n_tile: int, pair: bool = False):
        """Input-projection weight as the streamed-weight image."""
        w_in, _, _ = self.packed(layer)
        cache = self.__dict__.setdefault("_stream_cache", {})
        key = (layer, precision, n_tile, pair, w_in.data_ptr())
        if key not in cache:
            cache[key] = (ops.pack_stream_weight_pair if pair else ops.pack_stream_weight)(w_in, n_tile, precision)
        return cache[key]


class ResLSTM(nn.Module):
    """y = LSTM(x^T) + x^T (vq/module.py:143-167); uni-directional only on the hot path."""

    def __init__(self, dimension: int, num_layers: int = 2, bidirectional: bool = False, skip: bool = True):
        super().__init__()
        if bidirectional:
            raise NotImplementedError("bidirectional ResLSTM is not used by any codec config and is not implemented")
        self.skip = skip
        self.lstm = _LSTMParams(dimension, dimension, num_layers)

    def _input_proj(self, h, l, precision):
        """W_ih h + b_ih + b_hh for every step: one dense K = 1 contraction."""
        H = self.lstm.hidden_size
        nt = _stream_tile(h.shape[2], 4 * H, 1, 1, 1, precision)
        if nt is not None:
            pair = STREAM_PAIR[0] and ops.stream_pair_ok(h.shape[2], 4 * H, 1, 1, 1, precision)
            return ops.conv1d_stream(h, self.lstm.input_proj_stream(l, precision, nt, pair), self.lstm.packed(l)[1], k=1,
                                     c_out=4 * H, t_out=h.shape[1], precision=precision, pair=pair)
        w_in, bias, prec = self.lstm.input_proj_for(l, precision)
        return ops.conv1d(h, w_in, bias, t_out=h.shape[1], precision=prec, geometry=(1, h.shape[2], 4 * H))

    def _wavefront_chunk(self, B, T, precision):
        """Chunk length of the two-layer wave front, or 0 when it does not apply: two layers, a tensor-core plan, and both
        layers' launches co-resident (twice the CTAs of one launch fit the SMs: batches up to 128 in split precision)."""
        if not LSTM_WAVEFRONT[0] or self.lstm.num_layers != 2 or precision == "fp32" or not LSTM_TENSOR_CORE[0]:
            return 0
        H = self.lstm.hidden_size
        ctas = ops.lstm_tc_ctas(B, H, precision)
        if ctas == 0 or 2 * ctas > _cabi_sm_count():
            return 0
        chunk = LSTM_WAVEFRONT_CHUNK[0]
        return chunk if T >= 3 * chunk else 0

    def _forward_wavefront(self, x_cl, precision, chunk):
        """Layer 1 runs on chunk c (input projection of layer 0's output + recurrence, side stream) while layer 0 runs on
        chunk c + 1 (current stream): T + chunk sequential steps instead of 2 T.  Same kernels, same arithmetic and same
        order of operations per element as the layer-by-layer schedule, so the results are bit-identical."""
        B, T, C = x_cl.shape
        H = self.lstm.hidden_size
        dev = x_cl.device
        main = torch.cuda.current_stream(dev)
        side = self.__dict__.setdefault("_side_stream", {}).setdefault(dev, torch.cuda.Stream(device=dev))
        pre0 = self._input_proj(x_cl, 0, precision)                                    # [B, T, 4H]
        img0, img1 = self.lstm.recurrent_image_for(0, precision), self.lstm.recurrent_image_for(1, precision)
        ws0, ws1 = ops.lstm_tc_workspace(B, H, precision, dev), ops.lstm_tc_workspace(B, H, precision, dev)
        rows = (B + 127) // 128 * 128
        c0, c1 = torch.empty((rows, H), device=dev), torch.empty((rows, H), device=dev)
        n_chunks = (T + chunk - 1) // chunk
        h0 = torch.empty((n_chunks, B, chunk, H), device=dev)                           # layer 0's output, chunk-major (dense per chunk)
        y = torch.empty((B, T, H), device=dev)
        done0 = [torch.cuda.Event() for _ in range(n_chunks)]
        for c in range(n_chunks):
            t0, t1 = c * chunk, min(T, (c + 1) * chunk)
            ops.lstm_recurrent_tc_chunk(pre0[:, t0:t1], img0, None, h0[c][:, : t1 - t0], ws0, c0, t0, precision)
            done0[c].record(main)
            with torch.cuda.stream(side):
                side.wait_event(done0[c])
                hc = h0[c][:, : t1 - t0]
                pre1 = self._input_proj(hc if t1 - t0 == chunk else hc.contiguous(), 1, precision)
                ops.lstm_recurrent_tc_chunk(pre1, img1, x_cl[:, t0:t1] if self.skip else None, y[:, t0:t1], ws1, c1, t0, precision)
        fin = torch.cuda.Event()
        fin.record(side)
        main.wait_event(fin)
        for t in (pre0, h0, ws0, ws1, c0, c1, y, x_cl):
            t.record_stream(side)
        return y

    def forward_cl(self, x_cl):
        h = x_cl
        n = self.lstm.num_layers
        precision = get_precision()
        H = self.lstm.hidden_size
        tc_batch = ops.lstm_tc_max_batch(H, precision) if LSTM_TENSOR_CORE[0] else 0
        if tc_batch > 0:
            chunk = self._wavefront_chunk(x_cl.shape[0], x_cl.shape[1], precision)
            if chunk:
                return self._forward_wavefront(x_cl, precision, chunk)
        for l in range(n):
            pre = self._input_proj(h, l, precision)
            skip = x_cl if (self.skip and l == n - 1) else None
            if tc_batch > 0:
                h = ops.lstm_recurrent_tc(pre, self.lstm.recurrent_image_for(l, precision), skip, precision, tc_batch)
            else:
                h = ops.lstm_recurrent(pre, self.lstm.packed(l)[2], skip)
        return h

    @torch.no_grad()
    def forward(self, x):  # [B, F, T]
        return ops.to_channels_first(self.forward_cl(ops.to_channels_last(x)))
