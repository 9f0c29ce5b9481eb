"""Host-side operator layer: thin, shape-checked wrappers over the C ABI.

All activations here are float32 CUDA tensors in CHANNELS-LAST layout ``[B, T, C]``
(contiguous).  Each function allocates its output with torch (device memory is
torch's job), passes raw pointers to the library on the current stream and returns
the output tensor.  No arithmetic happens in Python.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from ._cabi import BC_CONV_SNAKE_IN, BC_CONV_TANH_OUT, PRECISIONS, check, load_library, ptr, require_cuda, stream_ptr


# launch accounting (bench.py's ``gpu_launches``) and optional per-call event timing of the
# dense contractions (bench.py's roofline leg).  PROFILE is None or a list that receives
# (kind, flops, start_event, end_event) per conv call.
STATS = {"launches": 0}
PROFILE = None


def _count(n: int = 1) -> None:
    STATS["launches"] += n


class _Timed:
    def __init__(self, kind, flops, device, kernel="?", nbytes=0.0):
        self.kind, self.flops, self.device, self.kernel, self.nbytes = kind, flops, device, kernel, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream(self.device))
            PROFILE.append((self.kind, self.flops, self.e0, self.e1, self.kernel, self.nbytes))
        return False


def _cl(x: torch.Tensor, name="x") -> torch.Tensor:
    require_cuda(x, name)
    if x.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {x.dtype}")
    if x.dim() != 3:
        raise ValueError(f"{name} must be [B,T,C], got {tuple(x.shape)}")
    return x if x.is_contiguous() else x.contiguous()


def to_channels_last(x_bct: torch.Tensor) -> torch.Tensor:
    """[B,C,T] (any strides) -> contiguous [B,T,C]."""
    require_cuda(x_bct, "x")
    if x_bct.dtype != torch.float32:
        raise TypeError(f"expected float32, got {x_bct.dtype}")
    B, C, T = x_bct.shape
    v = x_bct.permute(0, 2, 1)
    if v.is_contiguous():          # already channels-last storage (or C == 1)
        return v
    x_bct = x_bct.contiguous()
    y = torch.empty((B, T, C), device=x_bct.device, dtype=torch.float32)
    check(load_library().bc_transpose_bct_to_btc(ptr(x_bct), ptr(y), B, C, T, stream_ptr(x_bct.device)),
          "bc_transpose_bct_to_btc")
    _count()
    return y


def to_channels_first(x_btc: torch.Tensor) -> torch.Tensor:
    """contiguous [B,T,C] -> contiguous [B,C,T]."""
    x_btc = _cl(x_btc)
    B, T, C = x_btc.shape
    if C == 1 or T == 1:
        return x_btc.reshape(B, C, T)
    y = torch.empty((B, C, T), device=x_btc.device, dtype=torch.float32)
    check(load_library().bc_transpose_btc_to_bct(ptr(x_btc), ptr(y), B, T, C, stream_ptr(x_btc.device)),
          "bc_transpose_btc_to_bct")
    _count()
    return y


def snake(x: torch.Tensor, a: torch.Tensor, ib: torch.Tensor, antialias: bool = False,
          fir: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SnakeBeta / anti-aliased Activation1d on a channels-last tensor."""
    x = _cl(x)
    B, T, C = x.shape
    y = torch.empty_like(x)
    check(load_library().bc_snake_fwd(ptr(x), ptr(y), ptr(a), ptr(ib), ptr(fir) if antialias else None,
                                      B, T, C, int(bool(antialias)), stream_ptr(x.device)), "bc_snake_fwd")
    _count()
    return y


def conv1d(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, stride: int = 1, dilation: int = 1,
           pad_left: int = 0, t_out: int, snake_a: Optional[torch.Tensor] = None,
           snake_ib: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None, tanh: bool = False,
           precision: str = "fp32", out: Optional[torch.Tensor] = None, geometry=None) -> torch.Tensor:
    """Dense conv.  ``w`` is the fp32 ``[K, C_in, C_out]`` array for precision 'fp32', or the bf16
    tensor-core image of ``pack_tc_weight`` (then ``geometry=(K, C_in, C_out)`` or the image's own
    shape gives the sizes); see bc_conv1d_fwd / bc_tc_plan."""
    x = _cl(x)
    B, T_in, C_in = x.shape
    if geometry is not None:
        K, wc_in, C_out = geometry
    elif w.dim() == 3:
        K, wc_in, C_out = w.shape
    else:  # [nt, nchunks, split, K, gpc, 2, n_tile, 8]
        K, wc_in, C_out = w.shape[3], w.shape[1] * w.shape[4] * 16, w.shape[0] * w.shape[6]
    if (precision == "fp32") != (w.dtype == torch.float32):
        raise TypeError(f"conv1d: weight dtype {w.dtype} does not match precision {precision!r}")
    if wc_in != C_in:
        raise ValueError(f"conv1d: input has {C_in} channels, weight expects {wc_in}")
    if t_out <= 0:
        raise ValueError(f"conv1d: input of {T_in} steps is too short for this layer (T_out={t_out})")
    y = out if out is not None else torch.empty((B, t_out, C_out), device=x.device, dtype=torch.float32)
    flags = (BC_CONV_SNAKE_IN if snake_a is not None else 0) | (BC_CONV_TANH_OUT if tanh else 0)
    if res is not None:
        res = _cl(res, "res")
        if tuple(res.shape) != (B, t_out, C_out):
            raise ValueError(f"conv1d: residual shape {tuple(res.shape)} != output {(B, t_out, C_out)}")
    with _Timed(("conv1d", C_in, C_out, K, stride, dilation, t_out, B, precision), 2.0 * B * t_out * C_out * C_in * K, x.device,
                "conv1d_f32_kernel" if precision == "fp32" else "conv1d_tc_kernel",
                4.0 * B * (T_in * C_in + t_out * C_out * (2 if res is not None else 1))):
        check(load_library().bc_conv1d_fwd(ptr(x), ptr(w), ptr(bias), ptr(snake_a), ptr(snake_ib), ptr(res), ptr(y),
                                           B, T_in, C_in, t_out, C_out, K, stride, dilation, pad_left,
                                           t_out, 1, 0, flags, PRECISIONS[precision], stream_ptr(x.device)),
              "bc_conv1d_fwd")
    _count()
    return y


def resunit(x: torch.Tensor, w7: torch.Tensor, b7, sa1, sib1, w1: torch.Tensor, b1, sa2, sib2, *, k: int,
            dilation: int, pad_left: int, precision: str) -> torch.Tensor:
    """Fused ResidualUnit (tensor-core modes): y = x + W1 snake2(W7 * snake1(x) + b7) + b1."""
    x = _cl(x)
    B, T, C = x.shape
    y = torch.empty_like(x)
    flops = 2.0 * B * T * C * C * (k + 1)
    plan = resunit_plan(C, k, dilation, precision)
    with _Timed(("resunit", C, C, k, 1, dilation, T, B, precision), flops, x.device,
                {1: "ru_persist_kernel", 2: "ru_group_kernel", 3: "ru_pair_kernel", 4: "ru_pair_kernel"}.get(plan[1] if plan else 0, "conv1d_tc_kernel"),
                8.0 * B * T * C):
        check(load_library().bc_resunit_fwd(ptr(x), ptr(w7), ptr(b7), ptr(sa1), ptr(sib1), ptr(w1), ptr(b1), ptr(sa2),
                                            ptr(sib2), ptr(y), B, T, C, k, dilation, pad_left, PRECISIONS[precision],
                                            stream_ptr(x.device)), "bc_resunit_fwd")
    _count()
    return y


def conv_transpose1d(x: torch.Tensor, w_phases: torch.Tensor, bias: Optional[torch.Tensor], *, stride: int,
                     padding: int, snake_a: Optional[torch.Tensor] = None, snake_ib: Optional[torch.Tensor] = None,
                     precision: str = "fp32", c_out: Optional[int] = None) -> torch.Tensor:
    """Transposed conv (k = 2*stride) from phase-packed weights ``[stride, 2, C_in, C_out]`` (fp32) or one
    tensor-core image per phase (``[stride, <pack_tc_weight image>]``)."""
    x = _cl(x)
    B, T_in, C_in = x.shape
    if (precision == "fp32") != (w_phases.dtype == torch.float32):
        raise TypeError(f"conv_transpose1d: weight dtype {w_phases.dtype} does not match precision {precision!r}")
    if precision == "fp32":
        s, two, wc_in, C_out = w_phases.shape
    else:
        img = w_phases.shape[1:]
        s, two, wc_in, C_out = w_phases.shape[0], img[3], img[1] * img[4] * 16, img[0] * img[6]
    if c_out is not None and c_out != C_out:
        raise ValueError("conv_transpose1d: weight does not match c_out")
    if s != stride or two != 2 or wc_in != C_in:
        raise ValueError(f"conv_transpose1d: weight {tuple(w_phases.shape)} does not match stride={stride}, C_in={C_in}")
    y = torch.empty((B, T_in * stride, C_out), device=x.device, dtype=torch.float32)
    flags = BC_CONV_SNAKE_IN if snake_a is not None else 0
    with _Timed(("convtr1d", C_in, C_out, 2 * stride, stride, 1, T_in * stride, B, precision),
                2.0 * B * T_in * stride * C_out * C_in * 2, x.device,
                "conv1d_f32_kernel" if precision == "fp32" else "conv1d_tc_kernel", 4.0 * B * T_in * (C_in + stride * C_out)):
        check(load_library().bc_convtr1d_fwd(ptr(x), ptr(w_phases), ptr(bias), ptr(snake_a), ptr(snake_ib), ptr(y),
                                             B, T_in, C_in, C_out, stride, padding, flags, PRECISIONS[precision],
                                             stream_ptr(x.device)), "bc_convtr1d_fwd")
    _count(stride)
    return y


def conv_transpose1d_stream(x: torch.Tensor, w_img: torch.Tensor, bias_tiled: Optional[torch.Tensor], *, stride: int,
                            padding: int, c_out: int, three_tap: bool, snake_a: Optional[torch.Tensor] = None,
                            snake_ib: Optional[torch.Tensor] = None, precision: str, pair: bool = False) -> torch.Tensor:
    """Transposed conv (k = 2*stride) as ONE launch of the persistent streamed-weight kernel: all output phases are the
    channel blocks of a conv with stride*C_out outputs (``pack_convtr_stream_weight``)."""
    x = _cl(x)
    B, T_in, C_in = x.shape
    y = torch.empty((B, T_in, stride * c_out), device=x.device, dtype=torch.float32)
    flags = BC_CONV_SNAKE_IN if snake_a is not None else 0
    with _Timed(("convtr1d", C_in, c_out, 2 * stride, stride, 1, T_in * stride, B, precision),
                2.0 * B * T_in * stride * c_out * C_in * 2, x.device, "conv_stream_kernel", 4.0 * B * T_in * (C_in + stride * c_out)):
        lib = load_library()
        if three_tap:
            fn = lib.bc_conv1d_stream_pair_fwd if pair else lib.bc_conv1d_stream_fwd
            check(fn(ptr(x), ptr(w_img), ptr(bias_tiled), ptr(snake_a), ptr(snake_ib), None,
                     ptr(y), B, T_in, C_in, T_in, stride * c_out, 3, 1, 1, 1, flags,
                     PRECISIONS[precision], stream_ptr(x.device)), "bc_conv1d_stream_fwd")
        else:
            fn = lib.bc_convtr1d_stream_pair_fwd if pair else lib.bc_convtr1d_stream_fwd
            check(fn(ptr(x), ptr(w_img), ptr(bias_tiled), ptr(snake_a), ptr(snake_ib), ptr(y),
                     B, T_in, C_in, c_out, stride, padding, flags, PRECISIONS[precision],
                     stream_ptr(x.device)), "bc_convtr1d_stream_fwd")
    _count()
    return y.view(B, T_in * stride, c_out)


def pack_convtr_stream_weight(w_phases: torch.Tensor, stride: int, padding: int, n_tile: int, precision: str,
                              pair: bool = False):
    """fp32 phase filters [stride, 2, C_in, C_out] -> (streamed-weight image, three_tap).  Two taps when every n-tile
    lies inside one phase (the kernel shifts the taps of the q = 1 phases by one row); otherwise three taps on rows
    m-1, m, m+1 with zeros where a phase does not reach (q = 0: taps 0, 1; q = 1: taps 1, 2)."""
    s, two, c_in, c_out = w_phases.shape
    three_tap = c_out % n_tile != 0
    if not three_tap:
        w = w_phases.permute(1, 2, 0, 3).reshape(2, c_in, s * c_out)
    else:
        w = torch.zeros((3, c_in, s * c_out), dtype=w_phases.dtype, device=w_phases.device)
        for ph in range(s):
            q = (ph + padding) // s
            w[q:q + 2, :, ph * c_out:(ph + 1) * c_out] = w_phases[ph]
    return (pack_stream_weight_pair if pair else pack_stream_weight)(w.contiguous(), n_tile, precision), three_tap


_TC_PLANS = {}


def tc_plan(c_in: int, c_out: int, k: int, stride: int, dilation: int, precision: str):
    """(n_tile, gpc, nchunks) of the tensor-core tiling, or None when this geometry stays on the fp32 kernel."""
    if precision == "fp32":
        return None
    key = (c_in, c_out, k, stride, dilation, precision)
    if key not in _TC_PLANS:
        import ctypes
        nt, g, nc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = load_library().bc_tc_plan(c_in, c_out, k, stride, dilation, PRECISIONS[precision],
                                       ctypes.byref(nt), ctypes.byref(g), ctypes.byref(nc))
        _TC_PLANS[key] = (nt.value, g.value, nc.value) if rc == 0 else None
    return _TC_PLANS[key]


def resunit_plan(c: int, k: int, dilation: int, precision: str):
    """((n_tile, gpc, nchunks) of the W7 image, kernel: 0 per-tile / 1 ru_persist / 2 ru_group) for the fused
    ResidualUnit kernel, or None."""
    if precision == "fp32":
        return None
    key = ("ru", c, k, dilation, precision)
    if key not in _TC_PLANS:
        import ctypes
        nt, g, nc, pers = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = load_library().bc_resunit_plan(c, k, dilation, PRECISIONS[precision], ctypes.byref(nt), ctypes.byref(g),
                                            ctypes.byref(nc), ctypes.byref(pers))
        _TC_PLANS[key] = ((nt.value, g.value, nc.value), int(pers.value)) if rc == 0 else None
    return _TC_PLANS[key]


def pack_tc_weight(w_kio: torch.Tensor, plan, precision: str) -> torch.Tensor:
    """fp32 [K, C_in, C_out] -> the bf16 smem image bc_conv1d_fwd consumes in tensor-core modes
    ([C_out/n_tile][nchunks][split][K][gpc][2][n_tile][8]; see bc_tc_plan)."""
    n_tile, gpc, nchunks = plan
    K, c_in, c_out = w_kio.shape
    w = w_kio.float()
    hi = w.to(torch.bfloat16)

    def image(t):
        return t.reshape(K, nchunks, gpc, 2, 8, c_out // n_tile, n_tile).permute(5, 1, 0, 2, 3, 6, 4)

    parts = [image(hi)]
    if precision == "bf16x3":
        parts.append(image((w - hi.float()).to(torch.bfloat16)))
    return torch.stack(parts, dim=2).contiguous()


def pack_pair_weights(w_kio: torch.Tensor, stacked: bool) -> torch.Tensor:
    """fp32 [K, C_in, C_out] -> the per-rank bf16 images of the CTA-pair ResidualUnit kernel (bc_resunit_plan kinds 3 / 4):
    [2 ranks][...], see include/bigcodec_b200.h."""
    K, c_in, c_out = w_kio.shape
    w = w_kio.float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    half = c_out // 2

    def block(m, r0, r1):   # rows [r0, r1) of m -> [K][c_in/16][2][rows][8], flattened
        t = m[:, :, r0:r1]
        return t.reshape(K, c_in // 16, 2, 8, r1 - r0).permute(0, 1, 2, 4, 3).reshape(-1)

    ranks = []
    for r in range(2):
        if stacked:
            ranks.append(torch.cat([block(hi if r == 0 else lo, 0, c_out), block(hi, r * half, (r + 1) * half)]))
        else:
            ranks.append(torch.cat([block(hi, r * half, (r + 1) * half), block(lo, r * half, (r + 1) * half)]))
    return torch.stack(ranks).contiguous()


def stream_plan(c_in: int, c_out: int, k: int, stride: int, dilation: int, precision: str, fused: bool = False):
    """n_tile of the streamed-weight persistent kernel (bc_stream_plan), or None when the geometry has no plan."""
    if precision == "fp32":
        return None
    key = ("stream", c_in, c_out, k, stride, dilation, precision, bool(fused))
    if key not in _TC_PLANS:
        import ctypes
        nt = ctypes.c_int()
        rc = load_library().bc_stream_plan(c_in, c_out, k, stride, dilation, PRECISIONS[precision], int(bool(fused)),
                                           ctypes.byref(nt))
        _TC_PLANS[key] = nt.value if rc == 0 else None
    return _TC_PLANS[key]


def pack_stream_weight(w_kio: torch.Tensor, n_tile: int, precision: str) -> torch.Tensor:
    """fp32 [K, C_in, C_out] -> bf16 image [C_out/n_tile][C_in/16][K][split][2][n_tile][8] of the streamed-weight
    kernel: one (16-channel group, tap) block after the other in the order the kernel consumes them."""
    K, c_in, c_out = w_kio.shape
    w = w_kio.float()
    hi = w.to(torch.bfloat16)

    def image(t):   # (k, g, h, e, nt, n) -> (nt, g, k, h, n, e)
        return t.reshape(K, c_in // 16, 2, 8, c_out // n_tile, n_tile).permute(4, 1, 0, 2, 5, 3)

    parts = [image(hi)]
    if precision == "bf16x3":
        parts.append(image((w - hi.float()).to(torch.bfloat16)))
    return torch.stack(parts, dim=3).contiguous()


def stream_pair_ok(c_in: int, c_out: int, k: int, stride: int, dilation: int, precision: str, fused: bool = False) -> bool:
    """True when the CTA-pair form (tcgen05 cta_group::2) of the streamed-weight kernel takes this geometry."""
    if precision == "fp32":
        return False
    key = ("pair", c_in, c_out, k, stride, dilation, precision, bool(fused))
    if key not in _TC_PLANS:
        _TC_PLANS[key] = bool(load_library().bc_stream_pair_ok(c_in, c_out, k, stride, dilation, PRECISIONS[precision], int(bool(fused))))
    return _TC_PLANS[key]


def pack_stream_weight_pair(w_kio: torch.Tensor, n_tile: int, precision: str) -> torch.Tensor:
    """fp32 [K, C_in, C_out] -> bf16 PAIR image [C_out/n_tile][2][C_in/16][K][split][2][n_tile/2][8] of the CTA-pair form:
    rank r of the pair holds rows [r*n_tile/2, (r+1)*n_tile/2) of every k-plane of every (16-channel group, tap) block."""
    K, c_in, c_out = w_kio.shape
    w = w_kio.float()
    hi = w.to(torch.bfloat16)
    h2 = n_tile // 2

    def image(t):   # (k, g, h, e, nt, r, n) -> (nt, r, g, k, h, n, e)
        return t.reshape(K, c_in // 16, 2, 8, c_out // n_tile, 2, h2).permute(4, 5, 1, 0, 2, 6, 3)

    parts = [image(hi)]
    if precision == "bf16x3":
        parts.append(image((w - hi.float()).to(torch.bfloat16)))
    return torch.stack(parts, dim=4).contiguous()


def conv1d_stream(x: torch.Tensor, w_img: torch.Tensor, bias: Optional[torch.Tensor], *, k: int, c_out: int,
                  stride: int = 1, dilation: int = 1, pad_left: int = 0, t_out: int,
                  snake_a: Optional[torch.Tensor] = None, snake_ib: Optional[torch.Tensor] = None,
                  res: Optional[torch.Tensor] = None, tanh: bool = False, precision: str, pair: bool = False) -> torch.Tensor:
    """Dense conv on the persistent streamed-weight kernel (``w_img`` from pack_stream_weight, or, with ``pair``, from
    pack_stream_weight_pair: the CTA-pair form)."""
    x = _cl(x)
    B, T_in, C_in = x.shape
    if t_out <= 0:
        raise ValueError(f"conv1d: input of {T_in} steps is too short for this layer (T_out={t_out})")
    y = torch.empty((B, t_out, c_out), device=x.device, dtype=torch.float32)
    flags = (BC_CONV_SNAKE_IN if snake_a is not None else 0) | (BC_CONV_TANH_OUT if tanh else 0)
    if res is not None:
        res = _cl(res, "res")
        if tuple(res.shape) != (B, t_out, c_out):
            raise ValueError(f"conv1d: residual shape {tuple(res.shape)} != output {(B, t_out, c_out)}")
    with _Timed(("conv1d", C_in, c_out, k, stride, dilation, t_out, B, precision), 2.0 * B * t_out * c_out * C_in * k, x.device,
                "conv_stream_kernel", 4.0 * B * (T_in * C_in + t_out * c_out * (2 if res is not None else 1))):
        fn = load_library().bc_conv1d_stream_pair_fwd if pair else load_library().bc_conv1d_stream_fwd
        check(fn(ptr(x), ptr(w_img), ptr(bias), ptr(snake_a), ptr(snake_ib), ptr(res),
                 ptr(y), B, T_in, C_in, t_out, c_out, k, stride, dilation, pad_left,
                 flags, PRECISIONS[precision], stream_ptr(x.device)),
              "bc_conv1d_stream_fwd")
    _count()
    return y


def resunit_stream(x: torch.Tensor, w7: torch.Tensor, b7, sa1, sib1, w1: torch.Tensor, b1, sa2, sib2, *, k: int,
                   dilation: int, pad_left: int, precision: str, out: Optional[torch.Tensor] = None,
                   pair: bool = False) -> torch.Tensor:
    """Fused ResidualUnit on the persistent streamed-weight kernel (wide layers).  ``out``: write the result
    into this contiguous [B,T,C] tensor (e.g. a slice of a larger batch buffer) instead of a fresh one."""
    x = _cl(x)
    B, T, C = x.shape
    if out is not None:
        if out.shape != x.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
            raise ValueError("resunit_stream: `out` must be a contiguous float32 tensor of the input's shape on its device")
        y = out
    else:
        y = torch.empty_like(x)
    flops = 2.0 * B * T * C * C * (k + 1)
    with _Timed(("resunit", C, C, k, 1, dilation, T, B, precision), flops, x.device, "conv_stream_kernel", 8.0 * B * T * C):
        fn = load_library().bc_resunit_stream_pair_fwd if pair else load_library().bc_resunit_stream_fwd
        check(fn(ptr(x), ptr(w7), ptr(b7), ptr(sa1), ptr(sib1), ptr(w1), ptr(b1),
                 ptr(sa2), ptr(sib2), ptr(y), B, T, C, k, dilation, pad_left,
                 PRECISIONS[precision], stream_ptr(x.device)),
              "bc_resunit_stream_fwd")
    _count()
    return y


LSTM_MAX_BATCH = 256


def lstm_recurrent(pre: torch.Tensor, w_hh_packed: torch.Tensor, skip: Optional[torch.Tensor]) -> torch.Tensor:
    """Recurrent part of one LSTM layer; ``pre`` = [B,T,4H] input projection (+ both biases)."""
    pre = _cl(pre, "pre")
    B, T, H4 = pre.shape
    H = H4 // 4
    lib = load_library()
    y = torch.empty((B, T, H), device=pre.device, dtype=torch.float32)
    if skip is not None:
        skip = _cl(skip, "skip")
    for b0 in range(0, B, LSTM_MAX_BATCH):
        b1 = min(B, b0 + LSTM_MAX_BATCH)
        ws = torch.empty(lib.bc_lstm_workspace_bytes(b1 - b0, H), device=pre.device, dtype=torch.uint8)
        with _Timed(("lstm", H, H, 0, 0, 0, T, b1 - b0, "fp32"), 2.0 * (b1 - b0) * T * 4 * H * H, pre.device, "lstm_rec_kernel"):
            check(lib.bc_lstm_recurrent_fwd(ptr(pre[b0:b1]), ptr(w_hh_packed),
                                            ptr(skip[b0:b1]) if skip is not None else None,
                                            ptr(y[b0:b1]), ptr(ws), b1 - b0, T, H, stream_ptr(pre.device)),
                  "bc_lstm_recurrent_fwd")
        _count()
    return y


def lstm_tc_max_batch(H: int, precision: str) -> int:
    """Largest batch one tensor-core LSTM launch takes (0: no tensor-core plan for this H / precision)."""
    if precision == "fp32":
        return 0
    return int(load_library().bc_lstm_tc_max_batch(H, PRECISIONS[precision]))


def pack_lstm_tc_weight(w_hh: torch.Tensor, precision: str) -> torch.Tensor:
    """W_hh [4H, H] fp32 -> bf16 image [4H/NS][H/16][2][split*NS][8] of bc_lstm_tc_recurrent_fwd: per 16-channel group
    and k-plane the NS rows of the hi slice followed (split precision) by the NS rows of the lo slice, so that
    [w_hi | w_lo] is one tcgen05 B operand of 2*NS rows."""
    NS = int(load_library().bc_lstm_tc_slice_cols(PRECISIONS[precision]))
    H = w_hh.shape[1]
    U = NS // 4
    w = w_hh.float().view(4, H // U, U, H).permute(1, 0, 2, 3).reshape(4 * H // NS, NS, H)
    hi = w.to(torch.bfloat16)

    def image(t):   # [S, NS, H] -> [S, H/16, 2, NS, 8]
        return t.reshape(-1, NS, H // 16, 2, 8).permute(0, 2, 3, 1, 4)

    parts = [image(hi)]
    if precision == "bf16x3":
        parts.append(image((w - hi.float()).to(torch.bfloat16)))
    return torch.cat(parts, dim=3).contiguous()


def lstm_recurrent_tc(pre: torch.Tensor, w_image: torch.Tensor, skip: Optional[torch.Tensor], precision: str,
                      max_batch: int) -> torch.Tensor:
    """Tensor-core recurrence; ``pre`` = [B,T,4H] input projection (+ both biases)."""
    pre = _cl(pre, "pre")
    B, T, H4 = pre.shape
    H = H4 // 4
    lib = load_library()
    y = torch.empty((B, T, H), device=pre.device, dtype=torch.float32)
    if skip is not None:
        skip = _cl(skip, "skip")
    for b0 in range(0, B, max_batch):
        b1 = min(B, b0 + max_batch)
        ws = torch.empty(lib.bc_lstm_tc_workspace_bytes(b1 - b0, H, PRECISIONS[precision]), device=pre.device,
                         dtype=torch.uint8)
        with _Timed(("lstm", H, H, 0, 0, 0, T, b1 - b0, precision), 2.0 * (b1 - b0) * T * 4 * H * H, pre.device, "lstm_tc_kernel"):
            check(lib.bc_lstm_tc_recurrent_fwd(ptr(pre[b0:b1]), ptr(w_image),
                                               ptr(skip[b0:b1]) if skip is not None else None, ptr(y[b0:b1]), ptr(ws),
                                               b1 - b0, T, H, PRECISIONS[precision], stream_ptr(pre.device)),
                  "bc_lstm_tc_recurrent_fwd")
        _count()
    return y


def lstm_tc_ctas(B: int, H: int, precision: str) -> int:
    """CTAs one tensor-core LSTM launch of batch ``B`` occupies (0: no plan)."""
    if precision == "fp32":
        return 0
    return int(load_library().bc_lstm_tc_ctas(B, H, PRECISIONS[precision]))


def lstm_tc_workspace(B: int, H: int, precision: str, device) -> torch.Tensor:
    return torch.empty(load_library().bc_lstm_tc_workspace_bytes(B, H, PRECISIONS[precision]), device=device, dtype=torch.uint8)


def lstm_recurrent_tc_chunk(pre: torch.Tensor, w_image: torch.Tensor, skip: Optional[torch.Tensor], y: torch.Tensor,
                            ws: torch.Tensor, c_state: torch.Tensor, t_base: int, precision: str) -> None:
    """One chunk of a chunked sequence (bc_lstm_tc_recurrent_chunk_fwd).  ``pre`` [B,Tc,4H], ``y`` / ``skip`` [B,Tc,H] may be
    time slices of larger tensors (batch stride = the full length); the innermost two dimensions must be dense."""
    B, Tc, H4 = pre.shape
    H = H4 // 4
    for t, inner in ((pre, H4), (y, H)) + (((skip, H),) if skip is not None else ()):
        if t.stride(2) != 1 or t.stride(1) != inner or (B > 1 and t.stride(0) % inner != 0):
            raise ValueError("lstm_recurrent_tc_chunk: tensors must be time slices of dense [B,T,C] tensors")
    if skip is not None and (skip.stride(0) != y.stride(0) and B > 1):
        raise ValueError("lstm_recurrent_tc_chunk: skip and y must share their batch stride")
    pre_rows = pre.stride(0) // H4 if B > 1 else Tc
    y_rows = y.stride(0) // H if B > 1 else Tc
    with _Timed(("lstm", H, H, 0, 0, 0, Tc, B, precision), 2.0 * B * Tc * 4 * H * H, pre.device, "lstm_tc_kernel"):
        check(load_library().bc_lstm_tc_recurrent_chunk_fwd(ptr(pre), ptr(w_image), ptr(skip), ptr(y), ptr(ws), ptr(c_state),
                                                            B, Tc, pre_rows, y_rows, int(t_base), H, PRECISIONS[precision],
                                                            stream_ptr(pre.device)), "bc_lstm_tc_recurrent_chunk_fwd")
    _count()


def vq_encode(z: torch.Tensor, w_in: Optional[torch.Tensor], b_in: Optional[torch.Tensor], cb_norm: torch.Tensor,
              want_margin: bool = False, want_ze: bool = False):
    """z [N,C] -> (idx int32 [N], margin [N] | None, z_e [N,D] | None)."""
    require_cuda(z, "z")
    z = z if z.is_contiguous() else z.contiguous()
    N, C = z.shape
    Kc, D = cb_norm.shape
    idx = torch.empty((N,), device=z.device, dtype=torch.int32)
    margin = torch.empty((N,), device=z.device, dtype=torch.float32) if want_margin else None
    z_e = torch.empty((N, D), device=z.device, dtype=torch.float32) if want_ze else None
    check(load_library().bc_vq_encode(ptr(z), ptr(w_in), ptr(b_in), ptr(cb_norm), ptr(idx), ptr(margin), ptr(z_e),
                                      N, C, D, Kc, stream_ptr(z.device)), "bc_vq_encode")
    _count()
    return idx, margin, z_e


def vq_dequant(idx: torch.Tensor, cb: torch.Tensor, w_out: Optional[torch.Tensor], b_out: Optional[torch.Tensor],
               C: int, *, z_q: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
               check_range: bool = True) -> torch.Tensor:
    """idx int32 [N] -> z_q [N,C] (accumulating into ``z_q`` when given; ``residual -= q``)."""
    require_cuda(idx, "idx")
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    idx = idx.contiguous()
    N = idx.numel()
    Kc, D = cb.shape
    accumulate = z_q is not None
    if z_q is None:
        z_q = torch.empty((N, C), device=idx.device, dtype=torch.float32)
    bad = torch.zeros((1,), device=idx.device, dtype=torch.int32) if check_range else None
    check(load_library().bc_vq_dequant(ptr(idx), ptr(cb), ptr(w_out), ptr(b_out), ptr(z_q), ptr(residual), ptr(bad),
                                       N, C, D, Kc, int(accumulate), stream_ptr(idx.device)), "bc_vq_dequant")
    _count()
    if check_range and int(bad.item()) != 0:
        raise IndexError(f"vq_dequant: {int(bad.item())} indices outside [0, {Kc})")
    return z_q


def fsq_encode(z: torch.Tensor, w_in: Optional[torch.Tensor], b_in: Optional[torch.Tensor], params: torch.Tensor,
               want_codes: bool = False, want_boundary: bool = False):
    """z [N,C] -> (idx int32 [N], codes [N,D] | None, boundary [N] | None); ``params`` = float32 [5,D] (bc_fsq_encode)."""
    require_cuda(z, "z")
    z = z if z.is_contiguous() else z.contiguous()
    N, C = z.shape
    D = params.shape[1]
    idx = torch.empty((N,), device=z.device, dtype=torch.int32)
    codes = torch.empty((N, D), device=z.device, dtype=torch.float32) if want_codes else None
    boundary = torch.empty((N,), device=z.device, dtype=torch.float32) if want_boundary else None
    check(load_library().bc_fsq_encode(ptr(z), ptr(w_in), ptr(b_in), ptr(params), ptr(idx), ptr(codes), ptr(boundary),
                                       N, C, D, stream_ptr(z.device)), "bc_fsq_encode")
    _count()
    return idx, codes, boundary


def indices_to_int16(idx: torch.Tensor) -> torch.Tensor:
    """int32 [n_q, N] -> int16 [N, n_q] (extract_indices.py:520-532 layout)."""
    require_cuda(idx, "idx")
    idx = idx.to(torch.int32).contiguous()
    n_q, N = idx.shape
    out = torch.empty((N, n_q), device=idx.device, dtype=torch.int16)
    check(load_library().bc_indices_to_int16(ptr(idx), ptr(out), n_q, N, stream_ptr(idx.device)), "bc_indices_to_int16")
    _count()
    return out
