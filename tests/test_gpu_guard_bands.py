"""GPU: output guard bands.  `compute-sanitizer` is closed on this pool (gpurun answers: "compute-sanitizer is closed on this
pool and stays closed"), so out-of-bounds WRITES of the kernels whose staging blocks alias live operand buffers or whose tiles
can over-run an item (ragged last tiles, phantom tiles of the CTA-pair kernels, partial chunks of the LSTM wave front) are
checked the direct way: every output the C ABI writes is a window inside a larger buffer pre-filled with a sentinel, and the
bytes before and after the window must still hold the sentinel afterwards -- while the window itself must equal the result of
the same call into an unguarded buffer."""
import ctypes

import pytest
import torch

from audiotokenization_b200 import _cabi, ops, synth
from audiotokenization_b200._cabi import PRECISIONS, check, ptr, stream_ptr
from audiotokenization_b200.vq import module as M
from audiotokenization_b200.vq import FSQ, FactorizedVectorQuantize

pytestmark = pytest.mark.gpu
DEV = "cuda"
GUARD = 4096          # elements on each side
SENT = -12345.0


class Guarded:
    def __init__(self, shape, dtype=torch.float32):
        self.n = int(torch.Size(shape).numel())
        self.full = torch.full((self.n + 2 * GUARD,), SENT, device=DEV, dtype=dtype) if dtype.is_floating_point else \
            torch.full((self.n + 2 * GUARD,), -12345, device=DEV, dtype=dtype)
        self.view = self.full[GUARD:GUARD + self.n].view(shape)
        self.sent = SENT if dtype.is_floating_point else -12345

    def intact(self):
        torch.cuda.synchronize()
        return bool((self.full[:GUARD] == self.sent).all()) and bool((self.full[GUARD + self.n:] == self.sent).all())


def gen(seed):
    return torch.Generator().manual_seed(seed)


def _ru_params(ru, precision):
    conv7, conv1, plan7, plan1, kind = ru._fused_plan(precision)
    if kind in (3, 4):
        w7 = ops.pack_pair_weights(conv7.packed()[0], stacked=(kind == 4))
        w1 = ops.pack_pair_weights(conv1.packed()[0], stacked=False)
    else:
        w7 = ops.pack_tc_weight(conv7.packed()[0], plan7, precision)
        w1 = ops.pack_tc_weight(conv1.packed()[0], plan1, precision)
    sa1, sib1 = ru.block[0].act.device_params()
    sa2, sib2 = ru.block[2].act.device_params()
    return conv7, conv1, w7, w1, sa1, sib1, sa2, sib2


@pytest.mark.parametrize("C,dil,B,T", [(64, 9, 3, 1000), (64, 1, 1, 129), (32, 3, 2, 777), (64, 3, 5, 128 * 3 + 1)])
def test_fused_residual_unit_kernels_stay_inside_their_output(C, dil, B, T):
    """ru_pair (C = 64: odd tile counts -> phantom tile of rank 1) and ru_group (C = 32), ragged last tiles."""
    ru = M.ResidualUnit(C, dilation=dil).to(DEV)
    x = torch.randn(B, T, C, generator=gen(C + dil + T)).to(DEV)
    conv7, conv1, w7, w1, sa1, sib1, sa2, sib2 = _ru_params(ru, "bf16x3")
    lib = _cabi.load_library()
    outs = []
    for g in (Guarded((B, T, C)), None):
        y = g.view if g is not None else torch.empty((B, T, C), device=DEV)
        check(lib.bc_resunit_fwd(ptr(x), ptr(w7), ptr(conv7.packed()[1]), ptr(sa1), ptr(sib1), ptr(w1), ptr(conv1.packed()[1]),
                                 ptr(sa2), ptr(sib2), ptr(y), B, T, C, 7, dil, conv7.left_pad, PRECISIONS["bf16x3"], stream_ptr(x.device)),
              "bc_resunit_fwd")
        outs.append(y)
        if g is not None:
            assert g.intact(), "write outside the output window"
    assert torch.equal(outs[0], outs[1]) and not bool((outs[0] == SENT).any())


@pytest.mark.parametrize("C,dil,B,T", [(128, 9, 3, 700), (256, 3, 1, 129), (128, 1, 5, 128 * 2 + 5)])
def test_streamed_fused_unit_pair_form_stays_inside_its_output(C, dil, B, T):
    """conv_stream pair form: odd tile counts per n-tile (phantom tile), ragged last tiles."""
    ru = M.ResidualUnit(C, dilation=dil).to(DEV)
    x = torch.randn(B, T, C, generator=gen(C + dil + T)).to(DEV)
    M.set_precision("bf16x3")
    try:
        g = Guarded((B, T, C))
        y = ru.forward_cl(x, out=g.view)
        assert g.intact(), "write outside the output window"
        ref = ru.forward_cl(x)
    finally:
        M.set_precision("fp32")
    assert torch.equal(y, ref) and not bool((y == SENT).any())


@pytest.mark.parametrize("cin,cout,stride,B,T", [(256, 128, 5, 2, 333), (512, 256, 5, 1, 129), (64, 32, 2, 3, 500)])
def test_single_launch_transposed_conv_stays_inside_its_output(cin, cout, stride, B, T):
    m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, padding=stride // 2 + stride % 2, output_padding=stride % 2).to(DEV)
    x = torch.randn(B, T, cin, generator=gen(cin + T)).to(DEV)
    lib = _cabi.load_library()
    M.set_precision("bf16x3")
    try:
        img, bias, three, pair = m.stream_image("bf16x3")
        ref = m.forward_cl(x)
    finally:
        M.set_precision("fp32")
    g = Guarded((B, T, stride * cout))
    if three:
        fn = lib.bc_conv1d_stream_pair_fwd if pair else lib.bc_conv1d_stream_fwd
        check(fn(ptr(x), ptr(img), ptr(bias), None, None, None, ptr(g.view), B, T, cin, T, stride * cout, 3, 1, 1, 1, 0,
                 PRECISIONS["bf16x3"], stream_ptr(x.device)), "conv1d_stream")
    else:
        fn = lib.bc_convtr1d_stream_pair_fwd if pair else lib.bc_convtr1d_stream_fwd
        check(fn(ptr(x), ptr(img), ptr(bias), None, None, ptr(g.view), B, T, cin, cout, stride, m.padding, 0,
                 PRECISIONS["bf16x3"], stream_ptr(x.device)), "convtr1d_stream")
    assert g.intact(), "write outside the output window"
    assert torch.equal(g.view.view(B, T * stride, cout), ref)


@pytest.mark.parametrize("B,T,chunk", [(3, 300, 128), (130, 140, 64)])
def test_lstm_chunk_launches_stay_inside_output_and_state(B, T, chunk):
    """Chunked tensor-core LSTM (the wave front's building block): y window of a larger [B, T, H] tensor, carried cell state."""
    H = 512
    lstm = M.ResLSTM(H, num_layers=1).to(DEV)
    img = lstm.lstm.recurrent_image_for(0, "bf16x3")
    pre = (torch.randn(B, T, 4 * H, generator=gen(B + T)) * 0.5).to(DEV)
    ref = ops.lstm_recurrent_tc(pre, img, None, "bf16x3", ops.lstm_tc_max_batch(H, "bf16x3"))
    gy = Guarded((B, T, H))
    rows = (B + 127) // 128 * 128
    gc = Guarded((rows, H))
    ws = ops.lstm_tc_workspace(B, H, "bf16x3", DEV)
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        ops.lstm_recurrent_tc_chunk(pre[:, t0:t1], img, None, gy.view[:, t0:t1], ws, gc.view, t0, "bf16x3")
    assert gy.intact() and gc.intact(), "write outside the output / state window"
    assert torch.equal(gy.view, ref)


@pytest.mark.parametrize("N,K", [(8192 + 77, 8192), (20000, 1000), (300, 512)])
def test_vq_and_fsq_encoders_stay_inside_their_outputs(N, K):
    g = gen(N + K)
    layer = FactorizedVectorQuantize(dim=64, codebook_size=K, codebook_dim=8, commitment=0.25).eval()
    layer._codebook.weight.data = torch.randn(K, 8, generator=g)
    layer = layer.to(DEV)
    z = torch.randn(N, 64, generator=g).to(DEV)
    w_in, b_in = layer._proj("in_proj")
    _, cbn = layer._codebooks()
    lib = _cabi.load_library()
    gi, gm, ge = Guarded((N,), torch.int32), Guarded((N,)), Guarded((N, 8))
    check(lib.bc_vq_encode(ptr(z), ptr(w_in), ptr(b_in), ptr(cbn), ptr(gi.view), ptr(gm.view), ptr(ge.view), N, 64, 8, K, stream_ptr(z.device)), "bc_vq_encode")
    assert gi.intact() and gm.intact() and ge.intact()
    idx, margin, z_e = ops.vq_encode(z, w_in, b_in, cbn, want_margin=True, want_ze=True)
    assert torch.equal(gi.view, idx) and torch.equal(gm.view, margin) and torch.equal(ge.view, z_e)
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    q = FSQ([8, 5, 5, 5], dim=64, channel_first=True).to(DEV)
    w, b = q._proj("project_in")
    gi2, gc2, gb2 = Guarded((N,), torch.int32), Guarded((N, 4)), Guarded((N,))
    check(lib.bc_fsq_encode(ptr(z), ptr(w), ptr(b), ptr(q._kernel_params()), ptr(gi2.view), ptr(gc2.view), ptr(gb2.view), N, 64, 4, stream_ptr(z.device)), "bc_fsq_encode")
    assert gi2.intact() and gc2.intact() and gb2.intact()
    assert int(gi2.view.min()) >= 0 and int(gi2.view.max()) < 1000


@pytest.mark.parametrize("B,T,C", [(2, 126 * 3 + 5, 32), (1, 7, 64), (3, 1000, 16)])
def test_antialiased_activation_stays_inside_its_output(B, T, C):
    g = gen(B + T + C)
    x = (torch.randn(B, T, C, generator=g) * 1.5).to(DEV)
    a = torch.exp(torch.randn(C, generator=g) * 0.3).to(DEV)
    ib = (1.0 / (torch.exp(torch.randn(C, generator=g) * 0.3) + 1e-9)).to(DEV)
    fir = synth.kaiser_sinc_filter12().reshape(-1).to(DEV)
    gy = Guarded((B, T, C))
    check(_cabi.load_library().bc_snake_fwd(ptr(x), ptr(gy.view), ptr(a), ptr(ib), ptr(fir), B, T, C, 1, stream_ptr(x.device)), "bc_snake_fwd")
    assert gy.intact()
    assert torch.equal(gy.view, ops.snake(x, a, ib, antialias=True, fir=fir))
