"""GPU: per-operator parity of the CUDA kernels (through the C ABI) against the CPU oracle on
the same seeded inputs.  Float32 mode: tolerances are float32 rounding (<= 2e-5 relative, far
inside the 1e-3 the north star states); integer outputs are bit-exact (margin-gated for argmax)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from audiotokenization_b200 import configs, ops, synth
from audiotokenization_b200.vq import module as M
from audiotokenization_b200.vq import (Activation1d, BigCodecDecoder, FactorizedVectorQuantize, ResidualVQ, SnakeBeta)
from oracle import bigcodec_oracle as oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gen(seed=0):
    return torch.Generator().manual_seed(seed)


def test_device_is_sm100():
    from audiotokenization_b200 import _cabi
    info = _cabi.device_info(0)
    assert info["cc"][0] == 10 and info["sm_count"] >= 100


@pytest.mark.parametrize("B,C,T", [(1, 1, 17), (2, 32, 1000), (3, 33, 129), (1, 512, 80), (2, 7, 31)])
def test_transpose_round_trip(B, C, T):
    x = torch.randn(B, C, T, generator=gen(1)).to(DEV)
    cl = ops.to_channels_last(x)
    assert cl.shape == (B, T, C) and torch.equal(cl, x.permute(0, 2, 1))
    back = ops.to_channels_first(cl.contiguous())
    assert torch.equal(back, x)


@pytest.mark.parametrize("aa", [False, True])
@pytest.mark.parametrize("B,C,T", [(1, 8, 1), (2, 32, 5), (1, 33, 31), (2, 64, 33), (1, 32, 1000), (3, 128, 257),
                                   (1, 4, 4099), (1, 16, 126), (2, 32, 127), (2, 64, 756), (1, 256, 13), (1, 32, 6),
                                   (1, 512, 2), (2, 16, 2521)])
def test_activation1d_matches_oracle(aa, B, C, T):
    g = gen(B * 1000 + C + T)
    act = Activation1d(SnakeBeta(C, alpha_logscale=True), antialias=aa)
    act.act.alpha.data = torch.randn(C, generator=g) * 0.3
    act.act.beta.data = torch.randn(C, generator=g) * 0.3
    x = torch.randn(B, C, T, generator=g) * 1.5
    sd = {"act.alpha": act.act.alpha.data.clone(), "act.beta": act.act.beta.data.clone()}
    want = oracle.activation1d(sd, "", x, aa)
    got = act.to(DEV)(x.to(DEV))
    assert got.shape == want.shape
    assert rel(got, want) <= TOL
    assert float((got.cpu() - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))


CONV_CASES = [
    # cin, cout, k, stride, dil, pad, causal, T
    (1, 32, 7, 1, 1, 3, False, 1000),
    (32, 32, 7, 1, 1, 3, False, 777),
    (32, 32, 7, 1, 3, 9, False, 300),
    (32, 32, 7, 1, 9, 27, False, 129),
    (32, 32, 7, 1, 9, 27, False, 20),       # shorter than the receptive field
    (64, 64, 1, 1, 1, 0, False, 513),
    (32, 64, 4, 2, 1, 1, False, 1001),
    (64, 128, 8, 4, 1, 2, False, 1000),
    (128, 256, 10, 5, 1, 3, False, 999),
    (256, 512, 10, 5, 1, 3, False, 400),
    (512, 512, 3, 1, 1, 1, False, 80),
    (512, 512, 7, 1, 1, 3, False, 50),
    (32, 1, 7, 1, 1, 3, False, 1500),
    (3, 5, 7, 1, 3, 9, False, 200),         # odd channel counts: scalar paths
    (16, 16, 7, 1, 3, 0, True, 333),        # causal dilated
    (16, 32, 4, 2, 1, 0, True, 335),        # causal strided
    (8, 8, 10, 5, 1, 0, True, 64),
]


@pytest.mark.parametrize("cin,cout,k,stride,dil,pad,causal,T", CONV_CASES)
@pytest.mark.parametrize("fused_act", [False, True])
def test_conv1d_matches_oracle(cin, cout, k, stride, dil, pad, causal, T, fused_act):
    g = gen(cin * 7 + cout + k + T)
    conv = M.WNConv1d(cin, cout, kernel_size=k, stride=stride, dilation=dil, padding=pad, causal=causal)
    inner = conv.conv if causal else conv
    inner.weight_g.data *= torch.exp(torch.randn(cout, 1, 1, generator=g) * 0.2)
    inner.bias.data = torch.randn(cout, generator=g) * 0.2
    x = torch.randn(2, cin, T, generator=g)
    sd = {("conv." if causal else "") + n: getattr(inner, n).data.clone() for n in ("weight_g", "weight_v", "bias")}
    xin = x
    act = None
    if fused_act:
        act = SnakeBeta(cin, alpha_logscale=True)
        act.alpha.data = torch.randn(cin, generator=g) * 0.3
        act.beta.data = torch.randn(cin, generator=g) * 0.3
        xin = oracle.snake_beta(x, act.alpha.data, act.beta.data)
        act = act.to(DEV)
    want = oracle.wn_conv1d(sd, "", xin, stride=stride, dilation=dil, padding=pad, causal=causal)
    conv = conv.to(DEV)
    got_cl = conv.forward_cl(ops.to_channels_last(x.to(DEV)), act=act)
    got = got_cl.permute(0, 2, 1)
    assert got.shape == want.shape
    assert rel(got, want) <= TOL


def test_conv1d_residual_and_tanh_epilogues():
    g = gen(5)
    conv = M.WNConv1d(32, 32, kernel_size=1)
    conv.bias.data = torch.randn(32, generator=g) * 0.1
    x = torch.randn(2, 32, 300, generator=g)
    r = torch.randn(2, 32, 300, generator=g)
    sd = {n: getattr(conv, n).data.clone() for n in ("weight_g", "weight_v", "bias")}
    want = oracle.wn_conv1d(sd, "", x) + r
    conv = conv.to(DEV)
    got = conv.forward_cl(ops.to_channels_last(x.to(DEV)), res=ops.to_channels_last(r.to(DEV))).permute(0, 2, 1)
    assert rel(got, want) <= TOL
    got_t = conv.forward_cl(ops.to_channels_last(x.to(DEV)), tanh=True).permute(0, 2, 1)
    assert rel(got_t, torch.tanh(oracle.wn_conv1d(sd, "", x))) <= TOL


@pytest.mark.parametrize("cin,cout,stride,causal,T", [(64, 32, 2, False, 501), (128, 64, 4, False, 250),
                                                      (512, 256, 5, False, 80), (256, 128, 5, False, 33),
                                                      (16, 8, 2, True, 100), (12, 6, 3, False, 50), (8, 4, 5, True, 7)])
def test_conv_transpose1d_matches_oracle(cin, cout, stride, causal, T):
    g = gen(cin + cout + stride + T)
    if causal:
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, causal=True)
        inner, pre = m.conv, "conv."
    else:
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, padding=stride // 2 + stride % 2,
                                output_padding=stride % 2)
        inner, pre = m, ""
    inner.weight_g.data *= torch.exp(torch.randn(cin, 1, 1, generator=g) * 0.2)
    x = torch.randn(2, cin, T, generator=g)
    act = SnakeBeta(cin, alpha_logscale=True)
    act.alpha.data = torch.randn(cin, generator=g) * 0.3
    act.beta.data = torch.randn(cin, generator=g) * 0.3
    sd = {pre + n: getattr(inner, n).data.clone() for n in ("weight_g", "weight_v", "bias")}
    want = oracle.wn_conv_transpose1d(sd, "", oracle.snake_beta(x, act.alpha.data, act.beta.data), stride=stride,
                                      causal=causal)
    got = m.to(DEV).forward_cl(ops.to_channels_last(x.to(DEV)), act=act.to(DEV)).permute(0, 2, 1)
    assert got.shape == want.shape == (2, cout, T * stride)
    assert rel(got, want) <= TOL


@pytest.mark.parametrize("H,layers,B,T", [(64, 2, 1, 1), (64, 2, 3, 17), (64, 1, 33, 40), (512, 2, 2, 60),
                                          (128, 2, 70, 9), (768, 1, 3, 12), (1536, 2, 2, 10), (1536, 1, 40, 5)])
def test_res_lstm_matches_oracle(H, layers, B, T):
    """H <= 592: W_hh resident in shared memory; H = 768 / 1536 (ngf 48, the original BigCodec config): rows
    streamed from L2, several unit groups per CTA, cell state in the workspace."""
    g = gen(H + layers + B + T)
    m = M.ResLSTM(H, num_layers=layers)
    sd = {"lstm." + k: v.data.clone() for k, v in m.lstm.named_parameters()}
    x = torch.randn(B, H, T, generator=g)
    want = oracle.res_lstm(sd, "", x, layers)
    got = m.to(DEV)(x.to(DEV))
    assert got.shape == want.shape
    assert rel(got, want) <= 5e-5


def _vq_layer(C, D, K, seed):
    g = gen(seed)
    layer = FactorizedVectorQuantize(dim=C, codebook_size=K, codebook_dim=D, commitment=0.25).eval()
    layer._codebook.weight.data = torch.randn(K, D, generator=g)
    sd = {"_codebook.weight": layer._codebook.weight.data.clone()}
    if C != D:
        layer.in_proj.weight_g.data *= 1.3
        for p in ("in_proj", "out_proj"):
            for n in ("weight_g", "weight_v", "bias"):
                sd[f"{p}.{n}"] = getattr(getattr(layer, p), n).data.clone()
    return layer, sd


@pytest.mark.parametrize("C,D,K,N", [(512, 8, 8192, 800), (64, 8, 512, 203), (512, 8, 32768, 130), (8, 8, 1024, 77),
                                     (1024, 8, 16384, 64), (32, 4, 100, 50), (48, 16, 333, 41)])
def test_vq_encode_indices_bit_exact_where_margin_allows(C, D, K, N):
    layer, sd = _vq_layer(C, D, K, C + K + N)
    z = torch.randn(1, C, N, generator=gen(N))
    z_q_want, idx_want, _, margin = oracle.vq_layer_forward(sd, "", z)
    layer = layer.to(DEV)
    idx, m_got, z_e = layer.encode_cl(ops.to_channels_last(z.to(DEV)), want_margin=True, want_ze=True)
    idx, m_got = idx.cpu().long(), m_got.cpu()
    decided = margin > 1e-5
    assert bool(decided.float().mean() > 0.95)
    assert torch.equal(idx[decided], idx_want[decided]), "index mismatch on a frame with margin > 1e-5"
    assert float((idx == idx_want).float().mean()) >= 0.999
    assert float((m_got - margin).abs().max()) <= 5e-6
    # full forward: quantised output and the reference's return triple
    z_q, idx2, loss = layer(z.to(DEV))
    assert z_q.shape == z.shape and idx2.dtype == torch.int64 and float(loss.abs().sum()) == 0.0
    same = (idx2.cpu() == idx_want).all(dim=0)
    assert rel(z_q.cpu()[:, :, same], z_q_want[:, :, same]) <= TOL


def test_vq_tie_break_is_lowest_index():
    layer, sd = _vq_layer(8, 8, 64, 3)
    cb = layer._codebook.weight.data
    cb[40] = cb[7]          # duplicate code: both attain the maximum
    z = cb[7].view(1, 8, 1).repeat(1, 1, 5) * 2.0
    layer = layer.to(DEV)
    idx, margin, _ = layer.encode_cl(ops.to_channels_last(z.to(DEV)), want_margin=True)
    assert idx.cpu().tolist() == [[7] * 5]
    assert float(margin.abs().max()) == 0.0


def test_vq_dequant_and_index_entry_points():
    cfg = configs.get_config("tiny")
    _, dsd = synth.make_state_dicts(cfg)
    dec = BigCodecDecoder(**cfg["codec_decoder"])
    dec.load_state_dict(dsd)
    dec = dec.to(DEV)
    codes = torch.randint(0, 512, (2, 37, 1), generator=gen(2))
    want = oracle.vq2emb(dsd, cfg["codec_decoder"], codes)
    got = dec.vq2emb(codes.to(DEV))
    assert got.shape == (2, 37, 64) and rel(got, want) <= TOL
    assert torch.equal(dec.get_emb()[0].cpu(), dsd["quantizer.layers.0._codebook.weight"])
    emb = dec.quantizer.layers[0].embed_code(codes[:, :, 0].to(DEV))
    assert torch.equal(emb.cpu(), F.embedding(codes[:, :, 0], dsd["quantizer.layers.0._codebook.weight"]))
    with pytest.raises(IndexError):
        dec.vq2emb(torch.full((1, 3, 1), 512).to(DEV))
    y = dec.inference_vq(got[0].transpose(0, 1).contiguous())
    assert y.shape == (1, 1, 37 * 40)


def test_residual_vq_two_quantizers_matches_oracle():
    cfg = configs.get_config("tiny")
    cfg["codec_decoder"]["vq_num_quantizers"] = 2
    _, dsd = synth.make_state_dicts(cfg)
    dec = BigCodecDecoder(**cfg["codec_decoder"])
    dec.load_state_dict(dsd)
    z = torch.randn(2, 64, 50, generator=gen(9))
    zq_want, idx_want, loss_want, margin = oracle.quantize(dsd, cfg["codec_decoder"], z)
    zq, idx, loss = dec.to(DEV)(z.to(DEV), vq=True)
    assert idx.shape == (2, 2, 50) and loss.shape == (2,)
    ok = (margin > 1e-5).all(dim=0)
    assert torch.equal(idx.cpu()[:, ok], idx_want[:, ok])
    same = (idx.cpu() == idx_want).all(dim=0)
    assert rel(zq.cpu().permute(0, 2, 1)[same], zq_want.permute(0, 2, 1)[same]) <= TOL


def test_indices_to_int16_layout():
    idx = torch.randint(0, 8192, (2, 91), generator=gen(4), dtype=torch.int32)
    out = ops.indices_to_int16(idx.to(DEV)).cpu().numpy()
    assert out.dtype == np.int16 and np.array_equal(out, idx.numpy().T.astype(np.int16))


def test_errors_are_exceptions_not_fallbacks():
    conv = M.WNConv1d(4, 4, kernel_size=3, padding=1).to(DEV)
    with pytest.raises(ValueError):
        conv.forward_cl(torch.zeros(1, 10, 5, device=DEV))            # wrong channel count
    with pytest.raises(TypeError):
        ops.to_channels_last(torch.zeros(1, 4, 10, device=DEV, dtype=torch.float64))
    strided = M.WNConv1d(4, 4, kernel_size=10, stride=5, padding=3).to(DEV)
    with pytest.raises(ValueError):
        strided.forward_cl(torch.zeros(1, 2, 4, device=DEV))          # too short for the layer


# ---------------------------------------------------------------------------------------------
# tensor-core (tcgen05) modes
# ---------------------------------------------------------------------------------------------
TC_CASES = [
    # cin, cout, k, stride, dil, pad, T
    (32, 32, 7, 1, 1, 3, 777),
    (32, 32, 7, 1, 9, 27, 300),
    (64, 64, 7, 1, 3, 9, 1000),
    (64, 64, 1, 1, 1, 0, 513),
    (128, 128, 7, 1, 9, 27, 260),
    (256, 256, 7, 1, 3, 9, 140),
    (32, 64, 4, 2, 1, 1, 1001),
    (64, 128, 8, 4, 1, 2, 1000),
    (128, 256, 10, 5, 1, 3, 999),
    (256, 512, 10, 5, 1, 3, 400),
    (512, 512, 3, 1, 1, 1, 80),
    (512, 2048, 1, 1, 1, 0, 100),
    (48, 96, 4, 2, 1, 1, 300),
    (16, 16, 7, 1, 3, 9, 50),
]


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


@pytest.mark.parametrize("cin,cout,k,stride,dil,pad,T", TC_CASES)
@pytest.mark.parametrize("precision,fused_act", [("bf16x3", True), ("bf16", False), ("bf16", True), ("bf16x3", False)])
def test_conv1d_tensor_core_modes(cin, cout, k, stride, dil, pad, T, precision, fused_act):
    """bf16x3 must be fp32-class (<= 3e-5 vs the float64 oracle); single-pass bf16 is compared with an
    oracle whose operands are rounded to bf16 the same way (<= 5e-5), and is ~4e-3 from the true result."""
    g = gen(cin * 7 + cout + k + T)
    conv = M.WNConv1d(cin, cout, kernel_size=k, stride=stride, dilation=dil, padding=pad)
    conv.weight_g.data *= torch.exp(torch.randn(cout, 1, 1, generator=g) * 0.2)
    conv.bias.data = torch.randn(cout, generator=g) * 0.2
    x = torch.randn(2, cin, T, generator=g)
    r = torch.randn(2, cout, conv.out_length(T), generator=g)
    w = oracle.fold_weight_norm(conv.weight_g.double(), conv.weight_v.double())
    xin = x.double()
    act = None
    if fused_act:
        act = SnakeBeta(cin, alpha_logscale=True)
        act.alpha.data = torch.randn(cin, generator=g) * 0.3
        act.beta.data = torch.randn(cin, generator=g) * 0.3
        xin = oracle.snake_beta(x, act.alpha.data, act.beta.data).double()   # fp32 snake like the kernel
        act = act.to(DEV)
    exact = F.conv1d(xin, w, conv.bias.double(), stride=stride, dilation=dil, padding=pad) + r.double()
    M.set_precision(precision)
    try:
        conv = conv.to(DEV)
        assert conv.packed_for(precision)[2] == precision, "geometry unexpectedly fell back to fp32"
        got = conv.forward_cl(ops.to_channels_last(x.to(DEV)), act=act,
                              res=ops.to_channels_last(r.to(DEV))).permute(0, 2, 1)
    finally:
        M.set_precision("fp32")
    assert got.shape == exact.shape
    if precision == "bf16x3":
        assert rel(got, exact) <= 3e-5
    else:
        emul = F.conv1d(_bf16(xin.float()).double(), _bf16(w.float()).double(), conv.bias.double().cpu(),
                        stride=stride, dilation=dil, padding=pad) + r.double()
        assert rel(got, emul) <= 5e-5
        assert rel(got, exact) <= 1e-2


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_edge_convs_stay_on_fp32_kernel_in_tensor_core_modes(precision):
    stem = M.WNConv1d(1, 32, kernel_size=7, padding=3).to(DEV)
    assert stem.packed_for(precision)[2] == "fp32"
    tail = M.WNConv1d(32, 1, kernel_size=7, padding=3).to(DEV)
    assert tail.packed_for(precision)[2] == "fp32"


@pytest.mark.parametrize("precision,tol", [("bf16x3", 3e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("cin,cout,stride,causal,T", [(64, 32, 2, False, 501), (128, 64, 4, False, 250), (512, 256, 5, False, 80),
                                                      (256, 128, 5, False, 1000), (128, 64, 4, True, 300), (64, 32, 2, True, 129),
                                                      (256, 128, 2, False, 77), (96, 48, 3, False, 200)])
def test_conv_transpose1d_tensor_core_modes(precision, tol, cin, cout, stride, causal, T):
    """Single-launch form on the streamed-weight kernel (all phases as channel blocks of one 2-tap conv, q = 1 phases
    shifted per n-tile; three zero-padded taps when an n-tile straddles phases, e.g. 64 -> 32), the causal variant, and
    agreement with the per-phase path."""
    g = gen(cin + cout + stride + T)
    if causal:
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, causal=True)
        inner = m.conv
    else:
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, padding=stride // 2 + stride % 2,
                                output_padding=stride % 2)
        inner = m
    x = torch.randn(2, cin, T, generator=g)
    act = SnakeBeta(cin, alpha_logscale=True)
    act.alpha.data = torch.randn(cin, generator=g) * 0.3
    act.beta.data = torch.randn(cin, generator=g) * 0.3
    pre = "conv." if causal else ""
    sd = {pre + n: getattr(inner, n).data.clone().double() for n in ("weight_g", "weight_v", "bias")}
    want = oracle.wn_conv_transpose1d(sd, "", oracle.snake_beta(x, act.alpha.data, act.beta.data).double(),
                                      stride=stride, causal=causal)
    M.set_precision(precision)
    try:
        m = m.to(DEV)
        streamed = inner.stream_image(precision) is not None
        got = m.forward_cl(ops.to_channels_last(x.to(DEV)), act=act.to(DEV)).permute(0, 2, 1)
        M.CONVTR_STREAM[0] = False
        per_phase = m.forward_cl(ops.to_channels_last(x.to(DEV)), act=act.to(DEV)).permute(0, 2, 1)
    finally:
        M.CONVTR_STREAM[0] = True
        M.set_precision("fp32")
    assert streamed == (cin % 16 == 0 and cin >= 32 and (stride * cout) % 64 == 0), streamed
    assert got.shape == want.shape
    assert rel(got, want) <= tol
    assert rel(got, per_phase) <= tol


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("bf16", 3e-2)])
def test_res_lstm_tensor_core_input_projection(precision, tol):
    m = M.ResLSTM(512, num_layers=2)
    sd = {"lstm." + k: v.data.clone() for k, v in m.lstm.named_parameters()}
    x = torch.randn(2, 512, 60, generator=gen(11))
    want = oracle.res_lstm(sd, "", x, 2)
    M.set_precision(precision)
    try:
        got = m.to(DEV)(x.to(DEV))
    finally:
        M.set_precision("fp32")
    assert rel(got, want) <= tol


@pytest.mark.parametrize("precision,tol", [("bf16x3", 5e-5), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("C,dil,causal,T", [(32, 1, False, 1000), (32, 9, False, 300), (64, 3, False, 777),
                                            (128, 9, False, 260), (64, 9, False, 30), (16, 3, True, 333),
                                            (48, 1, False, 200), (96, 3, False, 150)])
def test_fused_residual_unit_matches_oracle(precision, tol, C, dil, causal, T):
    """One-kernel ResidualUnit (conv7 -> snake -> conv1 -> +x, intermediate kept in TMEM/smem) against the
    float64 oracle, and against the unfused two-kernel path of the same precision."""
    g = gen(C + dil + T)
    ru = M.ResidualUnit(C, dilation=dil, causal=causal)
    sd = {}
    for name, prm in ru.named_parameters():
        if name.endswith(("alpha", "beta")):
            prm.data = torch.randn(prm.shape, generator=g) * 0.3
        elif name.endswith("bias"):
            prm.data = torch.randn(prm.shape, generator=g) * 0.2
        elif name.endswith("weight_g"):
            prm.data = prm.data * torch.exp(torch.randn(prm.shape, generator=g) * 0.2)
    for name, prm in ru.named_parameters():
        sd[name] = prm.data.clone().double()
    x = torch.randn(2, C, T, generator=g)
    want = oracle.residual_unit(sd, "", x.double(), dil, causal, False)
    ru = ru.to(DEV)
    M.set_precision(precision)
    try:
        if not (precision == "bf16x3" and C > 64):   # C=128 split mode: too much smem for 2 CTAs/SM -> two launches
            assert ru._fused_plan(precision) is not None
        got = ru(x.to(DEV))
        M.FUSE_RESUNIT[0] = False
        unfused = ru(x.to(DEV))
    finally:
        M.FUSE_RESUNIT[0] = True
        M.set_precision("fp32")
    assert got.shape == want.shape
    assert rel(got, want) <= tol
    assert rel(got, unfused) <= tol


_PAIR_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {repo!r}); sys.path.insert(0, {repo!r} + "/tests")
from audiotokenization_b200.vq import module as M
from oracle import bigcodec_oracle as oracle
C = 64
for dil, B, T in [(1, 3, 40000), (3, 1, 128 * 151), (9, 5, 9000), (9, 1, 100), (3, 2, 20000)]:
    g = torch.Generator().manual_seed(1000 + dil + B + T)
    ru = M.ResidualUnit(C, dilation=dil)
    for name, prm in ru.named_parameters():
        if name.endswith(("alpha", "beta")):
            prm.data = torch.randn(prm.shape, generator=g) * 0.3
        elif name.endswith("bias"):
            prm.data = torch.randn(prm.shape, generator=g) * 0.2
    sd = {{name: prm.data.clone().double() for name, prm in ru.named_parameters()}}
    x = torch.randn(B, C, T, generator=g)
    want = oracle.residual_unit(sd, "", x.double(), dil, False, False)
    ru = ru.cuda()
    M.set_precision("bf16x3")
    plan = ru._fused_plan("bf16x3")
    assert plan is not None and plan[4] in (3, 4), plan          # the CTA-pair kernel is what runs
    got = ru(x.cuda()); again = ru(x.cuda())
    M.FUSE_RESUNIT[0] = False
    unfused = ru(x.cuda())
    M.FUSE_RESUNIT[0] = True
    rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
    assert torch.equal(got, again)
    assert rel(got, want) <= 5e-5 and rel(got, unfused) <= 5e-5, (rel(got, want), rel(got, unfused))
    err = (got.cpu().double() - want).abs().amax(dim=1)           # per time step: no tile may be wrong
    assert float(err.max()) <= 2e-4 * max(1.0, float(want.abs().max())), (dil, B, T, int(err.argmax()))
    print("pair ok", dil, B, T, plan[4])
"""


def test_cta_pair_residual_unit_many_tiles_odd_counts():
    """ru_pair.cu (tcgen05 cta_group::2, C = 64, split precision; the default, pinned here with BC_RU_PAIR=1): more
    tile pairs than CTA pairs (every pair walks several rounds through both activation slots and accumulator stages),
    odd tile counts (the phantom tile of rank 1), a single tile, tiles that straddle items, stacked and plain weight
    images -- against the float64 oracle and the unfused two-launch path."""
    import os, subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BC_RU_PAIR="1")
    r = subprocess.run([sys.executable, "-c", _PAIR_SCRIPT.format(repo=repo)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("pair ok") == 5, r.stdout


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("H,layers,B,T", [(128, 2, 1, 1), (128, 1, 3, 40), (512, 2, 130, 25), (256, 2, 33, 64),
                                          (512, 1, 257, 7), (512, 2, 512, 30), (512, 1, 300, 33), (256, 1, 385, 12)])
def test_res_lstm_tensor_core_recurrence(precision, tol, H, layers, B, T):
    """tcgen05 recurrence (W_hh resident in smem, h exchanged as bf16 hi/lo images, per-tile step counters):
    several batch tiles, partial tiles, one-step sequences; and it must agree with the CUDA-core kernel."""
    g = gen(H + layers + B + T)
    m = M.ResLSTM(H, num_layers=layers)
    sd = {"lstm." + k: v.data.clone() for k, v in m.lstm.named_parameters()}
    x = torch.randn(B, H, T, generator=g)
    want = oracle.res_lstm(sd, "", x, layers)
    m = m.to(DEV)
    M.set_precision(precision)
    try:
        assert ops.lstm_tc_max_batch(H, precision) >= 128
        got = m(x.to(DEV))
        M.LSTM_TENSOR_CORE[0] = False
        ref = m(x.to(DEV))
    finally:
        M.LSTM_TENSOR_CORE[0] = True
        M.set_precision("fp32")
    assert got.shape == want.shape
    assert rel(got, want) <= tol
    assert rel(got, ref) <= tol


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("H,B,T", [(512, 64, 800), (512, 1, 1000), (256, 100, 500), (128, 3, 384), (512, 128, 390)])
def test_res_lstm_two_layer_wave_front_is_bit_identical(precision, H, B, T):
    """Small batches: layer 1 runs one chunk behind layer 0 on a second stream (chunked launches with the hidden state in
    the workspace and the cell state carried in a buffer).  Same kernels and the same per-element arithmetic as the
    layer-by-layer schedule: results must be bit-identical, partial last chunks included."""
    g = gen(H + B + T)
    m = M.ResLSTM(H, num_layers=2).to(DEV)
    x = torch.randn(B, H, T, generator=g).to(DEV)
    M.set_precision(precision)
    try:
        assert m._wavefront_chunk(B, T, precision) == 128
        got = m(x)
        again = m(x)
        M.LSTM_WAVEFRONT[0] = False
        ref = m(x)
    finally:
        M.LSTM_WAVEFRONT[0] = True
        M.set_precision("fp32")
    assert torch.equal(got, again)
    assert torch.equal(got, ref)
    sd = {"lstm." + k: v.data.clone().cpu() for k, v in m.lstm.named_parameters()}
    if B * T <= 60000:
        want = oracle.res_lstm(sd, "", x.cpu(), 2)
        assert rel(got, want) <= (1e-4 if precision == "bf16x3" else 3e-2)


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_res_lstm_tensor_core_gates_saturate_instead_of_overflowing(precision):
    """the cell state is unbounded (|c| grows by up to 1 per step when i, f saturate), and a
    (1 - e) / (1 + e) tanh with e = exp(-2x) turns into 0 or NaN for x < -44.  Drive |c| and the g gate far past
    that (pre-activations up to 1e3, 120 steps) and require finite output that matches the float32 oracle."""
    H, B, T = 128, 5, 120
    g = gen(77)
    m = M.ResLSTM(H, num_layers=1)
    with torch.no_grad():
        m.lstm.weight_hh_l0.mul_(0.05)
        # input gate and forget gate wide open, cell candidate strongly negative for half of the units
        b = m.lstm.bias_ih_l0
        b[:H] = 50.0
        b[H:2 * H] = 50.0
        b[2 * H:3 * H] = torch.where(torch.arange(H) % 2 == 0, torch.tensor(-1000.0), torch.tensor(1000.0))
        b[3 * H:] = 30.0
        m.lstm.weight_ih_l0.mul_(0.01)
    sd = {"lstm." + k: v.data.clone() for k, v in m.lstm.named_parameters()}
    x = torch.randn(B, H, T, generator=g)
    want = oracle.res_lstm(sd, "", x, 1)          # c reaches -120 / +120: tanh(c) = -1 / +1 exactly in float32
    assert torch.isfinite(want).all() and float((want - x).abs().max()) > 0.99
    m = m.to(DEV)
    M.set_precision(precision)
    try:
        got = m(x.to(DEV))
    finally:
        M.set_precision("fp32")
    assert torch.isfinite(got).all(), "NaN / inf out of the tensor-core LSTM"
    assert rel(got, want) <= 1e-4


# ---------------------------------------------------------------------------------------------
# persistent streamed-weight kernels (csrc/conv_stream.cu)
# ---------------------------------------------------------------------------------------------
STREAM_CONV_CASES = [
    # cin, cout, k, stride, dil, pad, B, T
    (128, 128, 7, 1, 9, 27, 2, 260),
    (256, 256, 7, 1, 3, 9, 1, 140),
    (64, 128, 8, 4, 1, 2, 2, 1000),
    (128, 256, 10, 5, 1, 3, 3, 999),
    (256, 512, 10, 5, 1, 3, 2, 400),
    (512, 512, 3, 1, 1, 1, 2, 80),
    (512, 2048, 1, 1, 1, 0, 1, 100),
    (64, 64, 1, 1, 1, 0, 2, 513),
    (128, 128, 7, 1, 1, 3, 5, 5000),      # more tiles than SMs: every CTA walks several tiles
    (256, 256, 1, 1, 1, 0, 3, 9000),
    (32, 64, 4, 2, 1, 1, 3, 10),          # T_out = 5: a store tile larger than the item (TMA clips), stride-2 slab
    (64, 128, 8, 4, 1, 2, 1, 4),          # one output step
    (512, 512, 3, 1, 1, 1, 130, 3),       # many tiny items
]


@pytest.mark.parametrize("cin,cout,k,stride,dil,pad,B,T", STREAM_CONV_CASES)
@pytest.mark.parametrize("precision,fused_act", [("bf16x3", True), ("bf16", True), ("bf16x3", False)])
def test_stream_conv_matches_oracle_and_tile_kernel(cin, cout, k, stride, dil, pad, B, T, precision, fused_act):
    g = gen(cin * 5 + cout + k + T)
    conv = M.WNConv1d(cin, cout, kernel_size=k, stride=stride, dilation=dil, padding=pad)
    conv.weight_g.data *= torch.exp(torch.randn(cout, 1, 1, generator=g) * 0.2)
    conv.bias.data = torch.randn(cout, generator=g) * 0.2
    x = torch.randn(B, cin, T, generator=g)
    r = torch.randn(B, cout, conv.out_length(T), generator=g)
    w = oracle.fold_weight_norm(conv.weight_g.double(), conv.weight_v.double())
    xin = x.double()
    act = None
    if fused_act:
        act = SnakeBeta(cin, alpha_logscale=True)
        act.alpha.data = torch.randn(cin, generator=g) * 0.3
        act.beta.data = torch.randn(cin, generator=g) * 0.3
        xin = oracle.snake_beta(x, act.alpha.data, act.beta.data).double()
        act = act.to(DEV)
    exact = F.conv1d(xin, w, conv.bias.double(), stride=stride, dilation=dil, padding=pad) + r.double()
    conv = conv.to(DEV)
    xc, rc = ops.to_channels_last(x.to(DEV)), ops.to_channels_last(r.to(DEV))
    M.set_precision(precision)
    try:
        assert M._stream_tile(cin, cout, k, stride, dil, precision) is not None
        got = conv.forward_cl(xc, act=act, res=rc).permute(0, 2, 1)
        M.STREAM[0] = False
        tile = conv.forward_cl(xc, act=act, res=rc).permute(0, 2, 1)
    finally:
        M.STREAM[0] = True
        M.set_precision("fp32")
    assert got.shape == exact.shape
    if precision == "bf16x3":
        assert rel(got, exact) <= 3e-5
    else:
        assert rel(got, exact) <= 1e-2
    assert rel(got, tile) <= (2e-6 if precision == "bf16" else 3e-5)   # same operands, other summation order


@pytest.mark.parametrize("precision,tol", [("bf16x3", 5e-5), ("bf16", 1.5e-2)])
@pytest.mark.parametrize("C,dil,causal,B,T", [(128, 1, False, 2, 1000), (128, 9, False, 5, 5000), (256, 3, False, 2, 777),
                                              (256, 9, False, 3, 7000), (64, 9, False, 2, 300), (128, 3, True, 1, 333),
                                              (128, 9, False, 1, 30)])
def test_stream_residual_unit_matches_oracle(precision, tol, C, dil, causal, B, T):
    g = gen(C + dil + T)
    ru = M.ResidualUnit(C, dilation=dil, causal=causal)
    sd = {}
    for name, prm in ru.named_parameters():
        if name.endswith(("alpha", "beta")):
            prm.data = torch.randn(prm.shape, generator=g) * 0.3
        elif name.endswith("bias"):
            prm.data = torch.randn(prm.shape, generator=g) * 0.2
        elif name.endswith("weight_g"):
            prm.data = prm.data * torch.exp(torch.randn(prm.shape, generator=g) * 0.2)
    for name, prm in ru.named_parameters():
        sd[name] = prm.data.clone().double()
    x = torch.randn(B, C, T, generator=g)
    want = oracle.residual_unit(sd, "", x.double(), dil, causal, False)
    ru = ru.to(DEV)
    M.set_precision(precision)
    old_min = M.STREAM_RU_MIN_C[0]
    try:
        M.STREAM_RU_MIN_C[0] = 64
        assert M._stream_tile(C, C, 7, 1, dil, precision, fused=True) is not None
        got = ru(x.to(DEV))
        M.STREAM[0] = False
        other = ru(x.to(DEV))
    finally:
        M.STREAM[0] = True
        M.STREAM_RU_MIN_C[0] = old_min
        M.set_precision("fp32")
    assert got.shape == want.shape
    assert rel(got, want) <= tol
    assert rel(got, other) <= tol


@pytest.mark.parametrize("levels,C,B,T", [([4, 4, 4, 8], 64, 2, 500), ([8, 5, 5, 5], 512, 3, 257), ([7, 5, 5, 5, 5], 96, 1, 1000),
                                          ([3], 32, 2, 33), ([8, 8, 8, 6, 5], 1024, 1, 130), ([2, 2], 2, 2, 64)])
def test_fsq_quantizer_matches_oracle(levels, C, B, T):
    """bc_fsq_encode + dequantisation over the implicit codebook against the oracle restatement of FSQ.forward
    (finite_scalar_quantization.py:203-259): indices bit-exact wherever no bounded component lies within 1e-5 of a
    rounding boundary (tanhf differs from the CPU library's tanh in the last ulp), outputs to float32 rounding;
    indices_to_codes round trip; odd and even level counts, identity projection (dim == number of levels)."""
    from audiotokenization_b200.vq import FSQ
    g = gen(sum(levels) + C + T)
    q = FSQ(levels, dim=C, channel_first=True)
    sd = {}
    if q.has_projections:
        q.project_in.weight.data = torch.randn(len(levels), C, generator=g) * (3.0 / C ** 0.5)
        q.project_in.bias.data = torch.randn(len(levels), generator=g) * 0.3
        sd = {"q." + k: v.data.clone() for k, v in q.named_parameters()}
    z = torch.randn(B, C, T, generator=g) * (1.5 if q.has_projections else 2.0)
    want, widx, wbound = oracle.fsq_forward(sd, "q.", z, levels)
    q = q.to(DEV)
    out, idx = q(z.to(DEV))
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, T) and tuple(out.shape) == (B, C, T)
    _, _, bound = q.encode_cl(ops.to_channels_last(z.to(DEV)), want_boundary=True)
    tol = 1e-5 * max(1.0, C / 256)            # the C-term projection sums in another order than the CPU library's GEMM
    decided = wbound > tol
    assert float(decided.float().mean()) > 0.99
    assert torch.equal(idx.cpu()[decided], widx[decided])
    assert float((bound.cpu() - wbound).abs().max()) <= tol
    same = (idx.cpu() == widx)
    assert rel(out.cpu().permute(0, 2, 1)[same], want.permute(0, 2, 1)[same]) <= 1e-6
    assert len(torch.unique(idx)) > min(q.codebook_size, B * T) // 8
    back = q.indices_to_codes(idx)
    assert torch.equal(back, out)
    with pytest.raises(IndexError):
        q.indices_to_codes(torch.full((1, 4), q.codebook_size, dtype=torch.int32, device=DEV))


@pytest.mark.parametrize("mode", ["0", "2"])
def test_stream_conv_other_tensor_map_modes(mode):
    """The plain streamed-weight convs run with TMA stores by default (BC_STREAM_TMA=1).  The policy is read once per
    process, so the two other forms -- no tensor maps at all (0) and x boxes by TMA as well (2) -- are pinned in a
    subprocess that re-runs the plain-conv, transposed-conv and guard-band cases under that setting."""
    import os, subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BC_STREAM_TMA=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                        os.path.join(repo, "tests", "test_gpu_ops.py"), os.path.join(repo, "tests", "test_gpu_guard_bands.py"),
                        "-k", "stream_conv_matches or conv_transpose1d_tensor_core or single_launch or res_lstm_tensor_core_input"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=repo)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-1000:]


_LSTM_PAIR_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {repo!r})
from audiotokenization_b200 import _cabi, ops
from audiotokenization_b200.vq import module as M
assert "lstm_pair=1" in _cabi.policy(), _cabi.policy()
for H, B, T in ((512, 512, 40), (512, 400, 33), (256, 1024, 12)):
    g = torch.Generator().manual_seed(H + B + T)
    m = M.ResLSTM(H, num_layers=1).cuda()
    img = m.lstm.recurrent_image_for(0, "bf16x3")
    pre = (torch.randn(B, T, 4 * H, generator=g) * 0.7).cuda()
    skip = torch.randn(B, T, H, generator=g).cuda()
    n_slices = ops.lstm_tc_ctas(128, H, "bf16x3")                       # one tile: one CTA per 32-column slice
    small = 128 * (148 // n_slices)                                     # what fits side by side with one tile per CTA
    assert B > small and ops.lstm_tc_ctas(B, H, "bf16x3") == (n_slices // 2) * ((B + 127) // 128)     # the pair plan
    got = ops.lstm_recurrent_tc(pre, img, skip, "bf16x3", ops.lstm_tc_max_batch(H, "bf16x3"))
    want = ops.lstm_recurrent_tc(pre, img, skip, "bf16x3", small)
    assert torch.isfinite(got).all() and torch.equal(got, want), (H, B, T)
    print("lstm pair ok", H, B, T)
"""


def test_res_lstm_cta_pair_recurrence_is_bit_identical():
    """lstm_pair_kernel (tcgen05 cta_group::2: two batch tiles x 64 gate columns per CTA pair; opt-in with BC_LSTM_PAIR=1,
    hence the subprocess) accumulates the same products in the same order as the one-tile-per-CTA kernel, so it must
    reproduce it (the same rows in launches that fit side by side) bit for bit -- partial last tile included."""
    import os, subprocess, sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _LSTM_PAIR_SCRIPT.format(repo=repo)], env=dict(os.environ, BC_LSTM_PAIR="1"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("lstm pair ok") == 3, r.stdout


_LSTM_COMPACT_SCRIPT = r"""
import sys, torch
sys.path.insert(0, {repo!r})
from audiotokenization_b200 import _cabi, ops
from audiotokenization_b200.vq import module as M
assert "lstm_compact=0" in _cabi.policy(), _cabi.policy()
out = {{}}
for H, B, T in {cases!r}:
    g = torch.Generator().manual_seed(H + B + T)
    torch.manual_seed(H + B + T)                  # the module's random initial weights
    m = M.ResLSTM(H, num_layers=1).cuda()
    img = m.lstm.recurrent_image_for(0, "bf16x3")
    pre = (torch.randn(B, T, 4 * H, generator=g) * 0.7).cuda()
    skip = torch.randn(B, T, H, generator=g).cuda()
    out[(H, B, T)] = ops.lstm_recurrent_tc(pre, img, skip, "bf16x3", ops.lstm_tc_max_batch(H, "bf16x3")).cpu()
torch.save(out, {path!r})
"""


def test_res_lstm_compact_exchange_image_is_bit_identical(tmp_path):
    """One-tile launches (B < 128) exchange h in a compact image of round_up(B, 8) rows per 8-channel plane (the MMA's rows
    beyond it alias other planes and only feed accumulator rows nobody stores).  Same arithmetic for the valid rows: the
    results must equal the full 128-row image (BC_LSTM_COMPACT=0, hence the subprocess) bit for bit."""
    import os, subprocess, sys
    from audiotokenization_b200 import _cabi
    assert "lstm_compact=1" in _cabi.policy()
    cases = [(512, 1, 30), (512, 5, 17), (512, 64, 40), (256, 100, 12), (128, 127, 9), (512, 8, 3)]
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = str(tmp_path / "full_image.pt")
    r = subprocess.run([sys.executable, "-c", _LSTM_COMPACT_SCRIPT.format(repo=repo, cases=cases, path=path)],
                       env=dict(os.environ, BC_LSTM_COMPACT="0"), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    want = torch.load(path)
    for H, B, T in cases:
        g = gen(H + B + T)
        torch.manual_seed(H + B + T)
        m = M.ResLSTM(H, num_layers=1).to(DEV)
        img = m.lstm.recurrent_image_for(0, "bf16x3")
        pre = (torch.randn(B, T, 4 * H, generator=g) * 0.7).to(DEV)
        skip = torch.randn(B, T, H, generator=g).to(DEV)
        got = ops.lstm_recurrent_tc(pre, img, skip, "bf16x3", ops.lstm_tc_max_batch(H, "bf16x3")).cpu()
        assert torch.isfinite(got).all() and torch.equal(got, want[(H, B, T)]), (H, B, T)
