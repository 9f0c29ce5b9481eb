"""CPU: host-side logic -- configs, state-dict compatibility, weight packing, sharding."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import REPO
from audiotokenization_b200 import configs, sharding, synth
from audiotokenization_b200.vq import BigCodecDecoder, BigCodecEncoder
from audiotokenization_b200.vq import module as M
from oracle import bigcodec_oracle as oracle


@pytest.mark.parametrize("name", ["tiny", "base", "debug", "debug_causal", "config9_base", "default", "tiny_fsq"])
@pytest.mark.parametrize("aa", [False, True])
def test_state_dict_keys_and_shapes_match_reference_layout(name, aa):
    cfg = configs.get_config(name, antialias=aa)
    esd, dsd = synth.make_state_dicts(cfg)
    enc = BigCodecEncoder(**cfg["codec_encoder"])
    dec = BigCodecDecoder(**cfg["codec_decoder"])
    enc.load_state_dict(esd, strict=True)
    dec.load_state_dict(dsd, strict=True)
    assert list(enc.state_dict()) == list(esd) and list(dec.state_dict()) == list(dsd)
    if name == "base":  # SURVEY.md Appendix C tensor counts
        assert (len(esd), len(dsd)) == ((214, 221) if aa else (156, 163))
        assert tuple(dsd["model.2.block.1.weight_g"].shape) == (512, 1, 1)      # ConvT: per INPUT channel
        assert tuple(dsd["quantizer.layers.0._codebook.weight"].shape) == (8192, 8)
    assert int(enc.hop_length) == int(np.prod(cfg["codec_encoder"]["up_ratios"]))


def test_yaml_reader_drops_rejected_keys(tmp_path):
    doc = textwrap.dedent("""
        codec_encoder: {type: bigcodec, out_channels: 64, ngf: 8, use_rnn: True, rnn_bidirectional: False,
                        rnn_num_layers: 2, up_ratios: [2, 4, 5], dilations: [1, 3, 9], causal: False, antialias: False}
        codec_decoder: {in_channels: 64, upsample_initial_channel: 64, ngf: 8, use_rnn: True, rnn_bidirectional: False,
                        rnn_num_layers: 2, up_ratios: [5, 4, 2], dilations: [1, 3, 9], causal: False, antialias: False,
                        vq_num_quantizers: 1, vq_dim: 64, vq_commit_weight: 0.25, vq_weight_init: False, fsq: False,
                        fsq_levels: [4, 4, 4, 8], vq_full_commit_loss: False, codebook_size: 512, codebook_dim: 8}
        mpd: {periods: [2, 3]}
    """)
    p = tmp_path / "m.yaml"
    p.write_text(doc)
    cfg = configs.load_model_yaml(str(p))
    assert "type" not in cfg["codec_encoder"] and "vq_dim" not in cfg["codec_decoder"]
    BigCodecEncoder(**cfg["codec_encoder"])
    BigCodecDecoder(**cfg["codec_decoder"])
    bad = tmp_path / "b.yaml"
    bad.write_text(doc.replace("type: bigcodec", "type: conformer_stft"))
    with pytest.raises(ValueError):
        configs.load_model_yaml(str(bad))


def test_conv_weight_packing_is_the_folded_reference_weight():
    conv = M.WNConv1d(6, 10, kernel_size=7, dilation=3, padding=9)
    conv.weight_g.data.mul_(1.7)
    conv.bias.data.normal_()
    w, b = conv.packed()
    ref = oracle.fold_weight_norm(conv.weight_g.double(), conv.weight_v.double())
    assert w.shape == (7, 6, 10)
    assert torch.allclose(w.permute(2, 1, 0).double(), ref, atol=1e-7)
    assert conv.out_length(100) == 100
    # cache invalidates on in-place parameter change
    conv.weight_g.data.mul_(2.0)
    conv.weight_g._version  # noqa: B018
    conv.load_state_dict(conv.state_dict())
    w2, _ = conv.packed()
    assert torch.allclose(w2, 2 * w, atol=1e-6)


@pytest.mark.parametrize("stride,causal", [(2, False), (4, False), (5, False), (3, True)])
def test_transposed_conv_phase_packing(stride, causal):
    cin, cout, T = 5, 4, 9
    if causal:
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, causal=True)
        p, conv = 0, m.conv
    else:
        p = stride // 2 + stride % 2
        m = M.WNConvTranspose1d(cin, cout, 2 * stride, stride=stride, padding=p, output_padding=stride % 2)
        conv = m
    w_ph, b = conv.packed()
    assert w_ph.shape == (stride, 2, cin, cout)
    x = torch.randn(2, cin, T, dtype=torch.float64)
    w = oracle.fold_weight_norm(conv.weight_g.double(), conv.weight_v.double())
    if causal:
        ref = F.conv_transpose1d(x, w, conv.bias.double(), stride=stride)[..., :-stride]
    else:
        ref = F.conv_transpose1d(x, w, conv.bias.double(), stride=stride, padding=p, output_padding=stride % 2)
    # emulate bc_convtr1d_fwd's phase loop on the CPU with the packed weights
    xc = x.transpose(1, 2)
    y = torch.zeros(2, T * stride, cout, dtype=torch.float64)
    for ph in range(stride):
        q = (ph + p) // stride
        for t in range(T):
            acc = b.double().expand(2, -1).clone()
            for k in range(2):
                g = t + k - (1 - q)
                if 0 <= g < T:
                    acc += xc[:, g] @ w_ph[ph, k].double()
            y[:, t * stride + ph] = acc
    assert torch.allclose(y.transpose(1, 2), ref, atol=1e-6)


def test_unsupported_configurations_are_errors():
    with pytest.raises(NotImplementedError):
        M.ResLSTM(32, bidirectional=True)
    cfg = configs.get_config("tiny")["codec_decoder"]
    with pytest.raises(AssertionError):          # vq/codec_decoder.py:47: codebook_size must equal prod(fsq_levels)
        BigCodecDecoder(**dict(cfg, fsq=True, codebook_size=128))
    from audiotokenization_b200.vq import FSQ
    with pytest.raises(NotImplementedError):
        FSQ([4, 4], num_codebooks=2)
    q = BigCodecDecoder(**dict(cfg, fsq=True, fsq_levels=[8, 5, 5], codebook_size=200)).quantizer
    assert q.codebook_size == 200 and tuple(q.implicit_codebook.shape) == (200, 3)
    assert q._basis.tolist() == [1, 8, 40] and sorted(k for k, _ in q.named_parameters()) == [
        "project_in.bias", "project_in.weight", "project_out.bias", "project_out.weight"]
    with pytest.raises(ValueError):
        M.set_precision("fp8")


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    owned = sharding.shard_by_cost([10, 1, 1, 1, 9, 2, 8], 3)
    assert sorted(i for o in owned for i in o) == list(range(7))
    loads = [sum([10, 1, 1, 1, 9, 2, 8][i] for i in o) for o in owned]
    assert max(loads) <= 12


def test_two_rank_gloo_shard_and_host_gather(tmp_path):
    """world_size-2 CPU run of the multi-GPU host logic: each rank takes its shard of a
    deterministic per-clip 'index' table and rank 0 reassembles it in utterance order."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {REPO!r})
        import numpy as np, torch, torch.distributed as dist
        from audiotokenization_b200 import sharding
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        n, tp = 11, 7
        full = (np.arange(n * tp * 1).reshape(n, tp, 1) % 8192).astype(np.int16)
        s, e = sharding.shard_range(n, r, w)
        out = sharding.gather_indices_to_rank0(full[s:e])
        if r == 0:
            assert out.dtype == np.int16 and np.array_equal(out, full), "gather mismatch"
            print("GATHER_OK")
        else:
            assert out is None
        dist.barrier(); dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GATHER_OK" in r.stdout


def test_synthetic_inputs_are_deterministic_and_bounded():
    a = synth.synth_batch(3, 2, 1600)
    b = synth.synth_batch(3, 2, 1600)
    assert torch.equal(a, b) and a.abs().max() <= 1.0 and a.std() > 0.1
    c = synth.fast_synth_batch(5, 4, 1600)
    assert torch.equal(c, synth.fast_synth_batch(5, 4, 1600)) and c.abs().max() <= 1.0
    assert not torch.equal(c[0], c[1])


def test_bench_reference_arm_prints_exactly_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): ONE stdout line, the contract's keys, the CPU
    oracle port as the implementation; rank > 0 under torchrun prints nothing."""
    import json
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-sample-clips", "1", "--clip-seconds", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_stream_weight_image_layout():
    """pack_stream_weight: [n_tile][16-ch group][tap][hi|lo][2 k-planes][N][8] with hi + lo == the fp32 weight to 2^-16."""
    from audiotokenization_b200 import ops
    K, cin, cout, N = 3, 64, 512, 256
    w = torch.randn(K, cin, cout, generator=torch.Generator().manual_seed(5))
    img = ops.pack_stream_weight(w, N, "bf16x3")
    assert tuple(img.shape) == (cout // N, cin // 16, K, 2, 2, N, 8) and img.dtype == torch.bfloat16
    nt, g, k, h, n, e = 1, 2, 1, 1, 77, 5
    ci, co = g * 16 + h * 8 + e, nt * N + n
    hi, lo = img[nt, g, k, 0, h, n, e].float(), img[nt, g, k, 1, h, n, e].float()
    assert hi == w[k, ci, co].to(torch.bfloat16).float()
    assert abs(float(hi + lo - w[k, ci, co])) <= 2.0 ** -16 * abs(float(w[k, ci, co])) + 1e-12
    assert tuple(ops.pack_stream_weight(w, N, "bf16").shape) == (cout // N, cin // 16, K, 1, 2, N, 8)


class _FakeEncoder:
    """Stand-in for BigCodecEncoder with the two-stage front-end interface, pure CPU tensors: shallow = x * 2 over
    [b, T, 1] -> [b, T // 2, 4], deep = sum over channels -> [b, T // 2, 1].  Records the batch sizes each stage saw."""

    def __init__(self):
        self.shallow_batches, self.deep_batches, self.whole_batches = [], [], []

    def front_shallow_cl(self, x_cl, out=None):
        self.shallow_batches.append(x_cl.shape[0])
        y = (x_cl[:, ::2, :] * 2.0).expand(-1, -1, 4).contiguous()
        if out is None:
            return y
        assert out.shape == y.shape and out.is_contiguous()
        out.copy_(y)
        return out

    def front_deep_cl(self, h):
        self.deep_batches.append(h.shape[0])
        return h.sum(dim=2, keepdim=True)

    def front_cl(self, x_cl):
        self.whole_batches.append(x_cl.shape[0])
        return self.front_deep_cl(self.front_shallow_cl(x_cl))


@pytest.mark.parametrize("sizes,deep,expect_deep", [
    ([2, 2, 2, 2], 4, [4, 4]),          # buffer fills exactly
    ([2, 2, 2], 5, [4, 2]),             # 2+2 | 2: a push that does not fit flushes first
    ([3, 3, 1], 64, [7]),               # never full: one deep launch at take()
    ([1], 2, [1]),
])
def test_front_pipeline_gathers_micro_batches_in_order(sizes, deep, expect_deep):
    """model._FrontPipeline: shallow stage per micro-batch into slices of one hand-off buffer, deep stage when it is full
    or at take(); the concatenated result keeps push order and equals the un-staged front end."""
    from audiotokenization_b200.model import _FrontPipeline
    enc = _FakeEncoder()
    pipe = _FrontPipeline(enc, deep)
    T = 10
    xs = [torch.arange(b * T, dtype=torch.float32).view(b, T, 1) + 100 * i for i, b in enumerate(sizes)]
    for x in xs:
        pipe.push(x)
    got = pipe.take()
    want = torch.cat([_FakeEncoder().front_cl(x) for x in xs], dim=0)
    assert torch.equal(got, want)
    assert enc.shallow_batches == sizes and enc.deep_batches == expect_deep and enc.whole_batches == []
    # the pipeline is reusable after take(): second group, same buffer
    pipe.push(xs[0])
    assert torch.equal(pipe.take(), want[: sizes[0]])


def test_front_pipeline_without_gathering_uses_the_plain_front_end():
    from audiotokenization_b200.model import _FrontPipeline
    enc = _FakeEncoder()
    pipe = _FrontPipeline(enc, 2)          # deep_batch <= micro-batch: nothing to gather
    x = torch.ones(2, 6, 1)
    pipe.push(x)
    pipe.push(x)
    assert pipe.take().shape == (4, 3, 1) and enc.whole_batches == [2, 2]


def test_front_pipeline_flushes_when_the_clip_length_changes():
    """Equal-length groups of different lengths may share one pipeline: the hand-off buffer is re-made, results of a
    take() can then only be returned per length (features of different lengths do not concatenate)."""
    from audiotokenization_b200.model import _FrontPipeline
    enc = _FakeEncoder()
    pipe = _FrontPipeline(enc, 8)
    a, b = torch.ones(2, 10, 1), torch.ones(3, 6, 1)
    pipe.push(a)
    fa = pipe.take()
    pipe.push(b)                       # different T after a take(): buffer must be re-allocated, not reused
    fb = pipe.take()
    assert fa.shape == (2, 5, 1) and fb.shape == (3, 3, 1)
    assert enc.deep_batches == [2, 3]


@pytest.mark.parametrize("fmt", ["state_dict", "model", "bare", "codecenc", "module_prefixed"])
def test_from_checkpoint_accepts_the_reference_checkpoint_formats(tmp_path, fmt):
    """extract_indices.py:309-323: ``{'state_dict'}`` (Lightning: encoder.* / decoder.* next to discriminator
    entries), ``{'model'}``, a bare state dict; the wrappers' ``model.CodecEnc`` / ``model.generator`` names."""
    import yaml
    from audiotokenization_b200.model import BigCodecModel
    cfg = configs.get_config("tiny")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=3)
    pe, pd = {"codecenc": ("model.CodecEnc.", "model.generator."), "module_prefixed": ("module.encoder.", "module.decoder.")}.get(
        fmt, ("encoder.", "decoder."))
    sd = {pe + k: v for k, v in enc_sd.items()}
    sd.update({pd + k: v for k, v in dec_sd.items()})
    sd["discriminator.mpd.0.weight"] = torch.zeros(3)                       # other modules of the Lightning checkpoint
    ckpt = {"state_dict": sd, "epoch": 3} if fmt in ("state_dict", "codecenc", "module_prefixed") else ({"model": sd} if fmt == "model" else sd)
    torch.save(ckpt, tmp_path / "last.ckpt")
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump({"model": {"codec_encoder": dict(cfg["codec_encoder"], type="bigcodec"),
                                  "codec_decoder": dict(cfg["codec_decoder"], vq_dim=8)}}, f)
    m = BigCodecModel.from_checkpoint(str(tmp_path / "last.ckpt"), str(tmp_path / "config.yaml"), device="cpu", precision="bf16x3")
    for k, v in enc_sd.items():
        assert torch.equal(m.encoder.state_dict()[k], v), k
    for k, v in dec_sd.items():
        assert torch.equal(m.decoder.state_dict()[k], v), k
    assert m.precision == "bf16x3" and m.encoder.precision == "bf16x3" and m.decoder.precision == "bf16x3"


def test_from_checkpoint_non_strict_fallback_and_refusal(tmp_path, capsys):
    import yaml
    from audiotokenization_b200.model import BigCodecModel
    cfg = configs.get_config("tiny")
    enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=4)
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump({"codec_encoder": cfg["codec_encoder"], "codec_decoder": cfg["codec_decoder"]}, f)
    # one tensor missing, one unexpected: strict fails, the reference's non-strict fallback loads the rest
    sd = {"encoder." + k: v for k, v in enc_sd.items()}
    sd.update({"decoder." + k: v for k, v in dec_sd.items()})
    dropped = next(k for k in sd if k.endswith("bias"))
    del sd[dropped]
    sd["decoder.extra.weight"] = torch.zeros(2)
    torch.save({"state_dict": sd}, tmp_path / "partial.ckpt")
    m = BigCodecModel.from_checkpoint(str(tmp_path / "partial.ckpt"), str(tmp_path / "config.yaml"), device="cpu")
    out = capsys.readouterr().out
    assert "Strict state_dict loading failed" in out and "1 missing" in out
    k0 = next(k for k in enc_sd if k.endswith("weight_v"))
    assert torch.equal(m.encoder.state_dict()[k0], enc_sd[k0])
    # nothing under a known prefix: an error, never a silently random-initialised model
    torch.save({"state_dict": {"foo." + k: v for k, v in enc_sd.items()}}, tmp_path / "alien.ckpt")
    with pytest.raises(ValueError, match="no encoder/decoder entries"):
        BigCodecModel.from_checkpoint(str(tmp_path / "alien.ckpt"), str(tmp_path / "config.yaml"), device="cpu")
    # right prefixes, wrong architecture: non-strict loading matches nothing -> refuse
    big = configs.get_config("debug")
    e2, d2 = synth.make_state_dicts(big, seed=1)
    sd2 = {"encoder." + k: v + 0 for k, v in e2.items() if v.dim() == 3}
    sd2.update({"decoder." + k: v for k, v in d2.items() if v.dim() == 3})
    torch.save(sd2, tmp_path / "wrong.ckpt")
    with pytest.raises(RuntimeError, match="matched no"):
        BigCodecModel.from_checkpoint(str(tmp_path / "wrong.ckpt"), str(tmp_path / "config.yaml"), device="cpu")


def test_precision_lives_on_the_modules_not_only_in_the_process_global():
    """``BigCodecModel(precision='bf16x3').encoder(x)`` (the reference's ``lm.model['CodecEnc'](x)`` call
    pattern) must run in the model's mode.  Checked on CPU through the mode the leaf ops would see."""
    from audiotokenization_b200.model import BigCodecModel
    cfg = configs.get_config("tiny")
    m = BigCodecModel(cfg, device="cpu", precision="bf16x3")
    assert M.get_precision() == "fp32"
    seen = []
    with M.module_scope(m.encoder):
        seen.append(M.get_precision())
    assert seen == ["bf16x3"] and M.get_precision() == "fp32"
    m.precision = "bf16"
    with M.module_scope(m.decoder):
        assert M.get_precision() == "bf16"
    plain = BigCodecEncoder(**cfg["codec_encoder"])
    with M.precision_scope("bf16x3"), M.module_scope(plain):                # no own mode: the enclosing scope's
        assert M.get_precision() == "bf16x3"
    with pytest.raises(ValueError):
        m.precision = "fp8"


def test_stream_pair_weight_image_layout():
    """CTA-pair image: [n_tile block][rank][16-channel group][tap][split][k-plane][n_tile/2 rows][8]; rank r holds rows
    [r*n_tile/2, (r+1)*n_tile/2) of every k-plane; hi + lo reproduces the weight to bf16x3 accuracy."""
    from audiotokenization_b200 import ops
    g = torch.Generator().manual_seed(3)
    K, cin, cout, nt = 3, 32, 256, 128
    w = torch.randn(K, cin, cout, generator=g)
    img = ops.pack_stream_weight_pair(w, nt, "bf16x3")
    assert tuple(img.shape) == (cout // nt, 2, cin // 16, K, 2, 2, nt // 2, 8) and img.dtype == torch.bfloat16
    single = ops.pack_stream_weight(w, nt, "bf16x3")            # [nt][g][k][split][2][N][8]
    for r in range(2):
        assert torch.equal(img[:, r], single[..., r * (nt // 2):(r + 1) * (nt // 2), :])
    # element check: image[nb, r, grp, k, sp, h, n, e] = part_sp(w[k, grp*16 + h*8 + e, nb*nt + r*nt/2 + n])
    hi = w.to(torch.bfloat16)
    for (nb, r, grp, k, h, n, e) in [(0, 0, 0, 0, 0, 0, 0), (1, 1, 1, 2, 1, 63, 7), (0, 1, 1, 1, 0, 5, 3)]:
        assert img[nb, r, grp, k, 0, h, n, e] == hi[k, grp * 16 + h * 8 + e, nb * nt + r * (nt // 2) + n]
    rec = (img[:, :, :, :, 0].float() + img[:, :, :, :, 1].float())
    ref = w.reshape(K, cin // 16, 2, 8, cout // nt, 2, nt // 2).permute(4, 5, 1, 0, 2, 6, 3)
    assert float((rec - ref).abs().max()) <= 2e-5 * float(w.abs().max())
