"""CPU: long-form chunk planning, receptive-field bookkeeping and the ordered feature hand-off (world size 2, gloo)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch
import torch.nn.functional as F

from conftest import REPO
from audiotokenization_b200 import configs, longform
from audiotokenization_b200.vq import BigCodecEncoder


def _toy_front(seed=0):
    """conv k7 -> dilated conv k7 d3 -> strided conv k4 s2 -> dilated conv k7 d9 -> strided conv k10 s5 (hop 10)."""
    g = torch.Generator().manual_seed(seed)
    spec = [(7, 1, 1, 3, 1, 4), (7, 1, 3, 9, 4, 4), (4, 2, 1, 1, 4, 6), (7, 1, 9, 27, 6, 6), (10, 5, 1, 3, 6, 5)]
    ws = [torch.randn(co, ci, k, generator=g, dtype=torch.float64) * 0.3 for k, s, d, p, ci, co in spec]

    def front(x_cl):                       # [B, T, 1] -> [B, T/10, 5]
        h = x_cl.permute(0, 2, 1).double()
        for (k, s, d, p, ci, co), w in zip(spec, ws):
            h = F.conv1d(h, w, stride=s, dilation=d, padding=p)
        return h.permute(0, 2, 1)
    chain = [(k, s, d, p) for k, s, d, p, ci, co in spec]
    return front, chain, 10


def test_plan_chunks_covers_every_frame_once_and_clips_halo():
    plan = longform.plan_chunks(1000, 300, 13)
    assert [(a, b) for a, b, _, _ in plan] == [(0, 300), (300, 600), (600, 900), (900, 1000)]
    assert plan[0][2:] == (0, 313) and plan[1][2:] == (287, 613) and plan[3][2:] == (887, 1000)
    assert longform.plan_chunks(5, 300, 13) == [(0, 5, 0, 5)]
    with pytest.raises(ValueError):
        longform.plan_chunks(0, 10, 1)


def test_front_context_matches_the_true_receptive_field():
    front, chain, hop = _toy_front()
    left, right = longform.front_context_samples(chain)
    T = 4000
    x = torch.zeros(1, T, 1, dtype=torch.float64)
    base = front(x)
    f = 200                                                   # frame 200 reads samples [2000 - left, 2000 + right]
    for off, inside in ((-left, True), (-left - 1, False), (right, True), (right + 1, False)):
        y = x.clone()
        y[0, f * hop + off, 0] = 1.0
        changed = bool((front(y)[0, f] - base[0, f]).abs().max() > 0)
        assert changed == inside, (off, inside)


@pytest.mark.parametrize("T_frames,chunk,mb", [(1000, 300, 2), (601, 200, 8), (120, 500, 1)])
def test_chunked_front_equals_unchunked(T_frames, chunk, mb):
    front, chain, hop = _toy_front(1)
    left, right = longform.front_context_samples(chain)
    halo = (max(left, right) + hop - 1) // hop + 1
    x = torch.randn(T_frames * hop, dtype=torch.float64, generator=torch.Generator().manual_seed(T_frames))
    whole = front(x.view(1, -1, 1))
    parts = longform.chunked_front(front, x, hop, chunk, halo, micro_batch=mb)
    got = longform.stitch(parts)
    assert got.shape == whole.shape and float((got - whole).abs().max()) <= 1e-12
    short = longform.stitch(longform.chunked_front(front, x, hop, chunk, halo - 2, micro_batch=mb))
    if T_frames > chunk:
        assert float((short - whole).abs().max()) > 1e-9      # a halo below the receptive field IS visible
    with pytest.raises(ValueError):
        longform.chunked_front(front, x[:-1], hop, chunk, halo)


def test_halo_of_the_real_encoder_configs():
    for name, want in (("base", 13), ("tiny", 14), ("debug", 14)):
        enc = BigCodecEncoder(**configs.get_config(name)["codec_encoder"])
        assert longform.halo_frames(enc) == want               # >= 13 frames for base
        left, right = longform.front_context_samples(longform.encoder_front_chain(enc))
        assert max(left, right) <= (want - 1) * int(enc.hop_length)


def test_feature_handoff_world_size_2_gloo(tmp_path):
    script = textwrap.dedent("""
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, %r)
        from audiotokenization_b200 import longform
        from audiotokenization_b200.sharding import shard_range
        dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
        rank = dist.get_rank()
        plan = longform.plan_chunks(1000, 150, 13)                      # 7 chunks: rank 0 owns 4, rank 1 owns 3
        c0, c1 = shard_range(len(plan), rank, 2)
        parts = [(i, torch.arange(plan[i][0], plan[i][1], dtype=torch.float32).view(-1, 1).repeat(1, 3) + 0.5 * i) for i in range(c0, c1)]
        out = longform.gather_features(parts, len(plan), lambda i: plan[i][1] - plan[i][0], 3, "cpu")
        if rank == 0:
            want = torch.cat([torch.arange(a, b, dtype=torch.float32).view(-1, 1).repeat(1, 3) + 0.5 * i for i, (a, b, _, _) in enumerate(plan)])
            assert out.shape == (1, 1000, 3) and torch.equal(out[0], want)
            print("OK")
        else:
            assert out is None
        dist.destroy_process_group()
    """ % REPO)
    path = tmp_path / "w.py"
    path.write_text(script)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(path)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0]
