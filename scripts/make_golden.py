#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Runs only in the build container (needs /root/reference).  It imports the
reference's own ``vq`` package (with a stub for the absent ``einx`` module,
which the hot path never calls -- SURVEY.md section 8c), builds
BigCodecEncoder / BigCodecDecoder from the model config, loads the seeded
state dicts of ``audiotokenization_b200.synth`` with strict=True (which also
proves key/shape compatibility), and stores inputs-by-seed + outputs.

The fixtures pin (a) the CPU oracle (tests/test_oracle_golden.py) and (b) the
CUDA path (tests/test_gpu_golden.py).  Nothing at test/bench time reads
/root/reference.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REF = "/root/reference/BigCodec_SSL"


def import_reference():
    stub = types.ModuleType("einx")
    stub.get_at = None
    stub.where = None
    sys.modules.setdefault("einx", stub)
    sys.path.insert(0, REF)
    import vq  # noqa: F401  (the reference package)
    from vq import BigCodecEncoder, BigCodecDecoder
    return BigCodecEncoder, BigCodecDecoder


CASES = [
    # name, config, antialias, clip samples, batch, input kind
    ("tiny", "tiny", False, 4000, 2, "tones"),
    ("tiny_aa", "tiny", True, 4000, 2, "tones"),
    ("tiny_ragged", "tiny", False, 3987, 1, "noise"),
    ("base_1s", "base", False, 16000, 1, "tones"),
    ("base_aa_1s", "base", True, 16000, 1, "tones"),
    ("debug_1s", "debug", False, 16000, 1, "noise"),
    ("debug_causal_1s", "debug_causal", False, 15999, 1, "noise"),
    ("debug_nodil_1s", "debug_nodil", False, 16000, 1, "tones"),
    ("config9_base_1s", "config9_base", False, 16000, 1, "tones"),
    ("default_half_s", "default", False, 8000, 1, "tones"),   # original BigCodec: ngf 48, 1024-d, channels 48..1536
    ("tiny_fsq", "tiny_fsq", False, 4000, 2, "tones"),        # the decoder's FSQ quantizer branch (fsq=True)
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--only", default=None, help="comma-separated case names (default: all)")
    args = ap.parse_args()
    only = set(args.only.split(",")) if args.only else None
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    Enc, Dec = import_reference()
    from audiotokenization_b200 import configs, synth

    for name, cfg_name, aa, nsamp, batch, kind in CASES:
        if only is not None and name not in only:
            continue
        cfg = configs.get_config(cfg_name, antialias=aa)
        enc_sd, dec_sd = synth.make_state_dicts(cfg, seed=0)
        x = synth.synth_batch(0, batch, nsamp, kind)
        out = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            enc = Enc(**cfg["codec_encoder"]).eval()
            dec = Dec(**cfg["codec_decoder"]).eval()
            enc.load_state_dict(enc_sd, strict=True)
            dec.load_state_dict(dec_sd, strict=True)
            enc, dec = enc.to(dt), dec.to(dt)
            with torch.no_grad():
                z = enc(x.to(dt))
                z_q, idx, loss = dec(z, vq=True)
                y = dec(z_q, vq=False)
                if cfg["codec_decoder"].get("fsq", False):
                    # FSQ: indices int32 [B,T']; the index -> embedding entry is indices_to_codes (channel-first), and the
                    # analogue of the cosine margin is the distance of the bounded latents to the nearest rounding boundary
                    q = dec.quantizer
                    emb = (q.indices_to_codes(idx).transpose(1, 2) if dt == torch.float32 else
                           q.project_out(q._indices_to_codes(idx).to(dt)))        # the reference's own method is float32-only
                    bounded = q.bound(q.project_in(z.transpose(1, 2)))
                    margin = (0.5 - (bounded - bounded.round()).abs()).amin(dim=-1)
                else:
                    # the int -> embedding entry (vq2emb is channel-last, residual_vq.py:42-48)
                    emb = dec.vq2emb(idx.permute(1, 2, 0))
                    # cosine margins from the reference's own quantizer parameters
                    layer = dec.quantizer.layers[0]
                    z_e = layer.in_proj(z.transpose(1, 2))
                    e = torch.nn.functional.normalize(z_e.reshape(-1, z_e.shape[-1]))
                    c = torch.nn.functional.normalize(layer.codebook.weight)
                    top2 = (e @ c.t()).topk(2, dim=1).values
                    margin = (top2[:, 0] - top2[:, 1]).view(z.shape[0], -1)
            out[f"z_{tag}"] = z.to(torch.float32).numpy()
            out[f"zq_{tag}"] = z_q.to(torch.float32).numpy()
            out[f"idx_{tag}"] = idx.numpy().astype(np.int32)
            out[f"y_{tag}"] = y.to(torch.float32).numpy()
            out[f"margin_{tag}"] = margin.to(torch.float32).numpy()
            out[f"emb_{tag}"] = emb.to(torch.float32).numpy()
            assert float(loss.abs().sum()) == 0.0
        out["meta"] = np.array([cfg_name, str(int(aa)), str(nsamp), str(batch), kind, "0"])
        path = os.path.join(args.out, name + ".npz")
        np.savez_compressed(path, **out)
        agree = float((out["idx_f32"] == out["idx_f64"]).mean())
        rel = float(np.linalg.norm(out["z_f32"] - out["z_f64"]) / np.linalg.norm(out["z_f64"]))
        print(f"{name:18s} z{out['z_f32'].shape} y{out['y_f32'].shape} idx32==idx64 {agree:.4f} "
              f"z rel(f32 vs f64) {rel:.2e} distinct codes {len(np.unique(out['idx_f32']))} "
              f"min margin {out['margin_f64'].min():.2e}  {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
