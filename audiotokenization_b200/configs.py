"""Model configurations of the BigCodec hot path.

These are the ``codec_encoder`` / ``codec_decoder`` blocks of the reference's
hydra model YAMLs, restated as plain dicts (hydra/omegaconf are not needed and
are not installed).  Keys the reference constructors reject (``type``,
``vq_dim`` -- see SURVEY.md section 5) are already dropped, so each dict can be
splatted into ``BigCodecEncoder(**cfg)`` / ``BigCodecDecoder(**cfg)`` of either
this package or the reference.

Sources (relative to ``/root/reference/BigCodec_SSL/``):
  base          cfgs/config11/model/base.yaml:1-31  (== config8/10/12 base)
  config9_base  cfgs/config9/model/base.yaml:1-31
  debug         config/model/debug.yaml:1-31
  default       config/model/default.yaml:1-32
  debug_causal  cfgs/config5/model/debug.yaml:1-31  (causal encoder)
  debug_nodil   cfgs/config3/model/debug.yaml:1-31  (dilations 1,1,1)
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml

_COMMON_DEC = dict(
    dilations=[1, 3, 9], causal=False, antialias=False, vq_num_quantizers=1,
    vq_commit_weight=0.25, vq_weight_init=False, fsq=False,
    fsq_levels=[4, 4, 4, 8], vq_full_commit_loss=False, codebook_size=8192,
    codebook_dim=8, rnn_bidirectional=False, use_rnn=True,
)

MODEL_CONFIGS: Dict[str, Dict[str, Dict[str, Any]]] = {
    "base": {
        "codec_encoder": dict(out_channels=512, ngf=32, use_rnn=True, rnn_bidirectional=False,
                              rnn_num_layers=2, up_ratios=[2, 4, 5, 5], dilations=[1, 3, 9],
                              causal=False, antialias=False),
        "codec_decoder": dict(_COMMON_DEC, in_channels=512, upsample_initial_channel=512, ngf=32,
                              rnn_num_layers=2, up_ratios=[5, 5, 4, 2]),
    },
    "config9_base": {
        "codec_encoder": dict(out_channels=512, ngf=32, use_rnn=True, rnn_bidirectional=False,
                              rnn_num_layers=2, up_ratios=[4, 4, 4, 5], dilations=[1, 3, 9],
                              causal=False, antialias=False),
        "codec_decoder": dict(_COMMON_DEC, in_channels=512, upsample_initial_channel=512, ngf=32,
                              rnn_num_layers=2, up_ratios=[5, 4, 4, 4]),
    },
    "debug": {
        "codec_encoder": dict(out_channels=512, ngf=16, use_rnn=False, rnn_bidirectional=False,
                              rnn_num_layers=1, up_ratios=[2, 2, 4, 4, 5], dilations=[1, 3, 9],
                              causal=False, antialias=False),
        "codec_decoder": dict(_COMMON_DEC, in_channels=512, upsample_initial_channel=512, ngf=16,
                              rnn_num_layers=1, up_ratios=[5, 4, 4, 2, 2]),
    },
    "default": {
        "codec_encoder": dict(out_channels=1024, ngf=48, use_rnn=True, rnn_bidirectional=False,
                              rnn_num_layers=2, up_ratios=[2, 2, 2, 5, 5], dilations=[1, 3, 9],
                              causal=False, antialias=False),
        "codec_decoder": dict(_COMMON_DEC, in_channels=1024, upsample_initial_channel=1536, ngf=48,
                              rnn_num_layers=2, up_ratios=[5, 5, 2, 2, 2]),
    },
}
MODEL_CONFIGS["debug_causal"] = copy.deepcopy(MODEL_CONFIGS["debug"])
MODEL_CONFIGS["debug_causal"]["codec_encoder"]["causal"] = True
MODEL_CONFIGS["debug_nodil"] = copy.deepcopy(MODEL_CONFIGS["debug"])
MODEL_CONFIGS["debug_nodil"]["codec_encoder"]["dilations"] = [1, 1, 1]
MODEL_CONFIGS["debug_nodil"]["codec_decoder"]["dilations"] = [1, 1, 1]

# A deliberately small model of the same family (same layer types, every
# stride/dilation kind, LSTM on both sides) used by unit tests and golden
# fixtures where base would be too slow for the CPU oracle.
MODEL_CONFIGS["tiny"] = {
    "codec_encoder": dict(out_channels=64, ngf=8, use_rnn=True, rnn_bidirectional=False,
                          rnn_num_layers=2, up_ratios=[2, 4, 5], dilations=[1, 3, 9],
                          causal=False, antialias=False),
    "codec_decoder": dict(_COMMON_DEC, in_channels=64, upsample_initial_channel=64, ngf=8,
                          rnn_num_layers=2, up_ratios=[5, 4, 2], codebook_size=512),
}

# the tiny model with the decoder's other quantizer (fsq=True: vq/codec_decoder.py:41-47); 4*4*4*8 = 512 codes
MODEL_CONFIGS["tiny_fsq"] = copy.deepcopy(MODEL_CONFIGS["tiny"])
MODEL_CONFIGS["tiny_fsq"]["codec_decoder"].update(fsq=True, fsq_levels=[4, 4, 4, 8], codebook_size=512)

SAMPLE_RATE = 16000  # config/dataset/default.yaml (dataset.sample_rate)


def get_config(name: str, antialias: bool | None = None) -> Dict[str, Dict[str, Any]]:
    """Return a deep copy of a named model config, optionally forcing ``antialias``."""
    if name not in MODEL_CONFIGS:
        raise KeyError(f"unknown model config {name!r}; have {sorted(MODEL_CONFIGS)}")
    cfg = copy.deepcopy(MODEL_CONFIGS[name])
    if antialias is not None:
        cfg["codec_encoder"]["antialias"] = bool(antialias)
        cfg["codec_decoder"]["antialias"] = bool(antialias)
    return cfg


def load_model_yaml(path: str) -> Dict[str, Dict[str, Any]]:
    """Read a reference ``model/*.yaml`` (or a saved ``hydra/config.yaml``) with plain yaml.

    Mirrors what ``OmegaConf.load`` + ``construct_model`` consume
    (lightning_module.py:88-139, extract_indices.py:286): only the
    ``codec_encoder`` / ``codec_decoder`` blocks, minus the keys the constructors
    reject.
    """
    with open(path, "r") as f:
        doc = yaml.safe_load(f)
    if "model" in doc and "codec_encoder" in doc["model"]:
        doc = doc["model"]
    enc = dict(doc["codec_encoder"])
    dec = dict(doc["codec_decoder"])
    enc_type = enc.pop("type", "bigcodec")
    if enc_type != "bigcodec":
        raise ValueError(f"only the 'bigcodec' encoder family is on the hot path, got {enc_type!r}")
    dec.pop("type", None)
    dec.pop("vq_dim", None)
    return {"codec_encoder": enc, "codec_decoder": dec}
