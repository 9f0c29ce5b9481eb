"""Driver-level wrapper: the ``BigCodecModel`` of extract_indices.py / inference_full.py.

``forward`` keeps the reference wrappers' call pattern and result dicts
(extract_indices.py:347-371, inference_full.py:557-561) with the semantic mapping
``lm.model['CodecEnc'] == encoder`` and ``lm.model['generator'] == decoder``
(SURVEY.md section 0).  ``extract_indices`` is the batched host-buffer entry point the
benchmark's end-to-end number goes through: pinned host waveforms in, int16 ``(T', n_q)``
index arrays out (the on-disk layout of extract_indices.py:520-532).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import _cabi, configs, ops
from .sharding import shard_range as sharding_range
from .vq import BigCodecDecoder, BigCodecEncoder, precision_scope


def _strip_prefix(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


class _FrontPipeline:
    """Two-stage schedule of the encoder's conv stack over equal-length utterances.

    ``push`` runs the shallow stage on one micro-batch, writing its result straight into a slice of a
    [deep_batch, T_mid, C_mid] hand-off buffer; whenever the buffer is full (or on ``take``) the deep stage runs
    over all of it, so that its few tiles per utterance still fill the last round of the persistent kernels."""

    def __init__(self, encoder, deep_batch: int):
        self.enc, self.cap = encoder, int(deep_batch)
        self.buf, self.fill, self.feats, self.t_in = None, 0, [], None

    def push(self, x_cl: torch.Tensor) -> None:
        b = x_cl.shape[0]
        if self.cap <= b:                                   # nothing to gather
            self.feats.append(self.enc.front_cl(x_cl))
            return
        if self.t_in is not None and x_cl.shape[1] != self.t_in:   # another clip length: new hand-off geometry
            self.flush()
            self.buf = None
        if self.fill + b > self.cap:
            self.flush()
        self.t_in = x_cl.shape[1]
        if self.buf is None:
            y = self.enc.front_shallow_cl(x_cl)
            self.buf = torch.empty((self.cap,) + tuple(y.shape[1:]), device=y.device, dtype=y.dtype)
            self.buf[:b].copy_(y)
        else:
            self.enc.front_shallow_cl(x_cl, out=self.buf[self.fill:self.fill + b])
        self.fill += b

    def flush(self) -> None:
        if self.fill:
            self.feats.append(self.enc.front_deep_cl(self.buf[:self.fill]))
            self.fill = 0

    def take(self) -> torch.Tensor:
        """Frame-rate features of everything pushed since the last ``take``, in push order."""
        self.flush()
        feat = self.feats[0] if len(self.feats) == 1 else torch.cat(self.feats, dim=0)
        self.feats = []
        return feat


class BigCodecModel(nn.Module):
    def __init__(self, cfg: dict, enc_state: Optional[dict] = None, dec_state: Optional[dict] = None,
                 device: str = "cuda", precision: str = "fp32"):
        super().__init__()
        _cabi.load_library()  # fail loudly before anything else if the CUDA library is missing
        self.cfg = cfg
        self.encoder = BigCodecEncoder(**cfg["codec_encoder"])
        self.decoder = BigCodecDecoder(**cfg["codec_decoder"])
        if enc_state is not None:
            self.encoder.load_state_dict(enc_state, strict=True)
        if dec_state is not None:
            self.decoder.load_state_dict(dec_state, strict=True)
        self.codebook_size = cfg["codec_decoder"]["codebook_size"]
        self.precision = precision      # property: also stored on encoder / decoder (direct sub-module calls honour it)
        self.to(device)
        self.eval()

    @property
    def precision(self) -> str:
        return self._precision

    @precision.setter
    def precision(self, mode: str) -> None:
        if mode not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}")
        self._precision = mode
        self.encoder.precision = mode
        self.decoder.precision = mode

    # -- checkpoint adapter (extract_indices.py:309-323) ---------------------------------
    # prefixes under which checkpoints carry the codec's two modules: the Lightning module's attribute names
    # (SURVEY.md section 5), the wrappers' ``lm.model[...]`` names, both optionally DataParallel-wrapped
    _CKPT_PREFIXES = (("encoder.", "decoder."), ("model.CodecEnc.", "model.generator."), ("CodecEnc.", "generator."))

    @classmethod
    def split_checkpoint(cls, ckpt) -> tuple:
        """(encoder state dict, decoder state dict, number of other keys) from a loaded checkpoint object:
        ``{'state_dict': ...}`` (Lightning), ``{'model': ...}`` or a bare state dict (extract_indices.py:309-315)."""
        sd = ckpt
        if isinstance(ckpt, dict):
            if "state_dict" in ckpt:
                sd = ckpt["state_dict"]
            elif "model" in ckpt and isinstance(ckpt["model"], dict):
                sd = ckpt["model"]
        if not isinstance(sd, dict) or not sd:
            raise ValueError("checkpoint holds no state dict")
        for lead in ("", "module.", "lm.", "lm.module."):
            for pe, pd in cls._CKPT_PREFIXES:
                enc, dec = _strip_prefix(sd, lead + pe), _strip_prefix(sd, lead + pd)
                if enc and dec:
                    return enc, dec, len(sd) - len(enc) - len(dec)
        raise ValueError("checkpoint has no encoder/decoder entries under any known prefix "
                         f"({', '.join(a + '|' + b for a, b in cls._CKPT_PREFIXES)}); first keys: {list(sd)[:4]}")

    @classmethod
    def from_checkpoint(cls, ckpt_path: str, config_path: str, device: str = "cuda", precision: str = "fp32"):
        """Checkpoint adapter of extract_indices.py:309-323: strict load first, then the reference's non-strict
        fallback -- which here still has to match at least one tensor per module (a model that silently keeps its
        random initialisation is never what the caller asked for)."""
        cfg = configs.load_model_yaml(config_path)
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        enc, dec, other = cls.split_checkpoint(ckpt)
        m = cls(cfg, device=device, precision=precision)
        try:
            m.encoder.load_state_dict(enc, strict=True)
            m.decoder.load_state_dict(dec, strict=True)
            print(f"State dict loaded strictly ({len(enc)} encoder + {len(dec)} decoder tensors"
                  + (f", {other} entries of other modules ignored" if other else "") + ").")
        except RuntimeError as e:  # the reference falls back to non-strict loading (extract_indices.py:318-323)
            print(f"Strict state_dict loading failed: {e}. Attempting non-strict loading.")
            for name, mod, part in (("encoder", m.encoder, enc), ("decoder", m.decoder, dec)):
                own = mod.state_dict()
                usable = {k: v for k, v in part.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
                if not usable:
                    raise RuntimeError(f"non-strict loading matched no {name} tensor: wrong config for this checkpoint?") from e
                res = mod.load_state_dict(usable, strict=False)
                skipped = sorted(set(part) - set(usable))
                print(f"  {name}: loaded {len(usable)} tensors, {len(res.missing_keys)} missing (keep their initial values), "
                      f"{len(skipped)} unexpected / mis-shaped skipped")
        return m

    # -- reference call pattern ---------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor, round_trip: bool = False):
        """x [B,1,T] on the GPU.  ``{'indices'}`` (extract_indices) or, with ``round_trip``,
        ``{'x_rec','indices','loss'}`` (inference_full)."""
        with precision_scope(self.precision):
            vq_emb = self.encoder(x)
            vq_post_emb, vq_code, _ = self.decoder(vq_emb, vq=True)
            if not round_trip:
                return {"indices": vq_code}
            recon = self.decoder(vq_post_emb, vq=False)
        return {"x_rec": recon, "indices": vq_code, "loss": {}}

    @torch.no_grad()
    def inference(self, wav: torch.Tensor) -> torch.Tensor:
        """CodecLightningModule.inference (lightning_module.py:280-285): wav [B,T] -> recon [B,T]."""
        return self.forward(wav.unsqueeze(1), round_trip=True)["x_rec"].squeeze(1)

    # -- device-resident fast path ---------------------------------------------------------------
    @torch.no_grad()
    def encode_indices_cl(self, x_cl: torch.Tensor, want_margin: bool = False):
        """x_cl [B,T,1] device -> (idx int32 [n_q,B,T'], margin [n_q,B,T'] | None, z_cl [B,T',C])."""
        with precision_scope(self.precision):
            z_cl = self.encoder.forward_cl(x_cl)
            _, idx, margin = self.decoder.quantizer.forward_cl(z_cl, want_margin=want_margin)
        return idx, margin, z_cl

    def _indices_from_features(self, feat):
        """frame-rate features [B,T',enc_dim] -> int16 [B,T',n_q] on the device."""
        z_cl = self.encoder.back_cl(feat)
        _, idx, _ = self.decoder.quantizer.forward_cl(z_cl, want_zq=False)    # index-only: no z_q is written
        n_q, B, Tp = idx.shape
        return ops.indices_to_int16(idx.reshape(n_q, B * Tp)).view(B, Tp, n_q)

    @torch.no_grad()
    def indices_device(self, x_dev: torch.Tensor, micro_batch: int = 8, rnn_batch: int = 512,
                       deep_batch: int = 64) -> torch.Tensor:
        """Device waveforms [N,1,T] -> int16 [N,T',n_q] on the device.

        Two-stage schedule: the convolutional front end runs in micro-batches (bounded activation
        memory: the stem's [mb, T, ngf] tensor is the largest), its frame-rate output (2 KB per frame) is
        collected for up to ``rnn_batch`` utterances (512: four 128-row tiles, two per CTA of the tensor-core LSTM kernel), and the sequential LSTM + final conv + VQ then run
        once over that whole group.  Inside the front end the last strided stages run over ``deep_batch``
        utterances at a time (`_FrontPipeline`)."""
        outs = []
        N = x_dev.shape[0]
        with precision_scope(self.precision):
            pipe = _FrontPipeline(self.encoder, deep_batch)
            for c0 in range(0, N, rnn_batch):
                c1 = min(N, c0 + rnn_batch)
                for b0 in range(c0, c1, micro_batch):
                    xb = x_dev[b0:min(c1, b0 + micro_batch)]
                    pipe.push(xb.reshape(xb.shape[0], xb.shape[2], 1))
                outs.append(self._indices_from_features(pipe.take()))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    @torch.no_grad()
    def indices_longform(self, x_dev: torch.Tensor, chunk_seconds: float = 60.0, micro_batch: int = 4, group=None,
                         halo: Optional[int] = None) -> Optional[torch.Tensor]:
        """One long recording [T] (device, T a multiple of the hop) -> int16 [1, T', n_q] on the device.

        The convolutional front end runs over hop-aligned chunks with a receptive-field halo (batched; dealt out over
        the ranks of ``group`` when torch.distributed is initialised), the LSTM + final conv + VQ once over the
        stitched frame-rate features on rank 0 (other ranks return None).  Bit-identical to encoding the recording
        in one piece (BASELINE.json configs[3])."""
        from . import longform
        import torch.distributed as dist
        x = x_dev.reshape(-1)
        hop = int(self.encoder.hop_length)
        if halo is None:
            halo = longform.halo_frames(self.encoder, antialias_extra=13 if self.cfg["codec_encoder"].get("antialias") else 0)
        chunk_frames = max(1, int(round(chunk_seconds * 16000)) // hop)
        plan = longform.plan_chunks(x.numel() // hop, chunk_frames, halo)
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(group) if world > 1 else 0
        c0, c1 = sharding_range(len(plan), rank, world)
        with precision_scope(self.precision):
            parts = longform.chunked_front(self.encoder.front_cl, x, hop, chunk_frames, halo, micro_batch, range(c0, c1))
            feat = longform.gather_features(parts, len(plan), lambda i: plan[i][1] - plan[i][0], self.encoder.enc_dim,
                                            x.device, group)
            if feat is None:
                return None
            return self._indices_from_features(feat)

    @torch.no_grad()
    def extract_indices(self, wave_host: torch.Tensor, micro_batch: int = 8, rnn_batch: int = 512,
                        deep_batch: int = 64) -> np.ndarray:
        """Host (ideally pinned) float32 waveforms [N,1,T] -> int16 numpy [N,T',n_q].

        Each micro-batch is copied H2D on a side stream while the previous one computes; the int16
        results of each LSTM group come back D2H asynchronously.  Equal-length clips only (ragged
        batches are not equivalent to the reference's per-utterance padding)."""
        if wave_host.is_cuda:
            raise ValueError("extract_indices takes HOST waveforms; use indices_device for device tensors")
        with precision_scope(self.precision):
            return self._extract_indices(wave_host, micro_batch, rnn_batch, deep_batch)

    def _extract_indices(self, wave_host, micro_batch, rnn_batch, deep_batch):
        dev = next(self.parameters()).device
        N, _, T = wave_host.shape
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        results = []
        # two device staging buffers, filled on the copy stream while the other one is being encoded; explicit
        # events instead of per-micro-batch allocations (no allocator traffic, no record_stream bookkeeping)
        nbuf = 2
        bufs = [torch.empty((micro_batch, T, 1), device=dev, dtype=torch.float32) for _ in range(nbuf)]
        copied = [torch.cuda.Event() for _ in range(nbuf)]
        consumed = [None] * nbuf

        def stage(i, b0, b1):
            k = i % nbuf
            with torch.cuda.stream(copy_stream):
                if consumed[k] is not None:
                    copy_stream.wait_event(consumed[k])
                bufs[k][: b1 - b0].copy_(wave_host[b0:b1].view(b1 - b0, T, 1), non_blocking=True)
                copied[k].record(copy_stream)

        spans = []
        for c0 in range(0, N, rnn_batch):
            c1 = min(N, c0 + rnn_batch)
            spans += [(b0, min(c1, b0 + micro_batch), min(c1, b0 + micro_batch) == c1) for b0 in range(c0, c1, micro_batch)]
        if spans:
            stage(0, spans[0][0], spans[0][1])
        pipe = _FrontPipeline(self.encoder, deep_batch)
        for i, (b0, b1, last_of_group) in enumerate(spans):
            k = i % nbuf
            if i + 1 < len(spans):
                stage(i + 1, spans[i + 1][0], spans[i + 1][1])
            main.wait_event(copied[k])
            pipe.push(bufs[k][: b1 - b0])
            consumed[k] = torch.cuda.Event()
            consumed[k].record(main)
            if last_of_group:
                i16 = self._indices_from_features(pipe.take())
                host = torch.empty(i16.shape, dtype=torch.int16, pin_memory=True)
                host.copy_(i16, non_blocking=True)
                results.append(host)
        torch.cuda.current_stream(dev).synchronize()
        return np.concatenate([r.numpy() for r in results], axis=0)
