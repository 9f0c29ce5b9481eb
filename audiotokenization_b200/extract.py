"""I/O shell around the hot path: the part of extract_indices.py that sits on either side of ``model(x)``.

What the reference's driver does per utterance (extract_indices.py:100-136, 495-561) and what is mirrored here:

  * waveform preparation: optional resample to the codec rate (``torchaudio.transforms.Resample``, :129-132),
    optional right-padding to a multiple of the hop (:134-136)                      -> ``prepare_waveform``
  * encode -> indices, squeeze(1) / permute to ``(T', n_q)``, ``int16``               -> ``BigCodecModel.extract_indices``
  * ``<out>/<subset>/<speaker>/<chapter>/<fileid>.npy`` (:534-561)                   -> ``index_file_path`` / ``save_indices``
  * per-utterance exceptions are printed, counted and skipped (:565-574)            -> ``extract_to_directory``

Instead of one utterance per model call the shell groups utterances of EQUAL length (right-padding unequal
lengths is not equivalent to the reference's per-utterance zero padding), encodes each
group through the pinned-host / double-buffered path and hands the int16 arrays to a small pool of writer
threads, so file output overlaps the GPU work.  Audio container decoding (soundfile in the reference) is not
part of this package: callers supply float32 arrays, or 16-bit PCM / float32 ``.wav`` files through
``load_wav``.
"""
from __future__ import annotations

import os
import struct
import wave
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Optional, Tuple

import numpy as np
import torch


def index_file_path(output_dir: str, subset: str, fileid: str) -> str:
    """extract_indices.py:534-556: speaker / chapter from the file id ('_' first, then '-', else 'unknown')."""
    try:
        if "_" in fileid:
            parts = fileid.split("_")
            speaker_id, chapter_id = parts[0], parts[1]
        elif "-" in fileid:
            parts = fileid.split("-")
            speaker_id, chapter_id = parts[0], parts[1]
        else:
            print(f"Warning: Could not determine speaker/chapter from fileid '{fileid}'. Using 'unknown'.")
            speaker_id, chapter_id = "unknown", "unknown"
    except IndexError:
        print(f"Warning: Could not parse speaker/chapter from fileid '{fileid}'. Using 'unknown'.")
        speaker_id, chapter_id = "unknown", "unknown"
    return os.path.join(output_dir, subset, speaker_id, chapter_id, f"{fileid}.npy")


def save_indices(output_dir: str, subset: str, fileid: str, indices: np.ndarray) -> str:
    """``np.save`` of the int16 index array ``(T', n_q)`` -- also for n_q = 1: the model returns ``[n_q, 1, T']``, the
    reference squeezes the batch axis and permutes (extract_indices.py:517-532), so one quantizer gives ``(T', 1)``."""
    arr = np.asarray(indices)
    if arr.ndim == 1:
        arr = arr[:, None]
    arr = arr.astype(np.int16, copy=False)
    path = index_file_path(output_dir, subset, fileid)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.save(path, arr)
    return path


def load_wav(path: str) -> Tuple[torch.Tensor, int]:
    """16-bit PCM or 32-bit float ``.wav`` -> (float32 [channels, T], sample rate); the ``always_2d`` ``(C, T)``
    convention of the reference's loader (extract_indices.py:100-105)."""
    with open(path, "rb") as f:
        head = f.read(44)
    fmt_tag = struct.unpack("<H", head[20:22])[0] if len(head) >= 22 else 1
    if fmt_tag == 3:   # IEEE float: the stdlib reader refuses it, parse the canonical header ourselves
        ch, sr = struct.unpack("<HI", head[22:28])
        with open(path, "rb") as f:
            raw = f.read()
        pos = raw.find(b"data")
        n = struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        x = np.frombuffer(raw, dtype="<f4", count=n // 4, offset=pos + 8).reshape(-1, ch).T
        return torch.from_numpy(np.ascontiguousarray(x)), sr
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        if width != 2:
            raise ValueError(f"{path}: only 16-bit PCM and 32-bit float wav files are supported (sample width {width})")
        x = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch).T.astype(np.float32) / 32768.0
    return torch.from_numpy(np.ascontiguousarray(x)), sr


def prepare_waveform(waveform: torch.Tensor, sample_rate: int, target_sample_rate: Optional[int] = 16000,
                     pad_to_stride: Optional[int] = None) -> torch.Tensor:
    """[C, T] or [T] float waveform -> mono-first [1, T'] at the codec rate, padded like extract_indices.py:129-136."""
    w = waveform if waveform.dim() == 2 else waveform.unsqueeze(0)
    if target_sample_rate and target_sample_rate != sample_rate:
        import torchaudio
        w = torchaudio.transforms.Resample(orig_freq=sample_rate, new_freq=target_sample_rate)(w.float())
    if pad_to_stride and w.size(1) % pad_to_stride != 0:
        w = torch.nn.functional.pad(w, (0, pad_to_stride - w.size(1) % pad_to_stride), mode="constant", value=0)
    return w[:1].float()


def extract_to_directory(model, items: Iterable[Tuple[torch.Tensor, str, str]], output_dir: str, *,
                         micro_batch: int = 8, rnn_batch: int = 512, group_size: int = 256, writers: int = 4,
                         max_buffered_bytes: int = 2 << 30, max_buffered_items: int = 4096,
                         verbose: bool = True) -> Tuple[int, int]:
    """Encode ``items`` = (waveform float32 [T] or [1, T] at the codec rate, subset, fileid) and write one ``.npy``
    per utterance in the reference's layout.  Returns (saved, errors) like the counters of extract_indices.py:492-579.

    Utterances are encoded in groups of equal length (up to ``group_size``), each group through
    ``model.extract_indices`` (pinned host buffer, H2D overlapped with compute).  The buffer of waiting
    utterances is bounded: once it holds more than ``max_buffered_bytes`` of samples or ``max_buffered_items``
    utterances, the largest waiting group is encoded -- with full-length files (``duration=None``, the reference's
    default) almost every length is unique, and an unbounded buffer would hold the whole corpus before the first
    encode.  A failing group is retried one utterance at a time, so one bad utterance costs one error like in the
    reference (extract_indices.py:565-574), never a whole group.  File writes run on ``writers`` threads and are
    reaped as they complete."""
    saved = errors = 0
    pending = []
    groups = {}
    buffered_bytes = buffered_items = 0

    def reap(block: bool) -> None:
        nonlocal saved, errors, pending
        keep = []
        for fileid, fut in pending:
            if not block and not fut.done():
                keep.append((fileid, fut))
                continue
            try:
                fut.result()
                saved += 1
            except Exception as e:
                print(f"\nError saving indices of {fileid}: {e}")
                errors += 1
        pending = keep

    def encode(batch, length):
        host = torch.empty((len(batch), 1, length), dtype=torch.float32, pin_memory=torch.cuda.is_available())
        for i, (w, _, _) in enumerate(batch):
            host[i, 0] = w
        return model.extract_indices(host, micro_batch=micro_batch, rnn_batch=rnn_batch)   # [N, T', n_q]

    def flush(length):
        nonlocal errors, buffered_bytes, buffered_items
        batch = groups.pop(length, [])
        if not batch:
            return
        buffered_bytes -= 4 * length * len(batch)
        buffered_items -= len(batch)
        try:
            results = list(encode(batch, length))
        except Exception as e:
            if len(batch) == 1:
                print(f"\nError processing {batch[0][2]}: {e}")
                errors += 1
                return
            print(f"\nError processing a group of {len(batch)} utterances of {length} samples: {e}; retrying one by one")
            results = []
            for item in batch:
                try:
                    results.append(encode([item], length)[0])
                except Exception as e1:   # the reference swallows per-utterance errors (extract_indices.py:565-574)
                    print(f"\nError processing {item[2]}: {e1}")
                    errors += 1
                    results.append(None)
        for (_, subset, fileid), arr in zip(batch, results):
            if arr is not None:
                pending.append((fileid, pool.submit(save_indices, output_dir, subset, fileid, arr)))
        reap(block=False)

    with ThreadPoolExecutor(max_workers=max(1, writers)) as pool:
        for item in items:
            try:
                w, subset, fileid = item
                w = torch.as_tensor(w, dtype=torch.float32).reshape(-1)
                if w.numel() == 0:
                    raise ValueError("empty waveform")
            except Exception as e:
                print(f"\nError processing batch item: {e}")
                errors += 1
                continue
            n = w.numel()
            groups.setdefault(n, []).append((w, subset, fileid))
            buffered_bytes += 4 * n
            buffered_items += 1
            if len(groups[n]) >= group_size:
                flush(n)
            while buffered_bytes > max_buffered_bytes or buffered_items > max_buffered_items:
                flush(max(groups, key=lambda k: (k * len(groups[k]), -k)))     # the group holding the most samples
        for length in sorted(groups):
            flush(length)
        reap(block=True)
    if verbose:
        print("\nExtraction complete.")
        print(f"Successfully saved {saved} index files.")
        if errors > 0:
            print(f"Encountered {errors} errors.")
    return saved, errors
