"""Long-form audio (BASELINE.json configs[3]: 10-minute clips): chunked convolutional front end, one sequential back end.

The convolutional front of the encoder (stem + EncoderBlocks) has a finite receptive field, so a long recording is
cut into hop-aligned chunks that overlap by a halo; each chunk is encoded independently (as one item of a batch, or
on another GPU), the halo frames are dropped, and the frame-rate features are concatenated in time.  The LSTM is
sequential over the whole recording, so it, the final conv and the VQ then run once over the stitched features
(SURVEY.md sections 5 and 8e).  Because chunk starts are multiples of the hop (every strided conv keeps its phase)
and each output element is the same sum in the same order, the stitched features equal the unchunked ones bit for bit.

With several ranks the chunks are dealt out contiguously; the features (2 KB per frame) make ONE ordered hand-off to
the rank that owns the LSTM (``torch.distributed.gather``: NCCL between GPUs, gloo in the CPU tests) -- the only
exchange step on this path.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .sharding import shard_range


def front_context_samples(kernel_sizes_strides_dilations: Sequence[Tuple[int, int, int, int]]) -> Tuple[int, int]:
    """(left, right): output element f of a chain of (kernel, stride, dilation, left_pad) convolutions (listed input ->
    output, total stride J) reads the input samples [f*J - left, f*J + right]."""
    left = right = 0
    jump = 1
    for k, s, d, pad_left in kernel_sizes_strides_dilations:
        span = (k - 1) * d
        left += pad_left * jump
        right += (span - pad_left) * jump
        jump *= s
    return left, right


def encoder_front_chain(encoder) -> List[Tuple[int, int, int, int]]:
    """The (kernel, stride, dilation, left_pad) chain of ``BigCodecEncoder.front_cl`` (stem, ResidualUnits, strided convs)."""
    from .vq.module import CausalConv1d, EncoderBlock, ResidualUnit, _Conv1dWN
    chain = []

    def conv_of(m):
        return m.conv if isinstance(m, CausalConv1d) else m

    def add(m):
        c = conv_of(m)
        chain.append((c.kernel_size, c.stride, c.dilation, c.left_pad))

    front, _, _, _ = encoder._split()
    for m in front:
        if isinstance(m, (CausalConv1d, _Conv1dWN)):
            add(m)
        elif isinstance(m, EncoderBlock):
            for sub in m.block:
                if isinstance(sub, ResidualUnit):
                    add(sub.block[1])
                    add(sub.block[3])
                elif isinstance(sub, (CausalConv1d, _Conv1dWN)):
                    add(sub)
    return chain


def halo_frames(encoder, antialias_extra: int = 0) -> int:
    """Frames of overlap per side that cover the front end's receptive field (+1 frame of slack)."""
    hop = int(encoder.hop_length)
    left, right = front_context_samples(encoder_front_chain(encoder))
    return (max(left, right) + hop - 1) // hop + 1 + antialias_extra


def plan_chunks(total_frames: int, chunk_frames: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """[(frame0, frame1, in_frame0, in_frame1)]: output frames [frame0, frame1) are computed from the input frames
    [in_frame0, in_frame1) (clipped to the recording, so true edges keep the reference's zero padding)."""
    if total_frames <= 0 or chunk_frames <= 0 or halo < 0:
        raise ValueError("plan_chunks: total_frames and chunk_frames must be positive, halo non-negative")
    out = []
    for f0 in range(0, total_frames, chunk_frames):
        f1 = min(total_frames, f0 + chunk_frames)
        out.append((f0, f1, max(0, f0 - halo), min(total_frames, f1 + halo)))
    return out


def chunked_front(front: Callable[[torch.Tensor], torch.Tensor], x: torch.Tensor, hop: int, chunk_frames: int, halo: int,
                  micro_batch: int = 8, chunk_ids: Optional[Sequence[int]] = None) -> List[Tuple[int, torch.Tensor]]:
    """Run ``front`` ([B, T, 1] -> [B, T/hop, C]) over the chunks of one recording ``x`` [T] (T a multiple of hop).

    Chunks with the same input length (all interior ones) go through ``front`` together, ``micro_batch`` at a time.
    Returns [(chunk id, features [frames, C] with the halo removed)] for ``chunk_ids`` (default: all), in id order."""
    T = x.numel()
    if T % hop != 0:
        raise ValueError(f"long-form input must be a multiple of the hop ({hop}); pad it (extract.prepare_waveform)")
    plan = plan_chunks(T // hop, chunk_frames, halo)
    ids = list(range(len(plan))) if chunk_ids is None else list(chunk_ids)
    by_len = {}
    for i in ids:
        f0, f1, a, b = plan[i]
        by_len.setdefault(b - a, []).append(i)
    feats = {}
    for n_in, members in by_len.items():
        for m0 in range(0, len(members), micro_batch):
            group = members[m0:m0 + micro_batch]
            xb = torch.stack([x[plan[i][2] * hop: plan[i][3] * hop] for i in group]).unsqueeze(-1)
            h = front(xb)                                                     # [len(group), n_in, C]
            for j, i in enumerate(group):
                f0, f1, a, _ = plan[i]
                feats[i] = h[j, f0 - a: f1 - a]
    return [(i, feats[i]) for i in sorted(feats)]


def stitch(parts: Sequence[Tuple[int, torch.Tensor]]) -> torch.Tensor:
    """[(chunk id, [frames, C])] -> [1, total frames, C] in chunk order."""
    return torch.cat([p for _, p in sorted(parts, key=lambda kv: kv[0])], dim=0).unsqueeze(0)


def gather_features(local_parts: Sequence[Tuple[int, torch.Tensor]], n_chunks: int, frames_of: Callable[[int], int],
                    channels: int, device, group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """The ordered hand-off: every rank contributes the features of its (contiguous) chunks; ``dst`` receives
    [1, total frames, C], the others None.  One gather of equal-sized (padded) blocks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stitch(local_parts)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    owned = [shard_range(n_chunks, r, world) for r in range(world)]
    frames = [sum(frames_of(i) for i in range(a, b)) for a, b in owned]
    cap = max(max(frames), 1)
    block = torch.zeros((cap, channels), device=device, dtype=torch.float32)
    if local_parts:
        mine = torch.cat([p for _, p in sorted(local_parts, key=lambda kv: kv[0])], dim=0)
        block[: mine.shape[0]] = mine
    bufs = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, frames)], dim=0).unsqueeze(0)
