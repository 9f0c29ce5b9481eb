"""Long-form audio (BASELINE.json configs[3]: 10-minute clips): chunked convolutional front end, one sequential back end.

The convolutional front of the encoder (stem + EncoderBlocks) has a finite receptive field, so a long recording is
cut into hop-aligned chunks that overlap by a halo; each chunk is encoded independently (as one item of a batch, or
on another GPU), the halo frames are dropped, and the frame-rate features are concatenated in time.  The LSTM is
sequential over the whole recording, so it, the final conv and the VQ then run once over the stitched features
(SURVEY.md sections 5 and 8e).  Because chunk starts are multiples of the hop (every strided conv keeps its phase)
and each output element is the same sum in the same order, the stitched features equal the unchunked ones bit for bit.

With several ranks the chunks are dealt out contiguously; the features (2 KB per frame) make ONE ordered hand-off to
the rank that owns the LSTM -- the only exchange step on this path.  Between GPUs it is a set of PEER STORES over NVLink
(``PeerFeatureBuffer``: the owner's receive buffer is mapped into every process through a CUDA IPC handle and each rank
copies its rows into place with one ``cudaMemcpyAsync``), not a collective and not NCCL; only the
64-byte handle and the completion barrier travel over the host-side (gloo) group.  CPU tensors (the gloo tests) use
``torch.distributed.gather``.
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


class PeerFeatureBuffer:
    """[rows, channels] float32 receive buffer on rank ``dst``'s GPU that every rank of ``group`` can store into.

    The owner allocates it with plain ``cudaMalloc`` (``bc_ipc_alloc``), the 64-byte IPC handle is broadcast once over the
    host-side group, the peers map it (``bc_ipc_open``).  ``put`` = one ``cudaMemcpyAsync`` per block of rows on the current
    stream; ``wait`` = stream synchronise + host barrier, after which the owner may read ``tensor``."""

    def __init__(self, rows: int, channels: int, device, host_group, dst: int = 0):
        import ctypes
        from . import _cabi
        self.lib = _cabi.load_library()
        self.rows, self.channels, self.device, self.group, self.dst = int(rows), int(channels), torch.device(device), host_group, dst
        self.rank = dist.get_rank(host_group)
        self.owner = self.rank == dst
        nbytes = max(1, self.rows * self.channels * 4)
        handle = [None]
        self.ptr = ctypes.c_void_p()
        if self.owner:
            _cabi.check(self.lib.bc_ipc_alloc(ctypes.byref(self.ptr), nbytes), "bc_ipc_alloc")
            buf = ctypes.create_string_buffer(64)
            _cabi.check(self.lib.bc_ipc_export(self.ptr, buf), "bc_ipc_export")
            handle[0] = buf.raw
        dist.broadcast_object_list(handle, src=dist.get_global_rank(host_group, dst) if host_group is not None else dst, group=host_group)
        if not self.owner:
            _cabi.check(self.lib.bc_ipc_open(handle[0], ctypes.byref(self.ptr)), "bc_ipc_open")
        self._tensor = None

    @property
    def tensor(self) -> torch.Tensor:
        """The buffer as a torch tensor (owner only)."""
        if not self.owner:
            raise RuntimeError("PeerFeatureBuffer.tensor: only the owning rank reads the buffer")
        if self._tensor is None:
            iface = {"shape": (self.rows, self.channels), "typestr": "<f4", "data": (self.ptr.value, False), "version": 3,
                     "strides": None}
            holder = type("_CudaBuffer", (), {"__cuda_array_interface__": iface})()
            self._holder = holder
            self._tensor = torch.as_tensor(holder, device=self.device)
        return self._tensor

    def put(self, row0: int, block: torch.Tensor) -> None:
        """Store ``block`` [n, channels] (this rank's GPU) at rows [row0, row0 + n) of the owner's buffer."""
        from . import _cabi
        if block.numel() == 0:
            return
        block = block.contiguous()
        if block.dtype != torch.float32 or block.shape[1] != self.channels or row0 < 0 or row0 + block.shape[0] > self.rows:
            raise ValueError("PeerFeatureBuffer.put: block does not fit the buffer")
        dst = self.ptr.value + row0 * self.channels * 4
        _cabi.check(self.lib.bc_peer_copy(dst, block.data_ptr(), block.numel() * 4, _cabi.stream_ptr(block.device)), "bc_peer_copy")

    def ready(self) -> None:
        """All ranks, before the first ``put`` of a round: the owner has finished reading the previous round's rows (its
        kernels on them were queued asynchronously), so the buffer may be overwritten."""
        if self.owner:
            torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    def wait(self) -> None:
        """All ranks: my stores are complete and, after the barrier, so are everybody else's."""
        torch.cuda.current_stream(self.device).synchronize()
        dist.barrier(group=self.group)

    def close(self) -> None:
        if self.ptr and self.ptr.value:
            if self.owner:
                self.lib.bc_ipc_free(self.ptr)
            else:
                self.lib.bc_ipc_close(self.ptr)
            self.ptr = None
            self._tensor = None


_HOST_GROUPS = {}
_PEER_BUFFERS = {}


def host_group_of(group=None):
    """A gloo group with the ranks of ``group`` (created once per process; every rank must reach this call)."""
    key = id(group) if group is not None else None
    if key not in _HOST_GROUPS:
        ranks = dist.get_process_group_ranks(group) if group is not None else None
        _HOST_GROUPS[key] = dist.new_group(ranks=ranks, backend="gloo")
    return _HOST_GROUPS[key]


def gather_features(local_parts: Sequence[Tuple[int, torch.Tensor]], n_chunks: int, frames_of: Callable[[int], int],
                    channels: int, device, group=None, dst: int = 0) -> Optional[torch.Tensor]:
    """The ordered hand-off: every rank contributes the features of its (contiguous) chunks; ``dst`` receives
    [1, total frames, C], the others None.  One gather of equal-sized (padded) blocks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stitch(local_parts)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    owned = [shard_range(n_chunks, r, world) for r in range(world)]
    frames = [sum(frames_of(i) for i in range(a, b)) for a, b in owned]
    if torch.device(device).type == "cuda":
        # peer stores over NVLink: every rank writes its rows straight into the owner's buffer (no collective)
        hg = host_group_of(group)
        total = sum(frames)
        key = (id(group), total, channels, dst)
        if key not in _PEER_BUFFERS:
            for old in [k for k in _PEER_BUFFERS if k[0] == id(group)]:
                _PEER_BUFFERS.pop(old).close()
            _PEER_BUFFERS[key] = PeerFeatureBuffer(total, channels, device, hg, dst=dst)
        buf = _PEER_BUFFERS[key]
        buf.ready()
        if local_parts:
            mine = torch.cat([p for _, p in sorted(local_parts, key=lambda kv: kv[0])], dim=0)
            buf.put(sum(frames[:rank]), mine)
        buf.wait()
        if rank != dst:
            return None
        return buf.tensor.unsqueeze(0)
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
