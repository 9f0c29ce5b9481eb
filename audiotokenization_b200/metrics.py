"""Codebook usage statistics on the device (host mirror of the reference's metric classes).

``CodebookPerplexity`` / ``CodebookUtilization`` keep the reference's names, constructor argument and
``update(indices)`` / ``compute()`` / ``reset()`` protocol (lightning_module.py:26-73, torchmetrics-style),
and ``calculate_perplexity`` the return pair of inference_full.py:570-604 -- but the bookkeeping is one
histogram kernel over int32 indices (``bc_code_histogram``) instead of a ``[N, K]`` one-hot or a host
``Counter``, and the entropy is a K-element device reduction (``bc_code_entropy``).
"""
from __future__ import annotations

import math

import torch

from ._cabi import check, load_library, ptr, require_cuda, stream_ptr


def code_histogram(indices: torch.Tensor, codebook_size: int, counts: torch.Tensor | None = None) -> torch.Tensor:
    """Accumulate the usage histogram of ``indices`` (any integer shape, CUDA) into ``counts`` (int64 [K])."""
    require_cuda(indices, "indices")
    idx = indices.reshape(-1).to(torch.int32).contiguous()
    if counts is None:
        counts = torch.zeros((codebook_size,), device=idx.device, dtype=torch.int64)
    if counts.dtype != torch.int64 or counts.numel() != codebook_size or not counts.is_contiguous():
        raise ValueError("counts must be a contiguous int64 tensor of codebook_size elements")
    bad = torch.zeros((1,), device=idx.device, dtype=torch.int32)
    check(load_library().bc_code_histogram(ptr(idx), idx.numel(), codebook_size, ptr(counts), ptr(bad),
                                           stream_ptr(idx.device)), "bc_code_histogram")
    nbad = int(bad.item())
    if nbad:
        raise IndexError(f"code_histogram: {nbad} indices outside [0, {codebook_size})")
    return counts


def _entropy_used_total(counts: torch.Tensor):
    out = torch.empty((3,), device=counts.device, dtype=torch.float64)
    check(load_library().bc_code_entropy(ptr(counts), counts.numel(), ptr(out), stream_ptr(counts.device)),
          "bc_code_entropy")
    e, used, total = out.tolist()
    return e, used, total


class _CountMetric:
    def __init__(self, codebook_size: int, device="cuda"):
        self.codebook_size = int(codebook_size)
        self.codebook_counts = torch.zeros((self.codebook_size,), device=device, dtype=torch.int64)

    def update(self, indices: torch.Tensor) -> None:
        code_histogram(indices.to(self.codebook_counts.device), self.codebook_size, self.codebook_counts)

    def reset(self) -> None:
        self.codebook_counts.zero_()

    def __call__(self, indices):
        self.update(indices)
        return self.compute()


class CodebookPerplexity(_CountMetric):
    """exp(entropy) of the empirical code distribution (lightning_module.py:26-51)."""

    @property
    def total_counts(self) -> torch.Tensor:
        return self.codebook_counts.sum()

    def compute(self) -> torch.Tensor:
        e, _, total = _entropy_used_total(self.codebook_counts)
        if total == 0:
            return torch.tensor(0.0, device=self.codebook_counts.device)
        return torch.tensor(math.exp(e), device=self.codebook_counts.device, dtype=torch.float32)


class CodebookUtilization(_CountMetric):
    """Fraction of the codebook that has been used (lightning_module.py:53-69)."""

    @property
    def used_codes(self) -> torch.Tensor:
        return self.codebook_counts > 0

    def compute(self) -> torch.Tensor:
        _, used, _ = _entropy_used_total(self.codebook_counts)
        return torch.tensor(used / self.codebook_size, device=self.codebook_counts.device, dtype=torch.float32)


def calculate_perplexity(counts, codebook_size: int):
    """(normalised perplexity, perplexity) of inference_full.py:570-604.  ``counts`` is the device histogram
    (int64 [K]) or a ``collections.Counter`` / dict {index: count} like the reference's; 0.0 when empty."""
    if not isinstance(counts, torch.Tensor):
        dev = torch.device("cuda")
        c = torch.zeros((codebook_size,), dtype=torch.int64)
        for i, n in dict(counts).items():
            if 0 <= int(i) < codebook_size:   # the reference drops invalid indices here
                c[int(i)] = int(n)
        total_all = sum(dict(counts).values())
        if total_all == 0:
            return 0.0
        counts = c.to(dev)
        e, _, total = _entropy_used_total(counts)
        if total != total_all and total > 0:
            # the reference normalises by the total INCLUDING out-of-range keys: p_i = c_i / total_all
            scale = total / total_all
            e = scale * (e - math.log(scale))
    else:
        e, _, total = _entropy_used_total(counts)
        if total == 0:
            return 0.0
    return math.exp(e / math.log(codebook_size)), math.exp(e)
