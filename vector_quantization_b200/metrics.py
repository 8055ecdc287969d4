"""Codebook usage statistics at validation (vq/tasks/image_tokenization/runners/metrics.py:25-73):
accumulate bincount(quant) over the validation set, all-reduce, then
    codebook_usage = #nonzero / K                      (:63)
    codebook_ppl   = categorical ENTROPY in nats       (:70-73; not its exponential)
The accumulation runs in the warp-aggregated histogram kernel (vqb_bincount_accumulate); the final K-sized
summary is a handful of scalar ops evaluated once per validation run.
"""
from __future__ import annotations

import re

import torch

from . import ops, parallel

__all__ = ['CodebookMixin', 'CodebookUsageMetric', 'CodebookPPLMetric']


def _get(memo, path: str):
    """todd.patches.py_.get_ for paths like '["quantizer"]["quant"]' or '.a.b'."""
    obj = memo
    for key in re.findall(r'\["([^"]+)"\]|\[\'([^\']+)\'\]|\.([A-Za-z_]\w*)', path):
        k = next(p for p in key if p)
        obj = obj[k] if isinstance(obj, dict) else getattr(obj, k)
    return obj


class CodebookMixin:

    def __init__(self, *args, quant: str, codebook_size: int, **kwargs) -> None:
        super().__init__()
        self._quant = quant
        self._codebook_size = codebook_size
        self._counts: torch.Tensor | None = None

    def forward(self, batch, memo: dict) -> dict:
        quant = _get(memo, self._quant).reshape(-1).contiguous().to(torch.int64)
        if self._counts is None:
            self._counts = torch.zeros(self._codebook_size, dtype=torch.int64, device=quant.device)
        ops.bincount_accumulate(quant, self._counts)
        return memo

    __call__ = forward

    def _summary(self, memo: dict, counts: torch.Tensor) -> float:
        raise NotImplementedError

    def summary(self, memo: dict) -> float:
        if self._counts is None:
            return 0.
        counts = parallel.all_reduce_sum_(self._counts.clone())
        return self._summary(memo, counts)


class CodebookUsageMetric(CodebookMixin):

    def _summary(self, memo, counts):
        return counts.bool().sum().item() / self._codebook_size


class CodebookPPLMetric(CodebookMixin):

    def _summary(self, memo, counts):
        p = counts / counts.sum()
        return torch.distributions.Categorical(p).entropy().item()
