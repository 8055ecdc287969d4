"""CVQ-VAE anchor samplers under the reference's registry names (vq/algorithms/cvqvae/anchors.py:23-166).

NearestAnchor: anchors[k] = x[argmin_n d[n, k]] (anchors.py:71-85).  The column arg-min comes from the fused
assignment kernel run with swapped operands (packed keys [K]); the exchange steps replace the reference's
all_gather of x, d[N x K], quant and p (anchors.py:50-57) / all_reduce-mean (anchors.py:64-67):
  sync=False: local nearest token per code, rows gathered locally, all_reduce(SUM); the callback divides by world.
  sync=True : keys carry the GLOBAL token index rank*N + n (same order as torch.cat(all_gather(x)));
              all_reduce(MIN) of the packed keys picks the global nearest (lowest index on ties, like
              argmin over the concatenation); each rank contributes only the rows it owns; all_reduce(SUM).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops, parallel
from .registry import AnchorRegistry

__all__ = ['BaseAnchor', 'NearestAnchor', 'MultinomialAnchor', 'CachedAnchor']


class BaseAnchor(nn.Module):

    def __init__(self, *args, sync: bool = False, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._sync = sync

    @property
    def sync(self) -> bool:
        return self._sync

    def gather(self, x: torch.Tensor, column_keys: torch.Tensor, n_local: int) -> torch.Tensor:
        raise NotImplementedError


@AnchorRegistry.register_()
class NearestAnchor(BaseAnchor):

    @torch.no_grad()
    def gather(self, x, column_keys, n_local):
        """-> [K, D] fp32 anchors, already summed over ranks (identical on every rank)."""
        if self._sync:
            offset = parallel.rank() * n_local          # column_keys were built with this offset
            parallel.all_reduce_min_keys_(column_keys)
            anchors = ops.gather_rows_by_key(x, column_keys, offset)  # zero rows for codes won by another rank
        else:
            anchors = ops.gather_rows_by_key(x, column_keys, 0)
        return parallel.all_reduce_sum_(anchors)


@AnchorRegistry.register_()
class MultinomialAnchor(BaseAnchor):
    """anchors.py:88-104 samples from softmax over the materialised distance columns — registered for config
    compatibility, not implemented on the B200 path (no shipped config uses it)."""

    def gather(self, x, column_keys, n_local):
        raise NotImplementedError('MultinomialAnchor needs the materialised N x K distance matrix')


@AnchorRegistry.register_()
class CachedAnchor(BaseAnchor):
    """anchors.py:107-166 (random permutation + cache) — registered, not implemented yet."""

    def gather(self, x, column_keys, n_local):
        raise NotImplementedError('CachedAnchor is not implemented on the B200 path yet')
