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

import random
from typing import Mapping

import torch
from torch import nn

from . import ops, parallel
from .registry import AnchorRegistry

__all__ = ['BaseAnchor', 'NearestAnchor', 'MultinomialAnchor', 'CachedAnchor']


class BaseAnchor(nn.Module):

    def __init__(self, *args, sync: bool = False, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._sync = sync

    needs_columns = True   # gather() consumes the per-code nearest-token keys of the column arg-min pass
    needs_distance = False  # True: gather() reads the materialised [N, K] distance matrix (compatibility mode)
    peer_exchange = False  # True: `gather_local` + the fused peer-memory exchange kernel replace gather()'s collectives

    @property
    def sync(self) -> bool:
        return self._sync

    def gather(self, x: torch.Tensor, column_keys: torch.Tensor | None, n_local: int,
               num_codes: int | None = None, distance: torch.Tensor | None = None) -> torch.Tensor:
        raise NotImplementedError


@AnchorRegistry.register_()
class NearestAnchor(BaseAnchor):

    peer_exchange = True

    @torch.no_grad()
    def gather_local(self, x, column_keys, n_local, out=None):
        """This rank's nearest-token row per code (before any exchange every key points into the local tokens)."""
        offset = parallel.rank() * n_local if self._sync else 0
        return ops.gather_rows_by_key(x, column_keys, offset, out=out)

    @torch.no_grad()
    def gather(self, x, column_keys, n_local, num_codes=None, distance=None):
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
    """anchors.py:88-104, COMPATIBILITY MODE: one token per code drawn from softmax over the tokens of that code's
    distance COLUMN.  The quantizer materialises the [N, K] matrix on demand (`vqb_distance_matrix`); the sampling
    is the reference's own call sequence (`d.T.softmax(1).multinomial(1)`), so the device RNG stream is consumed
    exactly as the reference consumes it; the rows are fetched by the row-gather kernel.  sync=True gathers tokens
    and distances of every rank first (anchors.py:50-53), sync=False averages the per-rank anchors (anchors.py:64-67)."""

    needs_columns = False
    needs_distance = True

    @torch.no_grad()
    def gather(self, x, column_keys, n_local, num_codes=None, distance=None):
        assert distance is not None, 'MultinomialAnchor.gather needs the materialised distance matrix'
        d = distance.detach()
        if self._sync and parallel.world_size() > 1:
            x = parallel.all_gather_rows(x)
            d = parallel.all_gather_rows(d)
        indices = d.t().softmax(1).multinomial(1).flatten()        # anchors.py:98-101
        anchors = ops.gather_rows_by_key(x.contiguous(), indices.to(torch.int64).contiguous(), 0)
        return anchors if self._sync else parallel.all_reduce_sum_(anchors)


def cached_rows_and_indices(x: torch.Tensor, num_codes: int, cache: torch.Tensor):
    """The sampling rule of CachedAnchor._anchors (anchors.py:140-166), which never looks at the distance VALUES:
    tokens (topped up with the previous anchors, then with uniform noise, when there are fewer tokens than codes)
    and K row indices — a random permutation when rows <= K, else a sample without replacement.  Same RNG calls
    in the same order as the reference (`torch.randperm` / `random.sample` on the host, `torch.rand` on the
    token device), so equal seeds give equal anchors."""
    K = int(num_codes)
    rows = x
    if rows.shape[0] < K and cache.numel() > 0:
        rows = torch.cat([rows.to(cache.dtype), cache])
    indices = torch.randperm(K) if rows.shape[0] <= K else torch.tensor(random.sample(range(rows.shape[0]), K))
    if rows.shape[0] < K:
        missing = torch.rand(K - rows.shape[0], rows.shape[1], device=rows.device)
        rows = torch.cat([rows.to(missing.dtype), missing])
    return rows, indices


@AnchorRegistry.register_()
class CachedAnchor(BaseAnchor):
    """anchors.py:107-166: K random token rows per step (no distances involved), remembered in the `_cache`
    buffer to top up small batches.  sync=True gathers the tokens of every rank first (anchors.py:50-51; every
    rank must then draw the same indices, as in the reference); sync=False averages the per-rank anchors
    (anchors.py:64-67: all_reduce here, the 1/world in the blend kernel)."""

    needs_columns = False

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.register_buffer('_cache', torch.empty(0))

    @property
    def cache(self) -> torch.Tensor:
        return self.get_buffer('_cache')

    def _load_from_state_dict(self, state_dict: Mapping[str, torch.Tensor], prefix: str, *args, **kwargs) -> None:
        cache = state_dict.get(f'{prefix}_cache')
        if cache is not None:
            self.cache.resize_(cache.shape)      # anchors.py:131-133
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    @torch.no_grad()
    def gather(self, x, column_keys, n_local, num_codes=None, distance=None):
        assert num_codes is not None, 'CachedAnchor.gather needs the codebook size'
        if self._sync and parallel.world_size() > 1:
            x = parallel.all_gather_rows(x)
        rows, indices = cached_rows_and_indices(x, num_codes, self.cache.to(x.device))
        keys = indices.to(device=rows.device, dtype=torch.int64).contiguous()   # row index in the low 32 key bits
        anchors = ops.gather_rows_by_key(rows.contiguous(), keys, 0)
        if not self._sync:
            parallel.all_reduce_sum_(anchors)
        # the reference caches what BaseAnchor.forward returns, i.e. the rank-averaged anchors (anchors.py:161-164)
        world = parallel.world_size()
        self.register_buffer('_cache', anchors.detach().clone() if self._sync or world == 1 else anchors / world)
        return anchors
