"""Multi-GPU exchange steps of the hot path (one process per GPU, `torch.distributed` plumbing).

The reference is data-parallel only: tokens are sharded by the sampler, the quantizer is replicated, and
only statistics are exchanged:
  * QuantStatistics bin_count / num_elements: two all_reduce(SUM)   vq/algorithms/vq/utils.py:35
  * VQ-KD centroid sums: all_reduce(SUM) of [K, D]                  vqkd/quantizers/callbacks.py:63-64
  * CVQ-VAE anchors sync=False: all_reduce(SUM) / world             cvqvae/anchors.py:64-67
  * CVQ-VAE anchors sync=True: all_gather of x, d[N x K], quant, p  cvqvae/anchors.py:50-57
Here: ONE all_reduce(SUM) of the fused [K*D sums | K counts] buffer; ONE all_reduce(SUM) of the int64
[K counts | numel] buffer; the sync=True anchor exchange is a packed (distance, global-token-index)
min-loc all_reduce(MIN) over [K] followed by a masked [K, D] all_reduce(SUM) of the winning rows — the
N x K matrix is never gathered.  The same packed min-loc reduce over [N] combines codebook shards.

All helpers are no-ops in a single process and work on CPU tensors with the gloo backend (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ['world_size', 'rank', 'all_reduce_sum_', 'all_reduce_min_keys_', 'shard_range']

_SIGN = -(1 << 63)  # 0x8000... as int64


def _on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if _on() else 1


def rank() -> int:
    return dist.get_rank() if _on() else 0


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if _on():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_min_keys_(keys: torch.Tensor) -> torch.Tensor:
    """MIN all-reduce of packed uint64 (score, index) keys stored in an int64 tensor.
    NCCL/gloo order int64 as signed, so bit 63 is flipped before and after (on CUDA by the
    vqb_keys_flip_sign kernel)."""
    if not _on():
        return keys
    if keys.is_cuda:
        from . import ops
        ops.keys_flip_sign(keys)
        dist.all_reduce(keys, op=dist.ReduceOp.MIN)
        ops.keys_flip_sign(keys)
    else:  # host-logic tests (gloo)
        keys.bitwise_xor_(_SIGN)
        dist.all_reduce(keys, op=dist.ReduceOp.MIN)
        keys.bitwise_xor_(_SIGN)
    return keys


def shard_range(total: int, r: int | None = None, w: int | None = None) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `total` rows owned by rank r (codebook sharding)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per = (total + w - 1) // w
    lo = min(total, r * per)
    return lo, min(total, lo + per)
