"""Distance tags under the reference's registry names (config builds `f'{strategy}Distance'`,
configs/vq/distance.py:7).

Reference: vq/algorithms/vq/distances.py:28-46 — L2Distance = torch.cdist (true distance),
CosineDistance = 1 - normalize(x) normalize(e)^T, both materialising [N, K].  On the B200 path the distance
is never materialised: the modules only select the metric of the fused tcgen05 assignment kernel
(argmin ||x-e|| == argmax <x,e> - 0.5|e|^2 ; argmin 1-cos == argmax <x, e/|e|>).
"""
from __future__ import annotations

import torch
from torch import nn

from .registry import VQITQuantizerDistanceRegistry

__all__ = ['BaseDistance', 'L2Distance', 'CosineDistance']


class BaseDistance(nn.Module):
    metric: str = ''

    def forward(self, x: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
        """COMPATIBILITY MODE: the materialised [N, K] matrix, as the reference's modules return it.  The quantizer
        itself never calls this (it hands `metric` to the fused assignment kernel); it serves user code that calls
        `quantizer.distance(x, e)` and the on-demand `memo['encode']['distance']`."""
        from . import functional as Fq
        return Fq.distance_matrix(x, e, self.metric)


@VQITQuantizerDistanceRegistry.register_()
class L2Distance(BaseDistance):
    metric = 'L2'


@VQITQuantizerDistanceRegistry.register_()
class CosineDistance(BaseDistance):
    metric = 'Cosine'
