"""Quantizer losses under the reference's registry names.

Reference: vq/algorithms/vq/losses.py:41-153 (CodebookLoss, CommitmentLoss, VQGANLoss, EntropyLoss) over
todd's MSELoss(norm=, reduction='mean', weight=1).  Here the MSE reductions are NOT computed by these
modules: the fused gather+STE+loss kernel already produced
    mse4 = [codebook, commitment, codebook(norm=True), commitment(norm=True)]
(equal values per pair, different gradient routing: "codebook" terms send gradient to the codebook row,
"commitment" terms to the token), and each loss module only selects / combines entries of mse4.

State-dict compatibility: todd losses carry a step-counter buffer `_weight._steps` (constant weight 1 in
every shipped config); the same buffer names exist here so that reference checkpoints load strictly
(tools/convert_checkpoints.py:239-243,321-322).
"""
from __future__ import annotations

from typing import Mapping

import torch
from torch import nn

from .registry import Config, VQITQuantizerLossRegistry, get_config

__all__ = ['BaseLoss', 'CodebookLoss', 'CommitmentLoss', 'VQGANLoss', 'EntropyLoss']

CODEBOOK, COMMITMENT, CODEBOOK_NORM, COMMITMENT_NORM = range(4)


class _StepWeight(nn.Module):
    """Name/shape-compatible stand-in for todd's loss-weight scheduler (`_weight._steps`)."""

    def __init__(self) -> None:
        super().__init__()
        self.register_buffer('_steps', torch.tensor(0))


class _MSE(nn.Module):
    """Holds the todd MSELoss options (`norm`) and its `_weight._steps` buffer; no arithmetic."""

    def __init__(self, norm: bool = False, reduction: str = 'mean', weight: float = 1.0) -> None:
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("only reduction='mean' (the reference default) is implemented")
        self.norm = bool(norm)
        self.scale = float(weight)
        self._weight = _StepWeight()


class BaseLoss(nn.Module):
    """A quantizer loss = a linear combination of mse4 entries."""

    def __init__(self, *args, weight: float = 1.0, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.scale = float(weight)
        self._weight = _StepWeight()

    @classmethod
    def build_pre_hook(cls, config: Mapping, registry, item) -> Mapping:
        return config

    def terms(self) -> dict[int, float]:
        """{mse4 index: coefficient}"""
        raise NotImplementedError

    uses_mse4 = True          # False: the loss is evaluated by its own forward(z, x, memo) (EntropyLoss)
    needs_distance = False    # True: forward reads the materialised distance matrix (compatibility mode)

    @property
    def needs_norm(self) -> bool:
        return any(i >= CODEBOOK_NORM for i in self.terms())

    def from_mse4(self, mse4) -> torch.Tensor:
        out = None
        for i, c in self.terms().items():
            t = mse4[i] if c == 1.0 else mse4[i] * c
            out = t if out is None else out + t
        return out

    def forward(self, z: torch.Tensor, x: torch.Tensor, memo: dict) -> torch.Tensor:
        if 'mse4' not in memo:
            raise RuntimeError('quantizer losses are evaluated by the fused quantize kernel; call the quantizer')
        return self.from_mse4(memo['mse4'])


class _MSELossBase(BaseLoss):

    def __init__(self, *args, mse: _MSE | None = None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._mse = mse if mse is not None else _MSE()

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        config['mse'] = _MSE(**get_config(config, 'mse'))
        return config


@VQITQuantizerLossRegistry.register_()
class CodebookLoss(_MSELossBase):
    """mse(z, x.detach()) — vq/algorithms/vq/losses.py:41-50."""

    def terms(self):
        return {CODEBOOK_NORM if self._mse.norm else CODEBOOK: self.scale * self._mse.scale}


@VQITQuantizerLossRegistry.register_()
class CommitmentLoss(_MSELossBase):
    """mse(z.detach(), x) — vq/algorithms/vq/losses.py:53-62."""

    def terms(self):
        return {COMMITMENT_NORM if self._mse.norm else COMMITMENT: self.scale * self._mse.scale}


@VQITQuantizerLossRegistry.register_()
class VQGANLoss(BaseLoss):
    """codebook + beta * commitment, beta = 0.25 — vq/algorithms/vq/losses.py:65-127."""

    def __init__(self, *args, codebook: CodebookLoss, commitment: CommitmentLoss, beta: float = 0.25,
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._codebook = codebook
        self._commitment = commitment
        self._beta = beta

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        config['codebook'] = VQITQuantizerLossRegistry.build(get_config(config, 'codebook'), type='CodebookLoss')
        config['commitment'] = VQITQuantizerLossRegistry.build(get_config(config, 'commitment'),
                                                               type='CommitmentLoss')
        return config

    def terms(self):
        out: dict[int, float] = {}
        for i, c in self._codebook.terms().items():
            out[i] = out.get(i, 0.0) + self.scale * c
        for i, c in self._commitment.terms().items():
            out[i] = out.get(i, 0.0) + self.scale * self._beta * c
        return out


@VQITQuantizerLossRegistry.register_()
class EntropyLoss(BaseLoss):
    """vq/algorithms/vq/losses.py:130-153, COMPATIBILITY MODE.  It consumes the full [N, K] distance matrix, which
    the quantizer materialises on demand for it (`vqb_distance_matrix`, differentiable); the softmax / entropy
    arithmetic below is the reference's, call for call, on that matrix (no shipped config uses this loss, and the
    reference looks the matrix up in the LOSS memo, `memo['distance']`, where nothing upstream puts it — here
    `BaseQuantizer.loss` places the encode memo's matrix there)."""

    uses_mse4 = False
    needs_distance = True

    def __init__(self, *args, temperature: float, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._temperature = temperature

    def terms(self):
        return {}

    def forward(self, z: torch.Tensor, x: torch.Tensor, memo: dict) -> torch.Tensor:
        affinity = memo['distance']
        flat_affinity = affinity.reshape(-1, affinity.shape[-1])
        flat_affinity = flat_affinity / self._temperature
        probs = flat_affinity.softmax(-1)
        log_probs = torch.log_softmax(flat_affinity + 1e-5, -1)
        avg_probs = probs.mean(0)
        avg_entropy = -torch.sum(avg_probs * torch.log(avg_probs + 1e-5))
        sample_entropy = -torch.mean(torch.sum(probs * log_probs, -1))
        return (sample_entropy - avg_entropy) * self.scale


def build_losses(config: Mapping):
    from .registry import build_module_dict
    return build_module_dict(VQITQuantizerLossRegistry, Config(config))
