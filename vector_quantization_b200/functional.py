"""Autograd-level operators of the quantizer hot path.  Each forward/backward is one launch of a
hand-written kernel behind the C-ABI (`ops.py`); nothing here computes on the host or in eager PyTorch.

    l2_normalize         F.normalize(x)                         normalize.py:24
    nearest_code         distance + argmin (no N x K matrix)    vq/algorithms/vq/quantizers.py:92-100
    column_nearest       d.argmin(0)                            vq/algorithms/cvqvae/anchors.py:83
    quantize_ste_loss    embedding gather + STE + MSE losses    quantizers.py:102-117, losses.py:41-62
    fsq_quantize         FSQ bound/round/pack                   vq/algorithms/fsq/quantizers.py:108-126
(paths relative to the reference root)
"""
from __future__ import annotations

import torch

from . import ops

PRECISION_PLANES = {'exact': 3, 'high': 2, 'fast': 1}


class _L2Normalize(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.l2norm_forward(x, torch.float32)

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        return ops.l2norm_backward(gy.contiguous().float(), x)


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """Row-wise F.normalize (fp32 result, like `norm`/`div` under the reference's autocast policy)."""
    return _L2Normalize.apply(x.contiguous())


def _planes_for(t: torch.Tensor, normalized: bool, precision: str) -> int:
    if t.dtype == torch.bfloat16 and not normalized:
        return 1  # bf16 values are exact in one plane
    return PRECISION_PLANES[precision]


@torch.no_grad()
def pack_codebook(W: torch.Tensor, metric: str, *, precision: str = 'exact', writeback_normalized: bool = False,
                  reset_keys: torch.Tensor | None = None) -> ops.Operand:
    """Codebook operand: cosine -> planes of F.normalize(W) (optionally written back to W in place, which is
    NormalizeCallback's `weight.data = normalize(weight)`, normalize.py:26-28); L2 -> planes of W and 0.5|e|^2."""
    cos = metric == 'Cosine'
    normalize = cos or writeback_normalized
    return ops.pack_rows(W, normalize=normalize, planes=_planes_for(W, normalize, precision),
                         want_half_sqnorm=not cos, writeback=W if writeback_normalized else None,
                         reset_keys=reset_keys)


@torch.no_grad()
def nearest_code(x: torch.Tensor, codebook: ops.Operand, metric: str, *, precision: str = 'exact',
                 keys: torch.Tensor | None = None, index_offset: int = 0,
                 tokens: ops.Operand | None = None) -> torch.Tensor:
    """Packed (score,index) keys [N] of the nearest code of every token (row arg-min of the distance).
    Cosine arg-min is invariant to the token norm, so raw tokens are packed (one exact plane for bf16)."""
    if tokens is None:
        tokens = ops.pack_rows(x, planes=_planes_for(x, False, precision))
    if keys is None:
        keys = ops.new_keys(x.shape[0], x.device)
    return ops.assign(tokens, codebook, keys, l2=metric == 'L2', index_offset=index_offset)


@torch.no_grad()
def column_nearest(x: torch.Tensor, codebook: ops.Operand, metric: str, *, precision: str = 'exact',
                   index_offset: int = 0) -> torch.Tensor:
    """Packed keys [K]: for every code, the nearest token (column arg-min) — the same kernel with the
    operands swapped.  Cosine needs normalised token planes here (the token norm now varies along the
    reduced axis); L2 needs the tokens' 0.5|x|^2."""
    cos = metric == 'Cosine'
    toks = ops.pack_rows(x, normalize=cos, planes=_planes_for(x, cos, precision), want_half_sqnorm=not cos)
    keys = ops.new_keys(codebook.rows, x.device)
    return ops.assign(codebook, toks, keys, l2=not cos, index_offset=index_offset)


class _QuantizeSTELoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, W, quant, want_norm):
        z, mse4 = ops.gather_ste_loss(x, W, quant, want_norm=want_norm, out_dtype=torch.float32)
        ctx.save_for_backward(x, W, quant)
        ctx.want_norm = want_norm
        ctx.mark_non_differentiable(quant)
        return z, mse4

    @staticmethod
    def backward(ctx, gz, g4):
        x, W, quant = ctx.saved_tensors
        if gz is None:
            gz = torch.zeros(x.shape, dtype=torch.float32, device=x.device)
        if g4 is None:
            g4 = torch.zeros(4, dtype=torch.float32, device=x.device)
        need_gW = ctx.needs_input_grad[1]
        gx, gW = ops.quantize_backward(gz.contiguous(), x, W, quant, g4.contiguous().float(),
                                       want_norm=ctx.want_norm, need_gW=need_gW)
        return (gx if ctx.needs_input_grad[0] else None), gW, None, None


def quantize_ste_loss(x: torch.Tensor, W: torch.Tensor, quant: torch.Tensor, want_norm: bool):
    """-> (z_ste [N,D] fp32 with value x + (W[q] - x) and gradient to x only,
           mse4 [4] = {codebook, commitment, codebook(norm), commitment(norm)} MSE terms)."""
    return _QuantizeSTELoss.apply(x.contiguous(), W, quant, bool(want_norm))


class _FSQ(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, params):
        zq, idx = ops.fsq_forward(x, params)
        ctx.save_for_backward(x)
        ctx.params = params
        ctx.mark_non_differentiable(idx)
        return zq, idx

    @staticmethod
    def backward(ctx, gz, _gidx):
        (x,) = ctx.saved_tensors
        return ops.fsq_backward(gz.contiguous(), x, ctx.params), None


def fsq_quantize(x: torch.Tensor, params):
    return _FSQ.apply(x.contiguous(), params)
