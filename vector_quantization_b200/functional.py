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
                 keys: torch.Tensor | None = None, index_offset: int = 0, normalize_tokens: bool = False,
                 tokens: ops.Operand | None = None, keys_are_reset: bool = False) -> torch.Tensor:
    """Packed (score,index) keys [N] of the nearest code of every token (row arg-min of the distance).
    Cosine arg-min is invariant to the token norm, so raw tokens are used: zero-copy for bf16 tokens
    whose D needs no padding, one exact plane otherwise.  L2 on normalised tokens (LlamaGen) packs the
    normalised planes."""
    cos = metric == 'Cosine'
    if keys is None:
        keys = torch.empty((x.shape[0],), dtype=torch.int64, device=x.device)
        keys_are_reset = False
    if tokens is None:
        norm = normalize_tokens and not cos
        tokens = None if norm else ops.as_operand(x)
        if tokens is None:
            tokens = ops.pack_rows(x, normalize=norm, planes=_planes_for(x, norm, precision),
                                   reset_keys=None if keys_are_reset else keys)
            keys_are_reset = True
    if not keys_are_reset:
        keys.fill_(-1)
    return ops.assign(tokens, codebook, keys, l2=not cos, index_offset=index_offset)


@torch.no_grad()
def column_nearest(x: torch.Tensor, codebook: ops.Operand, metric: str, *, precision: str = 'exact',
                   index_offset: int = 0) -> torch.Tensor:
    """Packed keys [K]: for every code, the nearest token (column arg-min) — the same kernel with the
    operands swapped.  Cosine needs normalised token planes here (the token norm now varies along the
    reduced axis); L2 needs the tokens' 0.5|x|^2."""
    cos = metric == 'Cosine'
    keys = ops.new_keys(codebook.rows, x.device)
    if cos:
        raw = ops.as_operand(x)
        if raw is not None:
            # bf16 tokens: ONE exact raw plane + a per-column 1/|x_n| scale in the epilogue instead of three planes
            # of the normalised tokens (halves the MMA work of this pass)
            raw.inv_norm = ops.row_inv_norm(x)
            return ops.assign(codebook, raw, keys, l2=False, scale_columns=True, index_offset=index_offset)
    toks = ops.pack_rows(x, normalize=cos, planes=_planes_for(x, cos, precision), want_half_sqnorm=not cos)
    return ops.assign(codebook, toks, keys, l2=not cos, index_offset=index_offset)


class _QuantizeSTELoss(torch.autograd.Function):
    """Outputs the four MSE terms as SEPARATE 0-dim tensors so that autograd hands their upstream gradients
    back as four device scalars (no select_backward / stack kernels between the loss and our backward)."""

    @staticmethod
    def forward(ctx, x, W, index, index_is_keys, key_offset, normalize_x, want_norm):
        z, mse4, quant, xn = ops.gather_ste_loss(
            x, W, quant=None if index_is_keys else index, keys=index if index_is_keys else None,
            key_offset=key_offset, normalize_x=normalize_x, want_norm=want_norm, want_quant=index_is_keys,
            want_xnorm=normalize_x)
        if quant is None:
            quant = index
        ctx.save_for_backward(x, W, quant)
        ctx.cfg = (normalize_x, want_norm)
        if xn is None:
            xn = x.detach()
        ctx.mark_non_differentiable(quant, xn)
        m0, m1, m2, m3 = mse4.unbind(0)
        return z, m0, m1, m2, m3, quant, xn

    @staticmethod
    def backward(ctx, gz, g0, g1, g2, g3, _gq, _gxn):
        x, W, quant = ctx.saved_tensors
        normalize_x, want_norm = ctx.cfg
        if gz is None:
            gz = torch.zeros(x.shape, dtype=torch.float32, device=x.device)
        g4 = [None if g is None else g.contiguous().float() for g in (g0, g1, g2, g3)]
        gx, gW = ops.quantize_backward(gz.contiguous().float(), x, W, quant, g4, normalize_x=normalize_x,
                                       want_norm=want_norm, need_gW=ctx.needs_input_grad[1])
        return (gx if ctx.needs_input_grad[0] else None), gW, None, None, None, None, None


def quantize_ste_loss(x: torch.Tensor, W: torch.Tensor, index: torch.Tensor, want_norm: bool, *,
                      index_is_keys: bool = False, key_offset: int = 0, normalize_x: bool = False):
    """One kernel: [x' = F.normalize(x)] -> gather W[q] -> straight-through -> MSE terms.
    -> (z_ste [N,D] fp32: value x' + (W[q] - x'), gradient to x only;
        mse4 = (codebook, commitment, codebook(norm), commitment(norm)) as four 0-dim tensors;
        quant int64 [N] (unpacked from the keys when index_is_keys);  x' (detached; x itself if not normalised))"""
    z, m0, m1, m2, m3, quant, xn = _QuantizeSTELoss.apply(x.contiguous(), W, index, bool(index_is_keys),
                                                          int(key_offset), bool(normalize_x), bool(want_norm))
    return z, (m0, m1, m2, m3), quant, xn


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
