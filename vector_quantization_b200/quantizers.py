"""Quantizer modules under the reference's registry names, config keys and forward contract
(SURVEY.md §8b):  forward(x [N, D], memo) -> (z [N, D], loss [], memo)  with
memo['x'], memo['quant'] (int64 [N]; int32 for FSQ), memo['encode'], memo['decode'], memo['loss'][name].

Reference:
    BaseQuantizer            vq/tasks/image_tokenization/models/quantizers/base.py:26-182
    VectorQuantizer          vq/algorithms/vq/quantizers.py:19-117
    VQGANQuantizer           vq/algorithms/vqgan/quantizer.py:11-21
    VQKDQuantizer            vq/algorithms/vqkd/quantizers/base.py:11-15
    ScalarQuantizer          vq/algorithms/sq/quantizers.py:9-11
    FiniteScalarQuantizer    vq/algorithms/fsq/quantizers.py:74-150
The arithmetic runs in the sm_100a kernels of libvqb200.so: tensors must live on a CUDA device and the
library must be present — there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import weakref
from typing import Mapping, Sequence

import torch
from torch import nn

from . import functional as Fq
from . import ops, parallel
from ._lib import VQBError
from .callbacks import ComposedCallback
from .distances import BaseDistance
from .registry import (Config, InitRegistry, ModelRegistry, ModuleDict, VQITQuantizerCallbackRegistry,
                       VQITQuantizerDistanceRegistry, VQITQuantizerLossRegistry, VQITQuantizerRegistry,
                       build_module_dict, get_config)

__all__ = ['BaseQuantizer', 'VectorQuantizer', 'VQGANQuantizer', 'VQKDQuantizer', 'ScalarQuantizer',
           'FiniteScalarQuantizer']


def get_memo(memo: dict, key: str) -> dict:
    """vq/utils/misc.py:30-38."""
    if key not in memo:
        memo[key] = dict()
    assert isinstance(memo[key], dict)
    return memo[key]


def _check_tokens(x: torch.Tensor, dim: int) -> torch.Tensor:
    if not x.is_cuda:
        raise VQBError('vector_quantization_b200 quantizers run on CUDA (sm_100a) tensors only; no CPU fallback')
    if x.dim() != 2 or x.shape[1] != dim:
        raise ValueError(f'expected tokens of shape [N, {dim}] ("b c h w -> (b h w) c"), got {tuple(x.shape)}')
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f'tokens must be float32 or bfloat16, got {x.dtype}')
    return x.contiguous()


class BaseQuantizer(nn.Module):

    def __init__(self, *args, callbacks: ComposedCallback, losses: ModuleDict, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._callbacks = callbacks
        self._losses = losses
        self._callbacks.bind(self)

    @classmethod
    def build_pre_hook(cls, config: Mapping, registry, item) -> Mapping:
        config['callbacks'] = VQITQuantizerCallbackRegistry.build(
            Config(type='ComposedCallback', callbacks=config.get('callbacks', [])))
        config['losses'] = build_module_dict(VQITQuantizerLossRegistry, get_config(config, 'losses'))
        return config

    @property
    def embedding_dim(self) -> int:
        raise NotImplementedError

    @property
    def codebook_size(self) -> int:
        raise NotImplementedError

    @property
    def embeddings(self) -> torch.Tensor:
        raise NotImplementedError

    def _init_weights(self, config: Mapping) -> bool:
        return True

    def init_weights(self, config: Mapping) -> bool:
        config = Config(config)
        before = config.pop('before_init_weights', Config())
        after = config.pop('after_init_weights', Config())
        self._callbacks.before_init_weights(before)
        recursive = self._init_weights(config)
        return self._callbacks.after_init_weights(after, recursive)

    # -- template (reference base.py:123-182) --------------------------------------------------------
    def _encode(self, x: torch.Tensor, memo: dict):
        raise NotImplementedError

    def encode(self, x: torch.Tensor, memo: dict):
        x = self._callbacks.before_encode(x, memo)
        enc = get_memo(memo, 'encode')
        for flag in ('_normalize_codebook', '_normalize_x', '_lazy_unpack', '_zero_fill'):  # requests of the callbacks / forward
            if flag in memo:
                enc[flag] = memo[flag]
        memo.pop('_normalize_codebook', None)
        memo.pop('_lazy_unpack', None)
        memo.pop('_zero_fill', None)
        quant, memo['encode'] = self._encode(x, enc)
        quant = self._callbacks.after_encode(x, quant, memo)
        memo['encode'].pop('_column_ctx', None)      # operands kept for the callbacks' column arg-min
        return x, quant, memo

    def _decode(self, quant: torch.Tensor, memo: dict):
        raise NotImplementedError

    def decode(self, quant: torch.Tensor, memo: dict):
        quant = self._callbacks.before_decode(quant, memo)
        z, memo['decode'] = self._decode(quant, get_memo(memo, 'decode'))
        z = self._callbacks.after_decode(z, memo)
        return z, memo

    def _loss(self, z: torch.Tensor, x: torch.Tensor, memo: dict):
        """reference base.py:151-160: evaluate the configured losses, record them in the (loss) memo, sum."""
        losses = self._losses(z, x, memo)
        memo.update(losses)
        loss = x.new_zeros([], dtype=torch.float32)
        for v in losses.values():
            loss = loss + v
        return loss, memo

    def loss(self, z: torch.Tensor, x: torch.Tensor, memo: dict):
        """reference base.py:162-171."""
        z, x = self._callbacks.before_loss(z, x, memo)
        loss_memo = get_memo(memo, 'loss')
        distance = memo.get('encode', {}).get('distance') if isinstance(memo.get('encode'), dict) else None
        if distance is not None:
            loss_memo.setdefault('distance', distance)     # where EntropyLoss looks it up (losses.py:142)
        loss, memo['loss'] = self._loss(z, x, loss_memo)
        memo['loss'].pop('distance', None)
        loss = self._callbacks.after_loss(loss, memo)
        return loss, memo

    def forward(self, x: torch.Tensor, memo: dict):
        """reference base.py:173-182 (the template; VectorQuantizer overrides it with the fused kernels)."""
        x, quant, memo = self.encode(x, memo)
        memo.update(x=x, quant=quant)
        z, memo = self.decode(quant, memo)
        loss, memo = self.loss(z, x, memo)
        return z, loss, memo


@VQITQuantizerRegistry.register_()
class VectorQuantizer(BaseQuantizer):
    """distance -> arg-min -> (codebook update callbacks) -> gather -> losses -> straight-through, as three
    kernel launches: pack+assign (tcgen05, no N x K matrix), [stats/update], fused gather+STE+loss."""

    def __init__(self, *args, embedding: nn.Embedding, distance: BaseDistance, precision: str = Fq.DEFAULT_PRECISION,
                 materialize_distance: bool = False, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if precision not in Fq.PRECISION_PLANES:
            raise ValueError(f'precision must be one of {sorted(Fq.PRECISION_PLANES)}')
        self._embedding = embedding
        self._distance = distance
        self.precision = precision
        # COMPATIBILITY / DEBUG switch: also build the reference's memo['encode']['distance'] ([N, K], quantizers.py:98)
        # for user callbacks that read it.  Implied by an EntropyLoss or a MultinomialAnchor in the config.
        self.materialize_distance = bool(materialize_distance)
        self._pending = weakref.WeakSet()   # CodebookRef of every forward whose backward has not run yet

    def protect_saved_codebook(self) -> None:
        """Call before ANY in-place write to the codebook: pending backwards that saved the live tensor get a private
        copy first (see functional.CodebookRef)."""
        live = self._embedding.weight.data_ptr()
        shared = [r for r in self._pending if r.tensor.data_ptr() == live]
        if shared:
            snapshot = shared[0].tensor.clone()
            for r in shared:
                r.tensor = snapshot
        self._pending.clear()

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        config['embedding'] = ModelRegistry.build_or_return(config['embedding'])
        config['distance'] = VQITQuantizerDistanceRegistry.build_or_return(config['distance'])
        return config

    @property
    def embedding(self) -> nn.Embedding:
        return self._embedding

    @property
    def distance(self) -> BaseDistance:
        return self._distance

    @property
    def embedding_dim(self) -> int:
        return self._embedding.embedding_dim

    @property
    def codebook_size(self) -> int:
        return self._embedding.num_embeddings

    @property
    def embeddings(self) -> torch.Tensor:
        return self._embedding.weight.clone()

    def _init_weights(self, config) -> bool:
        if config:                      # an empty node (no `init_weights` in the config) keeps nn.Embedding's own init
            InitRegistry.build(config)(self._embedding.weight)
        return False

    def _weight(self) -> torch.Tensor:
        W = self._embedding.weight
        if W.dtype != torch.float32:
            raise TypeError('the codebook (nn.Embedding weight) must be float32, as in the reference')
        if not W.is_cuda:
            raise VQBError('the codebook must live on a CUDA device; there is no CPU fallback')
        return W

    @property
    def wants_distance(self) -> bool:
        return (self.materialize_distance or self._callbacks.needs_distance
                or any(loss.needs_distance for loss in self._losses.values()))

    def _encode(self, x: torch.Tensor, memo: dict):
        if self.wants_distance:
            # compatibility mode: the reference's differentiable [N, K] matrix against the codebook as it is BEFORE
            # this step's update (`self.embeddings` clones, quantizers.py:84-85,97-98)
            W = self._weight()
            if memo.get('_normalize_codebook', False):
                self.protect_saved_codebook()
                ops.pack_rows(W.data, normalize=True, planes=1, writeback=W.data)     # NormalizeCallback, in place
            memo['distance'] = Fq.distance_matrix(_check_tokens(x, self.embedding_dim), W.clone(), self._distance.metric)
        return self._encode_nograd(x, memo)

    @torch.no_grad()
    def _encode_nograd(self, x: torch.Tensor, memo: dict):
        """Nearest code per token.  Launches: codebook pack (normalise-in-place + operand planes + key reset),
        [token pack unless zero-copy], tcgen05 assignment, [key unpack unless deferred to the gather kernel]."""
        x = _check_tokens(x.detach(), self.embedding_dim)
        W = self._weight().data
        metric = self._distance.metric
        keys = torch.empty((x.shape[0],), dtype=torch.int64, device=x.device)
        writeback = memo.pop('_normalize_codebook', False)
        if writeback:
            self.protect_saved_codebook()   # the codebook is normalised in place
        zero_fill = memo.pop('_zero_fill', None)
        book = Fq.pack_codebook(W, metric, precision=self.precision, writeback_normalized=writeback, reset_keys=keys,
                                tokens=x, zero_fill=zero_fill)
        if zero_fill is not None:
            memo['_zeroed'] = True
        normalize_tokens = memo.pop('_normalize_x', False)
        want_columns = self.training and self._callbacks.needs_column_nearest
        tokens = None
        if want_columns and book.fmt != 'bf16' and x.dtype == torch.bfloat16:
            tokens = ops.pack_rows(x, fmt='f16')      # one fp16 token plane shared by the row and the column pass
        Fq.nearest_code(x, book, metric, precision=self.precision, keys=keys, keys_are_reset=True,
                        normalize_tokens=normalize_tokens, tokens=tokens)
        memo['keys'] = keys
        if want_columns:
            # the column arg-min (per code, the nearest token) runs when the callback asks for it: CVQVAECallback first
            # works out WHICH codes need it this step (`column_keys`)
            memo['_column_ctx'] = dict(x=x, book=book, tokens=tokens, metric=metric)
        if memo.pop('_lazy_unpack', False):
            return keys, memo            # forward(): the fused gather kernel unpacks the indices
        return ops.unpack_keys(keys), memo

    @torch.no_grad()
    def column_keys(self, enc_memo: dict, rows=None) -> torch.Tensor:
        """Packed keys [K] of the nearest token per code (NearestAnchor's `d.argmin(0)`, cvqvae/anchors.py:83), computed
        on the operands `_encode` left in its memo; `rows` restricts the pass to a device-side list of codes.  The result
        is also stored as memo['encode']['column_keys']."""
        ctx = enc_memo.pop('_column_ctx')
        x = ctx['x']
        offset = parallel.rank() * x.shape[0] if self._callbacks.column_nearest_global else 0
        keys = Fq.column_nearest(x, ctx['book'], ctx['metric'], precision=self.precision, index_offset=offset,
                                 tokens=ctx['tokens'], rows=rows)
        enc_memo['column_keys'] = keys
        return keys

    def _decode(self, quant: torch.Tensor, memo: dict):
        """`nn.Embedding` gather, any index shape (decode_from_quant).  Differentiable w.r.t. the codebook when it
        requires grad (the unfused template path); plain gather kernel otherwise."""
        W = self._weight()
        if torch.is_grad_enabled() and W.requires_grad:
            return Fq.embedding_lookup(W, quant), memo
        return ops.embedding_gather(W.data, quant.contiguous()), memo

    def _loss(self, z: torch.Tensor, x: torch.Tensor, memo: dict):
        """Standalone `loss(z, x)` of the reference template (base.py:151-160) on ARBITRARY (z, x): the fused kernel
        is reused with z as the "codebook" and the identity assignment, which yields exactly mse(z, x) with the
        codebook-role gradient routed to z and the commitment-role gradient to x."""
        x2, z2 = x.reshape(-1, x.shape[-1]), z.reshape(-1, z.shape[-1]).float()
        index = torch.arange(x2.shape[0], dtype=torch.int64, device=x2.device)
        _, mse4, _, _ = Fq.quantize_ste_loss(_check_tokens(x2, self.embedding_dim), z2.contiguous(), index,
                                             self._loss_terms())
        losses = {name: (m.from_mse4(mse4) if m.uses_mse4 else m(z, x, memo)) for name, m in self._losses.items()}
        memo.update(losses)
        loss = mse4[0].new_zeros([])
        for v in losses.values():
            loss = loss + v
        return loss, memo

    def can_fuse_nchw(self) -> bool:
        """May the caller's NCHW <-> token-major rearranges (models/base.py:124,126-127) be folded into the kernels?
        Yes when forward() takes the fused path and no callback changes the tokens it is handed in before_encode
        (NormalizeCallback defers its normalisation into the fused kernels when `lazy_normalize_ok`)."""
        from .callbacks import BaseCallback, NormalizeCallback
        hooked = any(self._callbacks.overrides(h) for h in ('before_decode', 'after_decode', 'before_loss', 'after_loss'))
        known = all(type(c).before_encode is BaseCallback.before_encode or isinstance(c, NormalizeCallback)
                    for c in self._callbacks)
        normalizes = any(isinstance(c, NormalizeCallback) for c in self._callbacks)
        lazy = self._callbacks.lazy_normalize_ok() and not self.wants_distance
        return not hooked and known and (lazy or not normalizes) and not self.wants_distance

    def _forward_template(self, x: torch.Tensor, memo: dict):
        """The reference's unfused template (base.py:173-182 + quantizers.py:110-117) for configurations whose
        callbacks hook decode / loss: every hook point is honoured; each stage is still a kernel of the C-ABI."""
        x = _check_tokens(x, self.embedding_dim)
        x, quant, memo = self.encode(x, memo)
        memo.update(x=x, quant=quant)
        z, memo = self.decode(quant, memo)
        loss, memo = self.loss(z, x, memo)
        z = x + (z - x).detach()                       # ste, utils/ste.py:9-10
        return z, loss, memo

    def _loss_terms(self):
        want_norm = False
        for loss in self._losses.values():
            want_norm = want_norm or loss.needs_norm
        return want_norm

    def forward(self, x: torch.Tensor, memo: dict):
        if any(self._callbacks.overrides(h) for h in ('before_decode', 'after_decode', 'before_loss', 'after_loss')):
            return self._forward_template(x, memo)
        x = _check_tokens(x, self.embedding_dim)
        nchw = memo.pop('_nchw', None)     # tokenizer.quantize: x is the token-major copy of these [b, c, h, w] latents
        # The packed keys go straight into the fused gather kernel, which also emits memo['quant'] — unless a callback
        # that overrides after_encode needs int64 indices first (VQKDCallback reads the keys itself).
        lazy_unpack = self._callbacks.packed_keys_ok()
        # NormalizeCallback may defer F.normalize(x) into the fused kernels only when nobody else reads x
        memo['_lazy_normalize'] = self._callbacks.lazy_normalize_ok() and not self.wants_distance
        memo['_lazy_unpack'] = lazy_unpack
        x, index, memo = self.encode(x, memo)
        memo.pop('_lazy_normalize', None)
        normalize_x = memo.pop('_normalize_x', False)
        W = self._weight()
        ref = None
        if torch.is_grad_enabled() and (x.requires_grad or W.requires_grad):
            ref = Fq.CodebookRef(W.detach())
            self._pending.add(ref)
        if nchw is not None:
            ref = ref or (Fq.CodebookRef(W.detach()) if torch.is_grad_enabled() and nchw.requires_grad else None)
            if ref is not None:
                self._pending.add(ref)
        z, mse4, quant, xn = Fq.quantize_ste_loss(x, W, index, self._loss_terms(), index_is_keys=lazy_unpack,
                                                  normalize_x=normalize_x, codebook_ref=ref, nchw=nchw)
        memo.update(x=xn if normalize_x else x, quant=quant)
        memo['decode'] = get_memo(memo, 'decode')
        loss_memo = get_memo(memo, 'loss')
        loss = None
        distance = memo['encode'].get('distance')
        if distance is not None:
            loss_memo['distance'] = distance              # where EntropyLoss looks it up (losses.py:142)
        for name, module in self._losses.items():
            value = module.from_mse4(mse4) if module.uses_mse4 else module(z, memo['x'], loss_memo)
            loss_memo[name] = value
            loss = value if loss is None else loss + value
        loss_memo.pop('distance', None)
        if loss is None:
            loss = mse4[0].new_zeros([])
        return z, loss, memo


@VQITQuantizerRegistry.register_()
class VQGANQuantizer(VectorQuantizer):

    def _init_weights(self, config) -> bool:
        if dict(config) == dict(type='vqgan'):
            config = Config(type='uniform_', a=-1.0 / self.codebook_size, b=1.0 / self.codebook_size)
        return super()._init_weights(config)


@VQITQuantizerRegistry.register_()
class VQKDQuantizer(VectorQuantizer):

    def _init_weights(self, config) -> bool:
        return False


@VQITQuantizerRegistry.register_()
class ScalarQuantizer(BaseQuantizer):
    pass


@VQITQuantizerRegistry.register_()
class FiniteScalarQuantizer(ScalarQuantizer):
    """tanh bound -> round (STE) -> mixed-radix index, one kernel (fsq/quantizers.py:108-126)."""

    def __init__(self, *args, eps: float = 1e-3, num_scalars_per_channel: Sequence[int], **kwargs) -> None:
        super().__init__(*args, **kwargs)
        levels = tuple(int(v) for v in num_scalars_per_channel)
        self._eps = eps
        self._levels = levels
        self._params = ops.fsq_params(levels, eps)
        max_per_digit = torch.tensor(levels, dtype=torch.int)
        cumprod = torch.tensor((1,) + levels[:-1]).cumprod(0)
        self._codebook_size = int(max_per_digit.prod().item())
        quant = torch.arange(self._codebook_size).unsqueeze(-1)
        digits = (quant // cumprod) % max_per_digit
        self.register_buffer('_embeddings', digits / (max_per_digit // 2) - 1)  # implicit codebook table, :90-93

    @property
    def embedding_dim(self) -> int:
        return len(self._levels)

    @property
    def codebook_size(self) -> int:
        return self._codebook_size

    @property
    def embeddings(self) -> torch.Tensor:
        return self.get_buffer('_embeddings')

    def _encode(self, x: torch.Tensor, memo: dict):
        x = _check_tokens(x, self.embedding_dim)
        zq, quant = Fq.fsq_quantize(x, self._params)
        memo['z'] = zq
        return quant, memo

    def _decode(self, quant: torch.Tensor, memo: dict):
        if 'z' in memo:
            return memo['z'], memo
        if not quant.is_cuda:
            raise VQBError('FSQ decode runs on CUDA tensors only; there is no CPU fallback')
        return ops.fsq_decode(quant.contiguous(), self._params), memo

    def decode(self, quant: torch.Tensor, memo: dict):
        enc, dec = get_memo(memo, 'encode'), get_memo(memo, 'decode')
        if 'z' in enc:
            dec['z'] = enc['z']
        return super().decode(quant, memo)

    def forward(self, x: torch.Tensor, memo: dict):
        x, quant, memo = self.encode(x, memo)
        memo.update(x=x, quant=quant)
        z, memo = self.decode(quant, memo)
        memo['loss'] = get_memo(memo, 'loss')
        return z, x.new_zeros([]), memo
