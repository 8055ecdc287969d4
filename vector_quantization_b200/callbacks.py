"""Quantizer callbacks under the reference's registry names and hook protocol.

Reference protocol: 9 hook points dispatched in priority order by ComposedCallback
(vq/tasks/image_tokenization/models/quantizers/callbacks/{base,composed,lazy_init_weights}.py).
Reference callbacks re-implemented on the B200 kernels:
    NormalizeCallback   vq/algorithms/vq/callbacks/normalize.py:19-29
    UpdateMixin         vq/algorithms/vq/callbacks/update.py:15-56
    VQKDCallback        vq/algorithms/vqkd/quantizers/callbacks.py:38-129   (k-means init + per-step EMA)
    CVQVAECallback      vq/algorithms/cvqvae/quantizer_callback.py:25-105   (usage EMA + anchor re-init)
Codebook updates keep the reference's ordering: they run in `after_encode`, i.e. BEFORE the gather, so
`z` and the losses see the updated codebook.
"""
from __future__ import annotations

import random
from typing import Iterable, Mapping

import torch

from . import functional as Fq
from . import ops, parallel
from .registry import AnchorRegistry, Config, VQITQuantizerCallbackRegistry

__all__ = ['BaseCallback', 'ComposedCallback', 'UpdateMixin', 'NormalizeCallback', 'LazyInitWeightsMixin',
           'VQKDCallback', 'VQGAN_VQKDCallback', 'CVQVAECallback', 'EMA']

HOOKS = ('bind', 'before_init_weights', 'after_init_weights', 'before_encode', 'after_encode', 'before_decode',
         'after_decode', 'before_loss', 'after_loss')


class EMA:
    """todd.utils.EMA stand-in: only `decay` is consumed (the blends run inside the update kernels).
    Default decay 0.99 as in upstream CVQ-VAE / BEiT-v2 (todd default UNVERIFIED)."""

    def __init__(self, decay: float = 0.99) -> None:
        self._decay = float(decay)

    @property
    def decay(self) -> float:
        return self._decay


class BaseCallback:
    # B200 path: set by callbacks whose `after_encode` consumes the per-code nearest token
    needs_column_nearest = False
    column_nearest_global = False
    # B200 path: True if this callback's `after_encode` copes with RAW tokens when NormalizeCallback defers the
    # token normalisation into the fused kernels (it must normalise whatever it reads from x itself)
    accepts_raw_tokens = False
    # B200 path: True if `after_encode` takes the PACKED (score, index) keys of the assignment in place of int64
    # indices (`quant is memo['encode']['keys']`); the fused gather kernel then emits memo['quant'] on its way
    accepts_packed_keys = False

    def __init__(self, *args, **kwargs) -> None:
        super().__init__()
        self._instance = None

    @classmethod
    def build_pre_hook(cls, config: Mapping, registry, item) -> Mapping:
        return config

    def bind(self, instance) -> None:
        self._instance = instance

    @property
    def quantizer(self):
        return self._instance

    @property
    def vector_quantizer(self):
        from .quantizers import VectorQuantizer
        assert isinstance(self._instance, VectorQuantizer)
        return self._instance

    def before_init_weights(self, config: Mapping) -> None:
        pass

    def after_init_weights(self, config: Mapping, recursive: bool) -> bool:
        return recursive

    def before_encode(self, x: torch.Tensor, memo: dict) -> torch.Tensor:
        return x

    def after_encode(self, x: torch.Tensor, quant: torch.Tensor, memo: dict) -> torch.Tensor:
        return quant

    def before_decode(self, quant: torch.Tensor, memo: dict) -> torch.Tensor:
        return quant

    def after_decode(self, z: torch.Tensor, memo: dict) -> torch.Tensor:
        return z

    def before_loss(self, z: torch.Tensor, x: torch.Tensor, memo: dict):
        return z, x

    def after_loss(self, loss: torch.Tensor, memo: dict) -> torch.Tensor:
        return loss


@VQITQuantizerCallbackRegistry.register_()
class ComposedCallback(BaseCallback):

    def __init__(self, *args, priorities: Iterable[Mapping[str, int]], callbacks: Iterable[BaseCallback],
                 **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._priorities = [dict(p) for p in priorities]
        self._callbacks = list(callbacks)

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        callbacks = [Config(c) if isinstance(c, Mapping) else c for c in config['callbacks']]
        config['priorities'] = [c.pop('priority', dict()) if isinstance(c, Mapping) else dict() for c in callbacks]
        config['callbacks'] = [VQITQuantizerCallbackRegistry.build_or_return(c) for c in callbacks]
        return config

    def __iter__(self):
        return iter(self._callbacks)

    def _queue(self, key: str):
        order = sorted(range(len(self._callbacks)), key=lambda i: -self._priorities[i].get(key, 0))
        return [self._callbacks[i] for i in order]

    @property
    def needs_column_nearest(self) -> bool:  # type: ignore[override]
        return any(c.needs_column_nearest for c in self._callbacks)

    @property
    def column_nearest_global(self) -> bool:  # type: ignore[override]
        return any(c.column_nearest_global for c in self._callbacks)

    @property
    def needs_distance(self) -> bool:
        """Some child reads the materialised [N, K] distance matrix (compatibility mode, e.g. MultinomialAnchor)."""
        return any(getattr(c, 'needs_distance', False) for c in self._callbacks)

    def lazy_normalize_ok(self) -> bool:
        """May NormalizeCallback hand RAW tokens down the pipeline (normalisation fused into the gather / backward
        kernels)?  Only if every callback that reads x in `after_encode` declares that it copes; otherwise the
        tokens are normalised up front, exactly like the reference (normalize.py:24), so that anchors, column
        arg-mins and user callbacks see F.normalize(x)."""
        return all(c.accepts_raw_tokens or type(c).after_encode is BaseCallback.after_encode
                   for c in self._callbacks) and not self.needs_column_nearest

    def packed_keys_ok(self) -> bool:
        """May `after_encode` receive the packed keys instead of int64 indices (no separate unpack launch)?"""
        return all(c.accepts_packed_keys or type(c).after_encode is BaseCallback.after_encode for c in self._callbacks)

    def overrides(self, hook: str) -> bool:
        """True if any child customises `hook` (the fused decode/loss path checks this)."""
        return any(getattr(type(c), hook) is not getattr(BaseCallback, hook) for c in self._callbacks)

    def bind(self, instance) -> None:
        super().bind(instance)
        for c in self._queue('bind'):
            c.bind(instance)

    def before_init_weights(self, config) -> None:
        for c in self._queue('before_init_weights'):
            c.before_init_weights(config)

    def after_init_weights(self, config, recursive: bool) -> bool:
        for c in self._queue('after_init_weights'):
            recursive = c.after_init_weights(config, recursive)
        return recursive

    def before_encode(self, x, memo):
        for c in self._queue('before_encode'):
            x = c.before_encode(x, memo)
        return x

    def after_encode(self, x, quant, memo):
        for c in self._queue('after_encode'):
            quant = c.after_encode(x, quant, memo)
        return quant

    def before_decode(self, quant, memo):
        for c in self._queue('before_decode'):
            quant = c.before_decode(quant, memo)
        return quant

    def after_decode(self, z, memo):
        for c in self._queue('after_decode'):
            z = c.after_decode(z, memo)
        return z

    def before_loss(self, z, x, memo):
        for c in self._queue('before_loss'):
            z, x = c.before_loss(z, x, memo)
        return z, x

    def after_loss(self, loss, memo):
        for c in self._queue('after_loss'):
            loss = c.after_loss(loss, memo)
        return loss


class UpdateMixin(BaseCallback):

    def __init__(self, *args, ema: EMA | None = None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        if ema is not None:
            self._ema = ema
        self._region = None      # parallel.PeerRegion once the multi-GPU exchange is set up; False = unavailable

    def _peer_region(self, device: torch.device, layout) -> 'parallel.PeerRegion | None':
        """The NVLink peer region of this quantizer (multi-GPU training only), created on first use — a collective
        call: every rank reaches its first training forward.  `layout`: [(name, shape, dtype)], the codebook 'W'
        first.  The codebook (and CVQ-VAE's `_probability`) are re-homed INTO the region, because the fused exchange
        kernels publish the updated rows straight into every rank's copy (the reference rebinds `weight.data` on
        every update as well, update.py:56)."""
        if self._region is None:
            self._region = False
            if parallel.peer_comm_enabled(device):
                nbytes = sum(((torch.Size(shape).numel() * torch.empty((), dtype=dt).element_size() + 511) // 512) * 512
                             for _, shape, dt in layout)
                region = parallel.try_peer_region(nbytes, device)
                if region is not None:
                    for name, shape, dt in layout:
                        region.alloc(name, shape, dt)
                    self._region = region
        region = self._region or None
        if region is not None:
            w = self.vector_quantizer.embedding.weight
            if w.data_ptr() != region.W.data_ptr():      # first use, or the module was moved / reloaded
                region.W.copy_(w.data)
                w.data = region.W
        return region

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        if config.get('ema') is not None and not isinstance(config['ema'], EMA):
            config['ema'] = EMA(**config['ema'])
        return config

    @property
    def with_ema(self) -> bool:
        return hasattr(self, '_ema')

    def _update_embedding(self, e: torch.Tensor) -> None:
        self._check_sync(e)
        self.vector_quantizer.embedding.weight.data = e

    def _check_sync(self, t: torch.Tensor, what: str = 'codebook') -> None:
        """The reference's only runtime cross-rank check (`todd.Store.DRY_RUN` + `todd.utils.is_sync`, update.py:54-55,
        cvqvae/anchors.py:52-53,62-63): with VQB_DRY_RUN=1 every codebook update asserts that all replicas hold the
        same values."""
        if parallel.DRY_RUN:
            parallel.assert_sync(t, what)


@VQITQuantizerCallbackRegistry.register_()
class NormalizeCallback(UpdateMixin, BaseCallback):
    """x <- normalize(x);  weight.data <- normalize(weight)  every forward, train and eval.
    Neither costs a pass of its own: the codebook is normalised in place by the kernel that also emits the
    bf16 operand planes (+0.5|e|^2) for the assignment (flag `_normalize_codebook`, consumed by
    `VectorQuantizer._encode`), and inside `forward` the token normalisation is deferred into the fused
    gather/STE/loss kernel and its backward (flag `_normalize_x`); a standalone `encode()` call
    normalises eagerly with the l2norm kernel."""

    def before_encode(self, x, memo):
        x = super().before_encode(x, memo)
        memo['_normalize_codebook'] = True
        if memo.pop('_lazy_normalize', False):
            memo['_normalize_x'] = True   # x stays raw; F.normalize is applied inside the fused kernels
            return x
        return Fq.l2_normalize(x)


class LazyInitWeightsMixin(BaseCallback):
    """Runs `lazy_init_weights(config, x, memo)` once, on the first forward (forward pre-hook)."""

    def lazy_init_weights(self, config: Mapping, x: torch.Tensor, memo: dict) -> None:
        raise NotImplementedError

    def before_init_weights(self, config) -> None:
        super().before_init_weights(config)
        lazy_cfg = config.pop('lazy_init_weights', Config()) if hasattr(config, 'pop') else Config()

        def forward_pre_hook(module, args):
            x, memo = args
            self.lazy_init_weights(lazy_cfg, x, memo)
            handle.remove()

        handle = self.quantizer.register_forward_pre_hook(forward_pre_hook)


@VQITQuantizerCallbackRegistry.register_()
class VQKDCallback(LazyInitWeightsMixin, NormalizeCallback):
    """Per training step: counts + sums of normalised tokens per code (warp-aggregated scatter-add) ->
    ONE all-reduce of the fused [K*D | K] buffer -> centroid / normalise / EMA / normalise kernel."""

    accepts_raw_tokens = True    # after_encode accumulates F.normalize(x) rows itself (scatter kernel flag)
    accepts_packed_keys = True   # ... and reads the code of a token straight from the packed keys

    @torch.no_grad()
    def lazy_init_weights(self, config, x, memo) -> None:
        """k-means init (callbacks.py:77-112) on the kernels of the training step, DISTRIBUTED: the reference gathers
        every rank's tokens to rank 0 (`distributed_cat`, :26-35), runs k-means there (CPU offload above 2^30
        distance elements, :97-100) and broadcasts.  Here every rank keeps its tokens: rank 0 draws the K seed
        indices into the concatenated token order with `random.sample` (the reference's RNG call, so equal seeds
        give equal seeds) and broadcasts them; each round is one assignment + one statistics pass per rank and an
        all-reduce of the [K*D | K] sums.  No N x K matrix, no gather, no offload; the result equals the
        rank-0 k-means up to fp32 summation order."""
        vq = self.vector_quantizer
        if not vq.training:
            return
        x = x.detach().contiguous()
        world, rank = parallel.world_size(), parallel.rank()
        n_local = x.shape[0]                       # equal on every rank (the reference's gather needs that too)
        W = vq.embedding.weight.data
        K = W.shape[0]
        iters = config.get('iters', 10)
        metric = vq.distance.metric
        if n_local * world < K:
            rows = parallel.all_gather_rows(x.float())            # fewer tokens than codes: raw tokens first (:91-92)
            W[:rows.shape[0]] = rows
        else:
            xn = ops.l2norm_forward(x)
            if rank == 0:
                indices = torch.tensor(random.sample(range(n_local * world), K), dtype=torch.int64)
            else:
                indices = torch.empty(K, dtype=torch.int64)
            indices = indices.to(x.device)
            if world > 1:
                torch.distributed.broadcast(indices, 0)
            # every rank contributes the seed rows it owns (zero rows elsewhere); the sum is x_cat[indices]
            seeds = ops.gather_rows_by_key(xn, indices, rank * n_local)
            W.copy_(parallel.all_reduce_sum_(seeds))
            for _ in range(iters):
                keys = ops.new_keys(n_local, xn.device)
                book = Fq.pack_codebook(W, metric, precision=vq.precision, writeback_normalized=True)
                quant = ops.unpack_keys(Fq.nearest_code(xn, book, metric, precision=vq.precision, keys=keys,
                                                        keys_are_reset=True))
                stats = parallel.all_reduce_sum_(ops.scatter_stats(xn, quant, K))
                ops.kmeans_ema_update(stats, W, 0.0)  # decay 0: W <- normalize(centroids | old row if unused)
        Fq.pack_codebook(W, metric, precision=vq.precision, writeback_normalized=True)

    def _stats_buffer(self, device: torch.device):
        """(region | None, stats): the per-step [K*D sums | K counts] buffer — inside the NVLink peer region when the
        fused multi-GPU exchange is active, a persistent private buffer otherwise (padded to 16 bytes for the
        fused zero-fill)."""
        W = self.vector_quantizer.embedding.weight.data
        K, D = W.shape
        layout = [('W', (K, D), torch.float32), ('stats', ((K * D + K + 3) // 4 * 4,), torch.float32)]
        if (K * D + K) * 4 <= parallel.LL_MAX_BYTES:      # small payload: latency matters, not the 2x wire bytes
            layout += ops.comm_ll_layout(K, D, parallel.world_size())
        region = self._peer_region(device, layout)
        if region is not None:
            return region, region.stats
        buf = getattr(self, '_stats', None)
        if buf is None or buf.device != device or buf.numel() != (K * D + K + 3) // 4 * 4:
            buf = self._stats = torch.zeros((K * D + K + 3) // 4 * 4, dtype=torch.float32, device=device)
        return None, buf

    def before_encode(self, x, memo):
        x = super().before_encode(x, memo)
        if self.vector_quantizer.training and x.is_cuda:
            # the statistics buffer is zeroed by the codebook-pack launch of `_encode` (no memset of its own)
            memo['_zero_fill'] = self._stats_buffer(x.device)[1]
        return x

    @torch.no_grad()
    def after_encode(self, x, quant, memo):
        quant = super().after_encode(x, quant, memo)
        vq = self.vector_quantizer
        if not vq.training:
            return quant
        W = vq.embedding.weight.data
        K, D = W.shape
        vq.protect_saved_codebook()
        packed = quant is memo['encode'].get('keys')          # forward() hands the packed keys down (no unpack launch)
        index = dict(quant=None, keys=quant) if packed else dict(quant=quant)
        region, stats = self._stats_buffer(x.device)
        if not memo['encode'].pop('_zeroed', False):
            stats.zero_()                                     # `_encode` was bypassed or overridden
        ops.scatter_stats(x.detach(), **index, K=K, normalize_x=True, out=stats)
        if region is not None:
            # the per-rank partial sums sit in the peer region: ONE fused launch reduces them over NVLink in fixed rank
            # order, applies the k-means/EMA update and publishes the new rows to every replica
            ops.comm_kmeans_ema_update(region, K, D, self._ema.decay)
            self._check_sync(region.W)
            return quant
        parallel.all_reduce_sum_(stats)
        ops.kmeans_ema_update(stats, W, self._ema.decay)
        self._check_sync(W)
        return quant


@VQITQuantizerCallbackRegistry.register_()
class VQGAN_VQKDCallback(VQKDCallback):  # noqa: N801 - the reference's registry name
    """vq/algorithms/exp/vqgan_vqkd/quantizer_callback.py:38-134: the k-means lazy init of VQKDCallback (:79-114,
    the same distributed implementation here) on a codebook that is trained by GRADIENT; per training step only
    `W <- normalize(ema(W, normalize(W)))` (:124-134) — no statistics, no exchange.  That is the k-means/EMA update
    kernel with every code "unused" (a zero statistics buffer: the centroid of an unused code is its old row)."""

    def before_encode(self, x, memo):
        return NormalizeCallback.before_encode(self, x, memo)     # no statistics buffer to zero

    @torch.no_grad()
    def after_encode(self, x, quant, memo):
        vq = self.vector_quantizer
        if not vq.training:
            return quant
        W = vq.embedding.weight.data
        K, D = W.shape
        zeros = getattr(self, '_zeros', None)
        if zeros is None or zeros.device != W.device or zeros.numel() != K * D + K:
            zeros = self._zeros = torch.zeros(K * D + K, dtype=torch.float32, device=W.device)
        vq.protect_saved_codebook()
        ops.kmeans_ema_update(zeros, W, self._ema.decay)
        self._check_sync(W)
        return quant


@VQITQuantizerCallbackRegistry.register_()
class CVQVAECallback(UpdateMixin, BaseCallback):
    """Usage-probability EMA + nearest-token anchors + per-code decay blend (training only)."""

    def __init__(self, *args, anchor, eps: float = 1e-3, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self._anchor = anchor
        self._eps = eps

    @classmethod
    def build_pre_hook(cls, config, registry, item):
        config = super().build_pre_hook(config, registry, item)
        config['anchor'] = AnchorRegistry.build_or_return(config['anchor'])
        return config

    @property
    def needs_column_nearest(self) -> bool:  # type: ignore[override]
        return bool(self._anchor.needs_columns)     # NearestAnchor: yes; CachedAnchor samples rows, no distances

    @property
    def column_nearest_global(self) -> bool:  # type: ignore[override]
        return bool(self._anchor.sync)

    @property
    def needs_distance(self) -> bool:
        return bool(self._anchor.needs_distance) and self.quantizer.training

    @property
    def accepts_packed_keys(self) -> bool:  # type: ignore[override]
        return not self.quantizer.training     # after_encode is a no-op outside training

    def before_init_weights(self, config) -> None:
        super().before_init_weights(config)
        if not self.quantizer.training:
            return
        dev = self.vector_quantizer.embedding.weight.device
        self.quantizer.register_buffer('_probability', torch.zeros(self.quantizer.codebook_size, device=dev))

    @property
    def probability(self) -> torch.Tensor:
        return self.quantizer.get_buffer('_probability')

    @torch.no_grad()
    def after_encode(self, x, quant, memo):
        quant = super().after_encode(x, quant, memo)
        vq = self.vector_quantizer
        if not vq.training:
            return quant
        W = vq.embedding.weight.data
        K, D = W.shape
        x = x.detach().contiguous()
        N = x.shape[0]
        vq.protect_saved_codebook()
        world = parallel.world_size()

        def column_keys(counts: torch.Tensor):
            """The column arg-min, restricted to the codes whose anchor can get a non-zero blend weight this step: in
            fp32 the per-code decay is exactly 1 for every code used at more than ~2 % of the uniform rate, its anchor
            is then multiplied by exactly 0 (results are bit-identical to computing every anchor)."""
            if not self._anchor.needs_columns or '_column_ctx' not in memo['encode']:
                return memo['encode'].get('column_keys')
            rows = ops.cvq_needy_codes(self.probability, counts, world * N, decay=self._ema.decay, eps=self._eps)
            return vq.column_keys(memo['encode'], rows=rows)

        if getattr(self._anchor, 'peer_exchange', False):
            region = self._peer_region(x.device, [('W', (K, D), torch.float32), ('prob', (K,), torch.float32),
                                                  ('counts', (K + 1,), torch.int64), ('anchors', (K, D), torch.float32),
                                                  ('keys', (K,), torch.int64)])
            if region is not None:
                if self.probability.data_ptr() != region.prob.data_ptr():
                    region.prob.copy_(self.probability)
                    self.quantizer._buffers['_probability'] = region.prob
                # per-rank partials in the peer region: usage counts, nearest-token rows (+ their keys when the
                # anchors are global); ONE fused launch reduces them, updates `_probability` and blends the anchors
                # into the codebook of every replica
                ops.bincount_accumulate(quant, region.counts.zero_(), K, total_slot=True)
                col_keys = column_keys(region.counts)        # this rank's counts: a lower bound of the global ones
                self._anchor.gather_local(x, col_keys, N, out=region.anchors)
                if self._anchor.sync:
                    region.keys.copy_(col_keys)
                ops.comm_cvq_update(region, K, D, decay=self._ema.decay, eps=self._eps, minloc=self._anchor.sync)
                self._check_sync(region.W)
                return quant
        # [K counts | numel] in one int64 buffer -> one all-reduce (utils.py:35 does two)
        cnt = torch.zeros(K + 1, dtype=torch.int64, device=x.device)
        ops.bincount_accumulate(quant, cnt, K, total_slot=True)
        parallel.all_reduce_sum_(cnt)
        col_keys = column_keys(cnt)
        anchors = self._anchor.gather(x, col_keys, N, num_codes=K, distance=memo['encode'].get('distance'))
        scale = 1.0 if self._anchor.sync else 1.0 / world
        if self._anchor.sync:
            self._check_sync(anchors, 'anchors')
        ops.cvq_update(W, anchors, self.probability, cnt[:K], cnt[K:], decay=self._ema.decay, eps=self._eps,
                       anchor_scale=scale)
        self._check_sync(W)
        return quant
