"""CPU oracle for the codebook-quantization hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32) restatement of the reference's
quantizer arithmetic (magic-research/vector_quantization, `vq/algorithms`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it; the product package
`vector_quantization_b200` never does.

Pinning status
--------------
* The reference ships no tests, golden vectors or fixtures.
* Everything that lives in files under /root/reference is pinned by
  `oracle/make_golden.py`: it executes the reference's OWN hot-path source
  files (unmodified, imported from /root/reference under the minimal `todd`
  shim in `oracle/todd_shim/`) and checks this restatement against them
  bit-for-bit; the resulting vectors are committed under `tests/golden/`.
* PARITY UNPINNED at the `todd` boundary: `todd.utils.EMA/ema`,
  `todd.models.losses.MSELoss` and `todd.patches.torch.all_gather` live in the
  third-party package `todd_ai` (git commit ed2a3ae75a66, reference
  `.todd_version:1`, `setup.py:6-14`), which is absent from /root/reference and
  not installable offline.  Their published semantics are restated here
  (`ema`, `mse_loss`) and flagged UNVERIFIED; `ema_decay` is an explicit
  parameter (assumed default 0.99 as in upstream CVQ-VAE / BEiT-v2).

All citations are `file:line` relative to /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import einops
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# primitives
# --------------------------------------------------------------------------- #


def normalize(v: torch.Tensor) -> torch.Tensor:
    """`F.normalize` defaults (dim=1, eps=1e-12) — vq/algorithms/vq/callbacks/normalize.py:24,27."""
    return F.normalize(v)


def ema(old: torch.Tensor, new: torch.Tensor, decay) -> torch.Tensor:
    """todd.utils.ema / todd.utils.EMA.__call__ (UNVERIFIED, third-party):
    `old * decay + new * (1 - decay)`.  Call sites:
    vq/algorithms/cvqvae/quantizer_callback.py:94,102, vq/algorithms/vqkd/quantizers/callbacks.py:127."""
    return old * decay + new * (1 - decay)


def l2_distance(x: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """vq/algorithms/vq/distances.py:28-32 — `torch.cdist(x, e)` (true, sqrt distance)."""
    return torch.cdist(x, e)


def cosine_distance(x: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    """vq/algorithms/vq/distances.py:35-46 — `1 - normalize(x) @ normalize(e).T` via einsum."""
    x = F.normalize(x)
    e = F.normalize(e)
    return 1 - torch.einsum('x d, e d -> x e', x, e)


def distance(kind: str, x: torch.Tensor, e: torch.Tensor) -> torch.Tensor:
    if kind == 'L2':
        return l2_distance(x, e)
    if kind == 'Cosine':
        return cosine_distance(x, e)
    raise ValueError(kind)


def encode(kind: str, x: torch.Tensor, weight: torch.Tensor):
    """VectorQuantizer._encode — vq/algorithms/vq/quantizers.py:92-100.
    Returns (quant int64 [N], distance [N,K])."""
    d = distance(kind, x, weight.clone())          # :85 (`embeddings` clones), :97
    quant = torch.argmin(d, dim=-1)                # :99  (first minimum wins)
    return quant, d


def decode(weight: torch.Tensor, quant: torch.Tensor) -> torch.Tensor:
    """VectorQuantizer._decode — vq/algorithms/vq/quantizers.py:102-108 (`nn.Embedding`)."""
    return F.embedding(quant, weight)


def ste(z: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """vq/tasks/image_tokenization/models/quantizers/utils/ste.py:9-10."""
    return x + (z - x).detach()


def mse_loss(pred: torch.Tensor, target: torch.Tensor, norm: bool = False) -> torch.Tensor:
    """todd.models.losses.MSELoss(norm=..., reduction='mean', weight=1) (UNVERIFIED,
    third-party; built at vq/algorithms/vq/losses.py:37)."""
    if norm:
        pred = F.normalize(pred)
        target = F.normalize(target)
    return F.mse_loss(pred, target)


def codebook_loss(z, x, norm=False):
    """vq/algorithms/vq/losses.py:41-50."""
    return mse_loss(z, x.detach(), norm)


def commitment_loss(z, x, norm=False):
    """vq/algorithms/vq/losses.py:53-62."""
    return mse_loss(z.detach(), x, norm)


def vqgan_loss(z, x, beta=0.25, norm=False):
    """vq/algorithms/vq/losses.py:65-127 (beta default :73)."""
    return codebook_loss(z, x, norm) + beta * commitment_loss(z, x, norm)


def entropy_loss(d: torch.Tensor, temperature: float) -> torch.Tensor:
    """vq/algorithms/vq/losses.py:130-153 (registered, unused by shipped configs)."""
    flat = d.reshape(-1, d.shape[-1]) / temperature
    probs = flat.softmax(-1)
    log_probs = torch.log_softmax(flat + 1e-5, -1)
    avg_probs = probs.mean(0)
    avg_entropy = -torch.sum(avg_probs * torch.log(avg_probs + 1e-5))
    sample_entropy = -torch.mean(torch.sum(probs * log_probs, -1))
    return sample_entropy - avg_entropy


# --------------------------------------------------------------------------- #
# statistics / codebook updates
# --------------------------------------------------------------------------- #


def bin_count(quants: Sequence[torch.Tensor], K: int) -> torch.Tensor:
    """QuantStatistics.bin_count with sync=True over ranks — vq/algorithms/vq/utils.py:40-43,35
    (all_reduce SUM of per-rank `bincount(minlength=K)`)."""
    out = torch.zeros(K, dtype=torch.int64, device=quants[0].device)
    for q in quants:
        out += q.bincount(minlength=K)
    return out


def frequency(quants: Sequence[torch.Tensor], K: int) -> torch.Tensor:
    """QuantStatistics.frequency — vq/algorithms/vq/utils.py:48-52 (int64 / int64 → fp32)."""
    cnt = bin_count(quants, K)
    numel = torch.tensor(sum(q.numel() for q in quants), dtype=torch.int64, device=cnt.device)
    return cnt / numel


def kmeans_centroids(xs: Sequence[torch.Tensor], quants: Sequence[torch.Tensor],
                     weight: torch.Tensor) -> torch.Tensor:
    """VQKDCallback._kmeans (sync over ranks) — vq/algorithms/vqkd/quantizers/callbacks.py:44-71."""
    e = weight.clone()
    K, D = e.shape
    occ = bin_count(quants, K).unsqueeze(1)                       # :52-58
    cent = torch.zeros_like(e)
    for x, q in zip(xs, quants):                                  # per rank scatter, then all_reduce :60-64
        c = torch.zeros_like(e)
        c.scatter_add_(0, q.unsqueeze(1).expand(-1, D), x)
        cent += c
    occurred = occ > 0                                            # :66
    occ = occ.clamp_min(1)                                        # :67
    cent = cent / occ                                             # :69
    return cent.where(occurred, e)                                # :70


def vqkd_update(xs, quants, weight, decay):
    """VQKDCallback.after_encode (training) — vqkd/quantizers/callbacks.py:114-129 with
    `_update_embedding` normalising again (:73-75)."""
    xs = [normalize(x) for x in xs]                               # :124
    e = kmeans_centroids(xs, quants, weight)                      # :125
    e = normalize(e)                                              # :126
    e = ema(weight.clone(), e, decay)                             # :127
    return normalize(e)                                           # :128 → :73-75


def vqkd_lazy_init(x: torch.Tensor, weight: torch.Tensor, iters: int = 10, distance_kind: str = 'Cosine') -> torch.Tensor:
    """VQKDCallback.lazy_init_weights on a single rank — vq/algorithms/vqkd/quantizers/callbacks.py:77-112: seeds
    drawn with `random.sample` (the CALLER seeds `random`), `iters` rounds of normalise-codebook / assign /
    centroids (unused codes keep their row), final normalise (`_update_embedding`, :73-75).  Fewer tokens than
    codes: the RAW tokens overwrite the first rows (:91-92).  The CPU offload above 2^30 elements (:97-100)
    only moves tensors."""
    import random
    e = weight.clone()                                            # :87 (`embeddings` clones)
    if x.shape[0] < e.shape[0]:
        e[:x.shape[0]] = x                                        # :91-92
    else:
        x = normalize(x)                                          # :94
        indices = random.sample(range(x.shape[0]), e.shape[0])    # :101
        e = x[indices]                                            # :102
        for _ in range(iters):                                    # :103-106
            w = normalize(e)                                      # _update_embedding
            quant, _ = encode(distance_kind, x, w)
            e = kmeans_centroids([x], [quant], w)
    return normalize(e)                                           # :111


def vqgan_vqkd_update(weight: torch.Tensor, decay) -> torch.Tensor:
    """VQGAN_VQKDCallback.after_encode (training) — vq/algorithms/exp/vqgan_vqkd/quantizer_callback.py:124-134 with
    `_update_embedding` normalising again (:75-77): no statistics, the codebook is only pulled towards the sphere."""
    e = normalize(weight.clone())                                 # :131-132
    e = ema(weight.clone(), e, decay)                             # :133
    return normalize(e)                                           # :134 -> :75-77


def multinomial_anchor(x: torch.Tensor, d: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """MultinomialAnchor._anchors — vq/algorithms/cvqvae/anchors.py:88-104 (consumes the torch RNG)."""
    indices = einops.rearrange(d, 'x e -> e x').softmax(1).multinomial(1)
    indices = einops.rearrange(indices, 'e 1 -> e')
    return x[indices], indices


def nearest_anchor(x: torch.Tensor, d: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """NearestAnchor._anchors — vq/algorithms/cvqvae/anchors.py:71-85. Returns (anchors, indices)."""
    idx = d.argmin(0)
    return x[idx], idx


def cvq_update(xs, ds, quants, weight, prob, decay, eps=1e-3, anchor_sync=False):
    """CVQVAECallback.after_encode (training) — vq/algorithms/cvqvae/quantizer_callback.py:75-105
    with BaseAnchor.forward — vq/algorithms/cvqvae/anchors.py:41-68.

    xs/ds/quants are per-rank lists; returns (new_weight, new_prob, anchors, anchor_indices).
    `prob` is the `_probability` buffer (identical on all ranks: it is an EMA of an all-reduced
    frequency starting from zeros, quantizer_callback.py:60-66).
    """
    K = weight.shape[0]
    e = weight.clone()                                            # :85
    freq = frequency(quants, K)                                   # :88-93 (sync=True)
    p = ema(prob, freq, decay)                                    # :94
    world = len(xs)
    if anchor_sync:                                               # anchors.py:50-57
        x = torch.cat(list(xs))
        d = torch.cat(list(ds))
        pp = torch.stack([p] * world).mean(0)                     # all ranks hold the same p
        anchors, idx = nearest_anchor(x, d)
        p_for_decay = p
        del pp
    else:
        per_rank = [nearest_anchor(x, d) for x, d in zip(xs, ds)]
        anchors = per_rank[0][0].clone()
        for a, _ in per_rank[1:]:
            anchors = anchors + a                                 # all_reduce SUM :66
        if world > 1:
            anchors = anchors / world                             # :67
        idx = torch.stack([i for _, i in per_rank])
        p_for_decay = p
    dec = 1 - torch.exp(-p_for_decay.unsqueeze(1) * K * 10 / (1 - decay) - eps)   # :98-101
    new_w = ema(e, anchors, dec)                                  # :102
    return new_w, p, anchors, idx


# --------------------------------------------------------------------------- #
# whole-quantizer forward (BaseQuantizer.forward template)
# --------------------------------------------------------------------------- #


@dataclass
class QuantizerSpec:
    """Mirror of the reference config keys that change arithmetic."""
    distance: str = 'L2'                    # configs/vq/distance.py:7
    callback: str | None = None             # None | 'NormalizeCallback' | 'VQKDCallback' | 'CVQVAECallback'
    ema_decay: float = 0.99                 # todd EMA default (UNVERIFIED)
    cvq_eps: float = 1e-3                   # quantizer_callback.py:32
    anchor_sync: bool = False               # configs/cvqvae/quantizer.py:4 / configs/cluster/model.py:28
    losses: dict = field(default_factory=dict)   # name -> dict(type=..., beta=?, norm=?)
    training: bool = True
    normalize: bool = False                 # a NormalizeCallback layered in front of `callback` (llamagen + cvqvae mixin)


def _losses(spec: QuantizerSpec, z, x):
    out = {}
    for name, cfg in spec.losses.items():
        t = cfg['type']
        norm = cfg.get('norm', False)
        if t == 'CodebookLoss':
            out[name] = codebook_loss(z, x, norm)
        elif t == 'CommitmentLoss':
            out[name] = commitment_loss(z, x, norm)
        elif t == 'VQGANLoss':
            out[name] = vqgan_loss(z, x, cfg.get('beta', 0.25), norm)
        else:
            raise ValueError(t)
    return out


def quantizer_forward(spec: QuantizerSpec, xs_in: Sequence[torch.Tensor], weight: torch.Tensor,
                      prob: torch.Tensor | None = None):
    """One `VectorQuantizer.forward` per rank with the reference's ordering
    (vq/tasks/image_tokenization/models/quantizers/base.py:123-182, vq/algorithms/vq/quantizers.py:110-117):
    before_encode → _encode → after_encode (codebook update, collective) → decode (UPDATED codebook)
    → losses → STE.  `xs_in` is the per-rank list of token matrices (len 1 = single process);
    tensors may require grad.

    Returns dict with per-rank lists `x`, `quant`, `z`, `z_ste`, `loss`, `losses`, `distance`, and the
    post-step `weight`, `prob`.
    """
    W = weight.detach().clone()
    xs = list(xs_in)
    # ---- before_encode ---------------------------------------------------------------------
    if spec.callback in ('NormalizeCallback', 'VQKDCallback') or spec.normalize:
        xs = [normalize(x) for x in xs]                           # normalize.py:24
        W = normalize(W)                                          # normalize.py:26-28
        if spec.callback == 'VQKDCallback':
            W = normalize(W)                                      # callbacks.py:73-75 (normalises again)
    # ---- _encode ----------------------------------------------------------------------------
    enc = [encode(spec.distance, x.detach(), W) for x in xs]
    quants = [q for q, _ in enc]
    ds = [d for _, d in enc]
    # ---- after_encode (training only) -------------------------------------------------------
    W_used = W
    new_prob = prob
    anchors = anchor_idx = None
    if spec.training and spec.callback == 'VQKDCallback':
        W_used = vqkd_update([x.detach() for x in xs], quants, W, spec.ema_decay)
    elif spec.training and spec.callback == 'CVQVAECallback':
        assert prob is not None
        W_used, new_prob, anchors, anchor_idx = cvq_update(
            [x.detach() for x in xs], ds, quants, W, prob, spec.ema_decay, spec.cvq_eps,
            spec.anchor_sync)
    # ---- decode / loss / STE ----------------------------------------------------------------
    Wp = W_used.clone().requires_grad_(True)                      # leaf so codebook grads can be checked
    out = dict(x=xs, quant=quants, distance=ds, z=[], z_ste=[], loss=[], losses=[], weight=W_used,
               weight_leaf=Wp, prob=new_prob, anchors=anchors, anchor_idx=anchor_idx)
    for x, q in zip(xs, quants):
        z = decode(Wp, q)
        ls = _losses(spec, z, x)
        loss = sum(ls.values(), x.new_zeros([]))                  # base.py:157-159
        out['z'].append(z)
        out['losses'].append(ls)
        out['loss'].append(loss)
        out['z_ste'].append(ste(z, x))                            # quantizers.py:116
    return out


# --------------------------------------------------------------------------- #
# FSQ  (vq/algorithms/fsq/quantizers.py)
# --------------------------------------------------------------------------- #


class FSQ:
    """FiniteScalarQuantizer + BaseConverter — vq/algorithms/fsq/quantizers.py:19-150."""

    def __init__(self, levels: Sequence[int], eps: float = 1e-3):
        levels = tuple(int(v) for v in levels)
        self.levels = levels
        self.eps = eps                                                          # :80
        self.cumprod = torch.tensor((1,) + levels[:-1]).cumprod(0)              # :30  int64
        self.max_per_digit = torch.tensor(levels, dtype=torch.int)              # :33
        self.codebook_size = int(self.max_per_digit.prod().item())              # :41
        quant = torch.arange(self.codebook_size)
        self.embeddings = self.from_decimal(quant) / (self.max_per_digit // 2) - 1   # :90-93

    def from_decimal(self, x):                                                  # :59-63
        x = x.unsqueeze(-1)
        x = x // self.cumprod
        return x % self.max_per_digit

    def to_decimal(self, x):                                                    # :65-68
        return (x * self.cumprod).sum(-1).to(torch.int)

    def constants(self):
        """The per-channel constants of `_encode` exactly as torch computes them (:114-115,:122)."""
        max_int = self.max_per_digit - 1
        max_ = max_int * (1 - self.eps)
        odd = max_int % 2
        shift = torch.atanh(odd / max_)
        half = self.max_per_digit // 2
        return max_, odd, shift, half

    def encode(self, x: torch.Tensor):
        """`_encode` :108-126.  Returns (quant int32 [N], z [N,D], pre-round value)."""
        max_, odd, _, half = self.constants()
        z = torch.tanh(x + torch.atanh(odd / max_)) * max_ - odd                # :118
        z = z / 2                                                               # :119
        pre = z
        z = ste(z.round(), z)                                                   # :120
        zq = z / half                                                           # :123
        quant = self.to_decimal(z + half)                                       # :124-125
        return quant, zq, pre

    def decode(self, quant: torch.Tensor) -> torch.Tensor:
        """decode-only branch of `_decode` :136-137."""
        digits = self.from_decimal(quant)
        return digits / (self.max_per_digit // 2) - 1

    def forward(self, x: torch.Tensor):
        """BaseQuantizer.forward for FSQ: loss = x.new_zeros([]) (base.py:159)."""
        quant, zq, pre = self.encode(x)
        return zq, x.new_zeros([]), quant, pre


# --------------------------------------------------------------------------- #
# usage metrics (vq/tasks/image_tokenization/runners/metrics.py:25-73)
# --------------------------------------------------------------------------- #


def model_quantize(spec: 'QuantizerSpec', x_nchw: torch.Tensor, weight: torch.Tensor, **kwargs):
    """`BaseModel.quantize` (vq/tasks/image_tokenization/models/base.py:116-129): rearrange
    'b c h w -> (b h w) c', quantizer forward, rearrange '(b h w) c -> b c h w' + contiguous."""
    b, c, h, w = x_nchw.shape
    rows = einops.rearrange(x_nchw, 'b c h w -> (b h w) c')
    out = quantizer_forward(spec, [rows], weight, **kwargs)
    z = einops.rearrange(out['z_ste'][0], '(b h w) c -> b c h w', b=b, c=c, h=h, w=w).contiguous()
    return z, out['loss'][0], dict(out, x_shape=(b, c, h, w))


def model_encode_to_quant(kind: str, x_nchw: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """`BaseModel.encode_to_quant` after the encoder (base.py:131-146), without callbacks:
    tokens [b, h, w] int64."""
    b, _, h, w = x_nchw.shape
    quant, _ = encode(kind, einops.rearrange(x_nchw, 'b c h w -> (b h w) c'), weight)
    return einops.rearrange(quant, '(b h w) -> b h w', b=b, h=h, w=w)


def codebook_usage(counts: torch.Tensor) -> float:
    """CodebookUsageMetric._summary :63."""
    return counts.bool().sum().item() / counts.numel()


def codebook_ppl(counts: torch.Tensor) -> float:
    """CodebookPPLMetric._summary :70-73 — categorical ENTROPY in nats (not its exp)."""
    p = counts / counts.sum()
    return torch.distributions.Categorical(p).entropy().item()


# --------------------------------------------------------------------------- #
# near-tie analysis helpers used by the parity tests
# --------------------------------------------------------------------------- #


def index_mismatch_report(d_oracle: torch.Tensor, q_oracle: torch.Tensor, q_test: torch.Tensor):
    """For rows where q_test != q_oracle, return the oracle's distance gap
    d[row, q_test] - d[row, q_oracle] (>= 0).  A mismatch is an accepted near-tie iff gap < eps."""
    rows = (q_oracle != q_test).nonzero().flatten()
    if rows.numel() == 0:
        return rows, d_oracle.new_zeros(0)
    gap = d_oracle[rows, q_test[rows]] - d_oracle[rows, q_oracle[rows]]
    return rows, gap


def synthetic_latents(N: int, K: int, D: int, seed: int = 3407, sigma: float = 0.5,
                      normalized_codebook: bool = False, clustered: bool = True):
    """Seeded synthetic inputs: trained-like codebook E0 ~ N(0,1) and latents
    x = E0[pi(n)] + sigma*std(E0)*eps (clustered) or x ~ N(0,1)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    E = torch.randn(K, D, generator=g)
    if normalized_codebook:
        E = F.normalize(E)
    if clustered:
        pi = torch.randint(0, K, (N,), generator=g)
        std = E.std() if E.numel() > 1 else torch.tensor(1.0)   # std of a single element is NaN
        x = E[pi] + sigma * std * torch.randn(N, D, generator=g)
    else:
        x = torch.randn(N, D, generator=g)
    return x, E
