"""Compatibility-mode components: the on-demand distance matrix and its consumers
(EntropyLoss, MultinomialAnchor, `materialize_distance`), the public `loss()` / `_loss()` of the template, the
hook-honouring unfused forward, VQGAN_VQKDCallback, and the streaming token writer — against the oracle."""
import pathlib

import pytest
import torch
import torch.nn.functional as F

import vector_quantization_b200 as vqb
from oracle import oracle as O
from vector_quantization_b200 import functional as Fq
from vector_quantization_b200 import ops

pytestmark = pytest.mark.gpu


def emb(K, D):
    return dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D)


@pytest.mark.parametrize('metric', ['L2', 'Cosine'])
@pytest.mark.parametrize('N,K,D', [(300, 130, 8), (1024, 512, 32), (257, 1000, 256), (64, 64, 5)])
def test_distance_matrix_forward_and_backward_vs_oracle(dev, metric, N, K, D):
    """`vqb_distance_matrix` == torch.cdist / 1 - cos of the oracle (1e-5), and the gradients of a scalar function of
    the matrix w.r.t. tokens and codebook equal autograd through the oracle's formula (1e-4)."""
    x, E = O.synthetic_latents(N, K, D, seed=N + K)
    g = torch.Generator().manual_seed(1)
    G = torch.randn(N, K, generator=g)
    # the oracle's formula evaluated in float64: an fp32 CPU cdist is itself only good to ~1e-6 * (|x|^2 + |e|^2) / d
    xo, Eo = x.double().requires_grad_(True), E.double().requires_grad_(True)
    d_ref = O.distance(metric, xo, Eo)
    (d_ref * G.double()).sum().backward()
    xg, Eg = x.to(dev).requires_grad_(True), E.to(dev).requires_grad_(True)
    d = Fq.distance_matrix(xg, Eg, metric)
    (d * G.to(dev)).sum().backward()
    torch.testing.assert_close(d.detach().cpu(), d_ref.detach().float(), rtol=1e-5, atol=2e-5)
    far = (d_ref.detach() > 1e-3)              # cdist's gradient is singular at d = 0
    assert far.float().mean() > 0.99
    torch.testing.assert_close(xg.grad.cpu(), xo.grad.float(), rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(Eg.grad.cpu(), Eo.grad.float(), rtol=2e-4, atol=2e-4)
    # the distance modules return the same matrix (what `quantizer.distance(x, e)` gives user code)
    mod = vqb.L2Distance() if metric == 'L2' else vqb.CosineDistance()
    assert torch.equal(mod(x.to(dev), E.to(dev)), d.detach())


def _quantizer(cfg, K, D, dev, training=True):
    q = vqb.build_quantizer(dict(cfg, embedding=emb(K, D)), training=training).to(dev)
    q._forward_pre_hooks.clear()
    return q


def test_entropy_loss_value_and_gradients_vs_oracle(dev):
    """VQGAN quantizer + EntropyLoss (losses.py:130-153): loss values, token gradient and codebook gradient equal
    the oracle (codebook/commitment MSE + entropy of softmax(d / T)), and memo['encode']['distance'] is there."""
    N, K, D, T = 512, 96, 16, 0.5
    x, E = O.synthetic_latents(N, K, D, seed=2)
    cfg = dict(type='VQGANQuantizer', distance=dict(type='L2Distance'),
               losses=dict(vqgan_loss=dict(type='VQGANLoss'), entropy_loss=dict(type='EntropyLoss', temperature=T)),
               init_weights=dict(type='vqgan'))
    q = _quantizer(cfg, K, D, dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    xg = x.to(dev).requires_grad_(True)
    z, loss, memo = q(xg, dict())
    loss.backward()
    xo = x.clone().requires_grad_(True)
    Wo = E.clone().requires_grad_(True)
    d = O.distance('L2', xo, Wo)
    quant = d.argmin(1)
    zo = O.decode(Wo, quant)
    want = O.vqgan_loss(zo, xo) + O.entropy_loss(d, T)
    want.backward()
    assert torch.equal(memo['quant'].cpu(), quant)
    torch.testing.assert_close(memo['encode']['distance'].detach().cpu(), d.detach(), rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(memo['loss']['entropy_loss'].detach().cpu(), O.entropy_loss(d, T).detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(loss.detach().cpu(), want.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(xg.grad.cpu(), xo.grad, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(q.embedding.weight.grad.cpu(), Wo.grad, rtol=1e-3, atol=1e-6)
    assert 'distance' not in memo['loss']


def test_materialize_distance_switch_for_user_callbacks(dev):
    """`materialize_distance=True` (config key) puts the reference's memo['encode']['distance'] back for user code;
    indices are its row arg-min; off by default."""
    N, K, D = 256, 64, 8
    x, E = O.synthetic_latents(N, K, D, seed=8, normalized_codebook=True)
    base = dict(type='VQGANQuantizer', distance=dict(type='CosineDistance'), callbacks=[dict(type='NormalizeCallback')],
                losses=dict(l=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
    q = _quantizer(dict(base, materialize_distance=True), K, D, dev, training=False)
    q0 = _quantizer(base, K, D, dev, training=False)
    for m in (q, q0):
        with torch.no_grad():
            m.embedding.weight.copy_(E)
    z, loss, memo = q(x.to(dev), dict())
    z0, loss0, memo0 = q0(x.to(dev), dict())
    assert 'distance' not in memo0['encode']
    d = memo['encode']['distance']
    torch.testing.assert_close(d.cpu(), O.distance('Cosine', x, E), rtol=1e-5, atol=2e-5)
    assert torch.equal(memo['quant'], memo0['quant']) and (d.argmin(1) == memo['quant']).float().mean() > 0.99
    torch.testing.assert_close(z, z0, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(loss, loss0, rtol=1e-6, atol=1e-8)


def test_multinomial_anchor_samples_from_softmax_of_the_distance_column(dev):
    """MultinomialAnchor (anchors.py:88-104): P(token n for code k) = softmax_n d[n, k].  With one token much
    FARTHER than all others (the reference samples the softmax of the DISTANCE, not of its negative) that token is
    the anchor of every code; and on generic data the empirical frequencies follow the softmax."""
    N, K, D = 64, 8, 4
    g = torch.Generator().manual_seed(3)
    E = torch.randn(K, D, generator=g)
    x = 0.1 * torch.randn(N, D, generator=g)
    x[17] = 50.0                                                  # distance ~100 to every code: softmax -> one-hot
    anchor = vqb.MultinomialAnchor()
    d = Fq.distance_matrix(x.to(dev), E.to(dev), 'L2')
    a = anchor.gather(x.to(dev), None, N, num_codes=K, distance=d)
    assert torch.equal(a.cpu(), x[17].expand(K, D))
    # frequencies
    x = torch.randn(N, D, generator=g)
    d = Fq.distance_matrix(x.to(dev), E.to(dev), 'L2')
    p = d.t().softmax(1).cpu()
    counts = torch.zeros(K, N)
    torch.manual_seed(0)
    trials = 400
    for _ in range(trials):
        rows = anchor.gather(x.to(dev), None, N, num_codes=K, distance=d).cpu()
        idx = (rows[:, None, :] == x[None, :, :]).all(-1).float().argmax(1)
        counts[torch.arange(K), idx] += 1
    assert ((counts / trials - p).abs() < 5 * (p * (1 - p) / trials).sqrt() + 0.01).all()


def test_cvqvae_training_step_with_multinomial_anchor(dev):
    """A CVQ-VAE step with MultinomialAnchor: every updated code row is the blend of its old row and ONE token row,
    with the oracle's per-code decay."""
    N, K, D = 256, 32, 8
    _, E = O.synthetic_latents(N, K, D, seed=6)
    g = torch.Generator().manual_seed(6)
    x = E[torch.randint(0, 8, (N,), generator=g)] + 0.05 * torch.randn(N, D, generator=g)   # only 8 of the 32 codes are used
    cfg = dict(type='VQGANQuantizer', distance=dict(type='L2Distance'),
               callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='MultinomialAnchor'))],
               losses=dict(l=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
    q = _quantizer(cfg, K, D, dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    z, loss, memo = q(x.to(dev).requires_grad_(True), dict())
    W = q.embedding.weight.detach().cpu()
    quant, _ = O.encode('L2', x, E)
    p = O.ema(torch.zeros(K), O.frequency([quant], K), 0.99)
    dec = (1 - torch.exp(-p * K * 10 / (1 - 0.99) - 1e-3)).unsqueeze(1)
    anchors = (W - E * dec) / (1 - dec)                             # solve the blend for the anchor row
    light = (1 - dec).flatten() > 0.5                               # rarely used codes: the anchor dominates the blend
    assert light.sum() >= 4
    nearest = torch.cdist(anchors[light].double(), x.double()).min(1).values
    assert (nearest < 1e-4 * x.norm(dim=1).max()).all(), 'every anchor must be one of the token rows'
    heavy = ~light                                                  # frequently used codes barely move
    torch.testing.assert_close(W[heavy], (E * dec)[heavy], rtol=0, atol=float((1 - dec)[heavy].max() * x.abs().max()) + 1e-6)
    torch.testing.assert_close(q.get_buffer('_probability').cpu(), p, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('norm', [False, True])
def test_public_loss_api_on_arbitrary_tensors(dev, norm):
    """`quantizer.loss(z, x, memo)` / `_loss` (base.py:151-171) standalone: values and BOTH gradients (codebook role ->
    z, commitment role -> x) equal the oracle's mse terms."""
    N, K, D = 300, 32, 16
    g = torch.Generator().manual_seed(4)
    z0, x0 = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g)
    cfg = dict(type='VQGANQuantizer', distance=dict(type='L2Distance'),
               losses=dict(a=dict(type='VQGANLoss', beta=0.4, codebook=dict(mse=dict(norm=norm)), commitment=dict(mse=dict(norm=norm))),
                           b=dict(type='CommitmentLoss', mse=dict(norm=norm))), init_weights=dict(type='vqgan'))
    q = _quantizer(cfg, K, D, dev)
    zg, xg = z0.to(dev).requires_grad_(True), x0.to(dev).requires_grad_(True)
    loss, memo = q.loss(zg, xg, dict())
    loss.backward()
    zo, xo = z0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
    want = O.vqgan_loss(zo, xo, 0.4, norm) + O.commitment_loss(zo, xo, norm)
    want.backward()
    torch.testing.assert_close(loss.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-7)
    assert set(memo['loss']) == {'a', 'b'}
    torch.testing.assert_close(zg.grad.cpu(), zo.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(xg.grad.cpu(), xo.grad, rtol=1e-4, atol=1e-7)


class _ScaleLoss(vqb.BaseCallback):
    """A user callback on the loss / decode hooks (the fused path cannot honour these: the template path runs)."""

    def after_decode(self, z, memo):
        memo['seen_after_decode'] = True
        return z

    def after_loss(self, loss, memo):
        return loss * 2.0


def test_hooked_template_forward_equals_fused_forward(dev):
    """Callbacks that override decode / loss hooks route forward() through the reference's unfused template; with a
    neutral decode hook and a loss-doubling hook the results equal the fused path (z, indices, gradients) and twice
    its loss."""
    N, K, D = 512, 64, 16
    x, E = O.synthetic_latents(N, K, D, seed=9)
    cfg = dict(type='VQGANQuantizer', distance=dict(type='L2Distance'), losses=dict(l=dict(type='VQGANLoss')),
               init_weights=dict(type='vqgan'))
    fused = _quantizer(cfg, K, D, dev)
    hooked = _quantizer(dict(cfg, callbacks=[_ScaleLoss()]), K, D, dev)
    outs = []
    for q in (fused, hooked):
        with torch.no_grad():
            q.embedding.weight.copy_(E)
        xg = x.to(dev).requires_grad_(True)
        z, loss, memo = q(xg, dict())
        (loss + z.sum()).backward()
        outs.append((z.detach(), loss.detach(), memo, xg.grad, q.embedding.weight.grad))
    (z0, l0, m0, gx0, gW0), (z1, l1, m1, gx1, gW1) = outs
    assert m1.get('seen_after_decode') and torch.equal(m0['quant'], m1['quant'])
    torch.testing.assert_close(z1, z0, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(l1, 2 * l0, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(gx1 - 1, 2 * (gx0 - 1), rtol=1e-4, atol=1e-6)   # d(z.sum())/dx = 1 on both paths
    torch.testing.assert_close(gW1, 2 * gW0, rtol=1e-4, atol=1e-7)


def test_vqgan_vqkd_callback_step(dev):
    """VQGAN_VQKDCallback (exp/vqgan_vqkd/quantizer_callback.py:124-134): per training step the normalised codebook is
    pulled to the sphere, W <- normalize(ema(W, normalize(W))); the codebook still receives its loss gradient."""
    N, K, D = 256, 64, 8
    x, E = O.synthetic_latents(N, K, D, seed=10)
    cfg = dict(type='VQGANQuantizer', distance=dict(type='L2Distance'), callbacks=[dict(type='VQGAN_VQKDCallback', ema=dict())],
               losses=dict(l=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
    q = _quantizer(cfg, K, D, dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    z, loss, memo = q(x.to(dev).requires_grad_(True), dict())
    loss.backward()
    W1 = F.normalize(E)                                           # NormalizeCallback.before_encode
    torch.testing.assert_close(q.embedding.weight.detach().cpu(), O.vqgan_vqkd_update(W1, 0.99), rtol=1e-5, atol=1e-6)
    quant, _ = O.encode('L2', F.normalize(x), W1)
    assert (memo['quant'].cpu() != quant).sum() <= 2
    assert q.embedding.weight.grad is not None and float(q.embedding.weight.grad.abs().sum()) > 0


def test_token_stream_writer_files_equal_save_tokens(dev, tmp_path):
    """TokenStreamWriter == TokenizeCallback's files ({iter}_{rank}.pth, `Tokens` dict, int64 [b, h, w]) for compact
    GPU ids, written asynchronously."""
    K = 500
    g = torch.Generator().manual_seed(0)
    batches = [torch.randint(0, K, (2 * 4 * 4,), generator=g) for _ in range(6)]
    with vqb.tokenizer.TokenStreamWriter(tmp_path / 'tokens', rank=1, depth=2) as w:
        for it, q in enumerate(batches):
            w.write(it, [f'id{it}a', f'id{it}b'], torch.tensor([it, it + 1]), q.to(torch.int32).to(torch.uint16).to(dev),
                    (2, 8, 4, 4))
    for it, q in enumerate(batches):
        rec = vqb.tokenizer.load_tokens(tmp_path / 'tokens' / f'{it}_1.pth')
        vqb.tokenizer.save_tokens(tmp_path / 'ref.pth', [f'id{it}a', f'id{it}b'], torch.tensor([it, it + 1]), q, (2, 8, 4, 4))
        want = vqb.tokenizer.load_tokens(tmp_path / 'ref.pth')
        assert rec['id_'] == want['id_'] and torch.equal(rec['category'], want['category'])
        assert rec['tokens'].dtype == torch.int64 and torch.equal(rec['tokens'], want['tokens'])
