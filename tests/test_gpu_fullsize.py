"""Parity at BASELINE.json's FULL sizes.

cfg 1-4 are compared DIRECTLY with the CPU oracle at their stated size and dtype (one whole quantizer step through
the drop-in module: indices under the near-tie policy, loss, updated codebook, `_probability`, z, token gradient).
On top of that come size-independent properties: identity on the codebook itself, idempotence, permutation
equivariance, agreement of the tensor-core kernel with the CUDA-core cross-check kernel, and shard-combine
associativity (cfg 5's 68.7 GB distance matrix cannot be materialised by any oracle).
"""
import pytest
import torch
import torch.nn.functional as F

import vector_quantization_b200 as vqb
from oracle import oracle as O
from vector_quantization_b200 import ops

pytestmark = pytest.mark.gpu


def emb(K, D):
    return dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D)


def test_cfg1_vqgan_full_size_vs_oracle(dev):
    """BASELINE.json configs[0]: codebook 8192 x 256 fp32, batch 64 x 16 x 16 latents, forward."""
    N, K, D = 64 * 16 * 16, 8192, 256
    x, E = O.synthetic_latents(N, K, D, seed=3407)
    spec = O.QuantizerSpec(distance='L2', losses={'vqgan_loss': dict(type='VQGANLoss')})
    out = O.quantizer_forward(spec, [x], E)
    q = vqb.build_quantizer(dict(type='VQGANQuantizer', embedding=emb(K, D), distance=dict(type='L2Distance'),
                                 losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))).to(dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    z, loss, memo = q(x.to(dev), dict())
    quant = memo['quant'].cpu()
    rows, gap = O.index_mismatch_report(out['distance'][0], out['quant'][0], quant)
    assert (gap <= 1e-5 * out['distance'][0][rows, out['quant'][0][rows]].clamp_min(1)).all()
    assert rows.numel() <= N // 1000
    same = quant == out['quant'][0]
    assert torch.equal(z.detach().cpu()[same], out['z_ste'][0].detach()[same])       # bit-exact x + (W[q] - x)
    torch.testing.assert_close(loss.detach().cpu(), out['loss'][0].detach(), rtol=1e-5, atol=1e-7)


def test_cfg1_degenerate_vqgan_init_near_tie_policy(dev):
    """The reference's real init U(-1/K, 1/K) makes every distance agree to ~1e-4 relative:
    two fp32 formulations already disagree on ~5 % of the indices.  Every disagreement must be a near-tie."""
    N, K, D = 4096, 8192, 256
    g = torch.Generator().manual_seed(1)
    E = (torch.rand(K, D, generator=g) * 2 - 1) / K
    x = torch.randn(N, D, generator=g)
    q_ref, d = O.encode('L2', x, E)
    a = ops.pack_rows(x.to(dev))
    b = ops.pack_rows(E.to(dev), want_half_sqnorm=True)
    keys = ops.new_keys(N, dev)
    ops.assign(a, b, keys, l2=True)
    quant = ops.unpack_keys(keys).cpu()
    rows, gap = O.index_mismatch_report(d, q_ref, quant)
    assert (gap <= 1e-5 * d[rows, q_ref[rows]].clamp_min(1)).all(), float(gap.max())


@pytest.mark.parametrize('K,D,metric', [(8192, 32, 'Cosine'), (16384, 8, 'L2'), (8192, 256, 'Cosine')])
def test_identity_and_idempotence_at_full_codebook_size(dev, K, D, metric):
    """Quantizing the codebook rows themselves returns arange(K) and zero loss; quantizing the output again
    returns the same indices (cfg 2 / 3 / 4 codebook sizes)."""
    g = torch.Generator().manual_seed(K + D)
    E = torch.randn(K, D, generator=g)
    if metric == 'Cosine':
        E = F.normalize(E)
    q = vqb.build_quantizer(dict(type='VQGANQuantizer', embedding=emb(K, D), distance=dict(type=f'{metric}Distance'),
                                 losses=dict(l=dict(type='CommitmentLoss')), init_weights=dict(type='vqgan'))).to(dev)
    q.eval()
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    z, loss, memo = q(E.to(dev), dict())
    assert torch.equal(memo['quant'].cpu(), torch.arange(K))
    assert float(loss.detach()) == 0.0 and torch.equal(z.detach().cpu(), E)
    z2, loss2, memo2 = q(z.detach(), dict())
    assert torch.equal(memo2['quant'], memo['quant']) and float(loss2.detach()) == 0.0


def test_cfg2_full_size_backends_agree_and_permutation_equivariance(dev):
    """cfg 2 at full size (65 536 x 8192 x 32, bf16 tokens, fp32 unit-norm codebook): the tcgen05 kernel and the
    CUDA-core cross-check kernel pick the same codes (scores equal to fp32 rounding), and permuting the tokens
    permutes the result."""
    N, K, D = 65536, 8192, 32
    g = torch.Generator().manual_seed(2)
    E = F.normalize(torch.randn(K, D, generator=g))
    x = (E[torch.randint(0, K, (N,), generator=g)] + 0.09 * torch.randn(N, D, generator=g)).to(torch.bfloat16).to(dev)
    book = ops.pack_rows(E.to(dev), normalize=True)
    tok = ops.as_operand(x)
    res = {}
    for name, backend in (('tc', ops.BACKEND_TCGEN05), ('simt', ops.BACKEND_SIMT)):
        keys = ops.new_keys(N, dev)
        ops.assign(tok, book, keys, l2=False, backend=backend)
        res[name] = ops.unpack_keys(keys, want_score=True)
    diff = res['tc'][0] != res['simt'][0]
    assert diff.float().mean() < 1e-3
    torch.testing.assert_close(res['tc'][1], res['simt'][1], rtol=1e-5, atol=2e-6)
    # both candidates of a disagreeing row have the same score to fp32 rounding
    assert (res['tc'][1][diff] - res['simt'][1][diff]).abs().max().item() < 2e-6 if diff.any() else True
    perm = torch.randperm(N, generator=g).to(dev)
    keys = ops.new_keys(N, dev)
    ops.assign(ops.as_operand(x[perm].contiguous()), book, keys, l2=False)
    assert torch.equal(ops.unpack_keys(keys), res['tc'][0][perm])


def test_cfg5_shape_shard_combine_is_associative(dev):
    """cfg 5 shape (768-d CLIP-sized features): arg-min over 4 codebook shards combined through the packed
    min-loc keys equals the arg-min over the whole codebook (what the 8-GPU min-loc all-reduce computes)."""
    N, K, D = 4096, 16384, 768
    g = torch.Generator().manual_seed(4)
    E = torch.randn(K, D, generator=g).to(dev)
    x = torch.randn(N, D, generator=g).to(torch.bfloat16).to(dev)
    tok = ops.as_operand(x)
    whole = ops.new_keys(N, dev)
    ops.assign(tok, ops.pack_rows(E, normalize=True, planes=1), whole, l2=False)
    sharded = ops.new_keys(N, dev)
    for r in range(4):
        lo, hi = r * K // 4, (r + 1) * K // 4
        ops.assign(tok, ops.pack_rows(E[lo:hi].contiguous(), normalize=True, planes=1), sharded, l2=False, index_offset=lo)
    assert torch.equal(whole, sharded)
    # and it is the true nearest code under the same bf16-plane arithmetic (fp32 check on a sample)
    idx = ops.unpack_keys(whole)[:256].cpu()
    ref = (F.normalize(x[:256].float()) @ F.normalize(E).to(torch.bfloat16).float().t()).argmax(1).cpu()
    assert (idx == ref).float().mean() > 0.98


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('D', [32, 64, 16])
def test_gather_and_backward_geometry_independent_of_row_count(dev, dtype, D):
    """Above ~57k rows the gather / backward kernels switch to half the lanes per row (one wave instead of 1.15):
    the per-row results must be bit-identical to the small-N geometry (run on two halves), the loss equal to
    reduction-order rounding."""
    N, K = 70000, 1024
    g = torch.Generator().manual_seed(D)
    x = torch.randn(N, D, generator=g).to(dtype).to(dev)
    W = F.normalize(torch.randn(K, D, generator=g)).to(dev)
    q = torch.randint(0, K, (N,), generator=g).to(dev)
    gz = torch.randn(N, D, generator=g).to(dev)
    g4 = torch.tensor([0.3, 0.7, 1.0, 0.25], device=dev)
    z, mse4, _, xn = ops.gather_ste_loss(x, W, quant=q, normalize_x=True, want_norm=True, want_xnorm=True)
    gx, gW = ops.quantize_backward(gz, x, W, q, g4, normalize_x=True, want_norm=True, need_gW=True)
    h = N // 2
    zs, xns, gxs, m4 = [], [], [], []
    gW2 = torch.zeros_like(W)
    for lo, hi in ((0, h), (h, N)):
        z_, m_, _, xn_ = ops.gather_ste_loss(x[lo:hi].contiguous(), W, quant=q[lo:hi].contiguous(), normalize_x=True,
                                             want_norm=True, want_xnorm=True)
        # the loss scale 2/(N D) of a half is twice the whole's: halve the upstream loss gradients
        gx_, gW_ = ops.quantize_backward(gz[lo:hi].contiguous(), x[lo:hi].contiguous(), W, q[lo:hi].contiguous(),
                                         g4 * ((hi - lo) / N), normalize_x=True, want_norm=True, need_gW=True)
        zs.append(z_); xns.append(xn_); gxs.append(gx_); m4.append(m_ * ((hi - lo) / N)); gW2 += gW_
    # the row norm is summed over a different lane split: equal to fp32 rounding, not bit for bit
    torch.testing.assert_close(z, torch.cat(zs), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(xn, torch.cat(xns), rtol=1e-6, atol=1e-7)
    z_raw = ops.gather_ste_loss(x, W, quant=q, normalize_x=False, want_norm=False)[0]
    z_raw2 = torch.cat([ops.gather_ste_loss(x[lo:hi].contiguous(), W, quant=q[lo:hi].contiguous(), normalize_x=False,
                                            want_norm=False)[0] for lo, hi in ((0, h), (h, N))])
    assert torch.equal(z_raw, z_raw2)                      # no reduction on the value path: bit-exact
    torch.testing.assert_close(mse4, m4[0] + m4[1], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(gx.float(), torch.cat(gxs).float(), rtol=1e-5 if dtype == torch.float32 else 1e-2, atol=1e-7)
    torch.testing.assert_close(gW, gW2, rtol=1e-4, atol=1e-6)


# ---- BASELINE.json configs[1..3] at their full size and dtype, one whole step vs the oracle --------------------
IDX_EPS = 1e-5     # near-tie: the oracle's own distance gap between the two candidates (cosine distances are O(1))

FULL = {
    # name: (quantizer config, oracle spec kwargs, N, K, D)
    'cfg2': (dict(type='VQKDQuantizer', distance=dict(type='CosineDistance'),
                  callbacks=[dict(type='VQKDCallback', ema=dict())],
                  losses=dict(commitment_loss=dict(type='CommitmentLoss', mse=dict(norm=True)))),
             dict(distance='Cosine', callback='VQKDCallback',
                  losses={'commitment_loss': dict(type='CommitmentLoss', norm=True)}), 65536, 8192, 32),
    'cfg3': (dict(type='VQKDQuantizer', distance=dict(type='CosineDistance'),
                  callbacks=[dict(type='VQKDCallback', ema=dict())],
                  losses=dict(commitment_loss=dict(type='CommitmentLoss', mse=dict(norm=True)))),
             dict(distance='Cosine', callback='VQKDCallback',
                  losses={'commitment_loss': dict(type='CommitmentLoss', norm=True)}), 65536, 16384, 8),
    'cfg4': (dict(type='VQGANQuantizer', distance=dict(type='CosineDistance'),
                  callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor'))],
                  losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan')),
             dict(distance='Cosine', callback='CVQVAECallback', losses={'vqgan_loss': dict(type='VQGANLoss')}),
             16384, 8192, 256),
}


@pytest.mark.parametrize('name,training', [('cfg2', False), ('cfg2', True), ('cfg3', True), ('cfg4', True)])
def test_full_size_step_vs_oracle(dev, name, training):
    """One whole step of BASELINE.json configs[1] (eval and training), [2] and [3] at their stated size, bf16 tokens,
    through the drop-in module, against `oracle.quantizer_forward` on the up-cast tokens.
    Tolerances: indices equal except near-ties (oracle distance gap of the two candidates < 1e-5); loss rel 1e-5;
    updated codebook rel 1e-5 / `_probability` 1e-6 on the codes no near-tie row touches; z rel 1e-5 and the bf16
    token gradient 2^-7 on the rows whose index and code row agree."""
    cfg, ospec, N, K, D = FULL[name]
    x32, E = O.synthetic_latents(N, K, D, seed=3407, normalized_codebook=True)
    xb = x32.to(torch.bfloat16)
    g = torch.Generator().manual_seed(7)
    gz = torch.randn(N, D, generator=g)
    is_cvq = ospec['callback'] == 'CVQVAECallback'
    prob0 = torch.rand(K, generator=g) / K if is_cvq else None    # a warm `_probability` (sums to ~0.5) ...
    if is_cvq:
        prob0[::2] *= 1e-3                                        # ... with every other code rarely used: its anchor counts

    q = vqb.build_quantizer(dict(cfg, embedding=emb(K, D)), training=training).to(dev)
    q._forward_pre_hooks.clear()
    with torch.no_grad():
        q.embedding.weight.copy_(E)
        if is_cvq:
            q.get_buffer('_probability').copy_(prob0)
    xg = xb.to(dev).requires_grad_(True)
    z, loss, memo = q(xg, dict())
    torch.autograd.backward((z, loss), (gz.to(dev), torch.ones([], device=dev)))
    torch.cuda.synchronize()

    spec = O.QuantizerSpec(training=training, **ospec)
    xo = xb.float().requires_grad_(True)
    out = O.quantizer_forward(spec, [xo], E, prob0)
    torch.autograd.backward((out['z_ste'][0], out['loss'][0]), (gz, torch.ones([])))
    d, q_ref = out['distance'][0], out['quant'][0]

    quant = memo['quant'].cpu()
    rows, gap = O.index_mismatch_report(d, q_ref, quant)
    assert (gap < IDX_EPS).all(), f'{name}: {int((gap >= IDX_EPS).sum())} index mismatches outside near-ties (max gap {float(gap.max())})'
    assert rows.numel() <= N // 200, f'{name}: {rows.numel()} near-tie rows'
    torch.testing.assert_close(loss.detach().cpu(), out['loss'][0].detach(), rtol=1e-5, atol=1e-7)

    # codes whose statistics a near-tie row (or, CVQ-VAE, a near-tie COLUMN arg-min) may have changed
    touched = torch.zeros(K, dtype=torch.bool)
    touched[q_ref[rows]] = True
    touched[quant[rows]] = True
    if training and is_cvq:
        # the column arg-min only runs for the codes whose anchor gets a non-zero fp32 blend weight; every code whose
        # weight IS non-zero in the oracle must have a key, and that key must be the oracle's nearest token
        ck = memo['encode']['column_keys'].cpu()
        has_key = ck != -1
        dec = 1 - torch.exp(-out['prob'] * K * 10 / (1 - spec.ema_decay) - spec.cvq_eps)
        needed = (1 - dec) != 0
        assert 0.05 < needed.float().mean() < 0.95 and has_key[needed].all() and has_key.float().mean() < 0.95
        col = ops.unpack_keys(memo['encode']['column_keys']).cpu()
        a_ref = out['anchor_idx'][0]
        bad = ((col != a_ref) & needed).nonzero().flatten()
        col_gap = d[col[bad], bad] - d[a_ref[bad], bad]
        assert (col_gap < IDX_EPS).all(), f'{name}: column arg-min mismatches outside near-ties'
        assert bad.numel() <= K // 100
        touched[bad] = True
        torch.testing.assert_close(q.get_buffer('_probability').cpu()[~touched], out['prob'][~touched], rtol=1e-6, atol=1e-9)
    W_gpu = q.embedding.weight.detach().cpu()
    torch.testing.assert_close(W_gpu[~touched], out['weight'][~touched], rtol=1e-5, atol=1e-6)
    assert int(touched.sum()) <= K // 20

    ok = (quant == q_ref) & ~touched[q_ref]
    assert ok.float().mean() > 0.9
    torch.testing.assert_close(z.detach().cpu()[ok], out['z_ste'][0].detach()[ok], rtol=1e-5, atol=1e-6)
    assert xg.grad.dtype == torch.bfloat16
    torch.testing.assert_close(xg.grad.float().cpu()[ok], xo.grad[ok], rtol=2 ** -7, atol=1e-5)
