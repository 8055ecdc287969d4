"""Host-side logic that needs no GPU: registries/configs, state-dict key compatibility, loss term tables,
callback dispatch order, key packing, failure modes."""
import pytest
import torch

import vector_quantization_b200 as vqb
from vector_quantization_b200 import parallel
from vector_quantization_b200.callbacks import BaseCallback, ComposedCallback


def emb(K, D):
    return dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D)


VQGAN = dict(type='VQGANQuantizer', embedding=emb(64, 8), distance=dict(type='L2Distance'),
             losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
VQKD = dict(type='VQKDQuantizer', embedding=emb(64, 8), distance=dict(type='CosineDistance'),
            callbacks=[dict(type='VQKDCallback', ema=dict())],
            losses=dict(commitment_loss=dict(type='CommitmentLoss', mse=dict(norm=True))))
CLUSTER = dict(type='VQGANQuantizer', embedding=emb(64, 8), distance=dict(type='CosineDistance'),
               callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor', sync=True))],
               losses=dict(vqgan_loss=dict(type='CodebookLoss')), init_weights=dict(type='vqgan'))


def test_registry_names_of_the_reference_are_all_present():
    R = vqb.registry
    for n in ('VectorQuantizer', 'VQGANQuantizer', 'VQKDQuantizer', 'FiniteScalarQuantizer', 'ScalarQuantizer'):
        assert n in R.VQITQuantizerRegistry
    for n in ('ComposedCallback', 'NormalizeCallback', 'CVQVAECallback', 'VQKDCallback'):
        assert n in R.VQITQuantizerCallbackRegistry
    for n in ('CodebookLoss', 'CommitmentLoss', 'VQGANLoss', 'EntropyLoss'):
        assert n in R.VQITQuantizerLossRegistry
    for n in ('L2Distance', 'CosineDistance'):
        assert n in R.VQITQuantizerDistanceRegistry
    for n in ('NearestAnchor', 'MultinomialAnchor', 'CachedAnchor'):
        assert n in R.AnchorRegistry


def test_state_dict_keys_match_reference_checkpoints():
    # tools/convert_checkpoints.py:239-243 (VQGAN) and :321-322 (VQ-KD) in the reference
    q = vqb.build_quantizer(VQGAN)
    assert set(q.state_dict()) == {
        '_embedding.weight', '_losses.vqgan_loss._weight._steps', '_losses.vqgan_loss._codebook._weight._steps',
        '_losses.vqgan_loss._codebook._mse._weight._steps', '_losses.vqgan_loss._commitment._weight._steps',
        '_losses.vqgan_loss._commitment._mse._weight._steps'}
    q = vqb.build_quantizer(VQKD)
    assert set(q.state_dict()) == {'_embedding.weight', '_losses.commitment_loss._weight._steps',
                                   '_losses.commitment_loss._mse._weight._steps'}
    q = vqb.build_quantizer(CLUSTER)
    assert '_probability' in q.state_dict() and q.state_dict()['_probability'].shape == (64,)
    q = vqb.build_quantizer(dict(type='FiniteScalarQuantizer', num_scalars_per_channel=[8, 8, 8, 5, 5, 5]))
    assert list(q.state_dict()) == ['_embeddings'] and q.codebook_size == 64000 and q.embedding_dim == 6


def test_vqgan_init_is_uniform_pm_one_over_k():
    q = vqb.build_quantizer(VQGAN)
    w = q.embedding.weight
    assert float(w.abs().max()) <= 1 / 64 and float(w.std()) > 0


def test_loss_term_tables():
    q = vqb.build_quantizer(VQGAN)
    assert q._losses['vqgan_loss'].terms() == {0: 1.0, 1: 0.25}       # codebook + 0.25 * commitment
    q = vqb.build_quantizer(VQKD)
    assert q._losses['commitment_loss'].terms() == {3: 1.0} and q._loss_terms() is True
    q = vqb.build_quantizer(CLUSTER)
    assert q._losses['vqgan_loss'].terms() == {0: 1.0}
    mse4 = torch.tensor([2.0, 2.0, 8.0, 8.0])
    q = vqb.build_quantizer(dict(VQGAN, losses=dict(l=dict(type='VQGANLoss', beta=0.5))))
    assert float(q._losses['l'].from_mse4(mse4)) == 3.0


def test_callback_flags_and_lazy_init_hook():
    q = vqb.build_quantizer(VQKD)
    assert len(q._forward_pre_hooks) == 1 and not q._callbacks.needs_column_nearest
    q = vqb.build_quantizer(CLUSTER)
    assert q._callbacks.needs_column_nearest and q._callbacks.column_nearest_global
    q = vqb.build_quantizer(dict(CLUSTER, callbacks=[dict(type='CVQVAECallback', ema=dict(decay=0.9),
                                                          anchor=dict(type='NearestAnchor'))]))
    assert not q._callbacks.column_nearest_global
    assert next(iter(q._callbacks))._ema.decay == 0.9


def test_composed_callback_priority_order():
    calls = []

    class A(BaseCallback):
        def before_encode(self, x, memo):
            calls.append('A')
            return x

    class B(BaseCallback):
        def before_encode(self, x, memo):
            calls.append('B')
            return x

    cc = ComposedCallback(priorities=[dict(), dict(before_encode=5)], callbacks=[A(), B()])
    cc.before_encode(torch.zeros(1), {})
    assert calls == ['B', 'A'] and cc.overrides('before_encode') and not cc.overrides('after_loss')


def test_unsupported_components_fail_loudly():
    with pytest.raises(vqb._lib.VQBError):          # the distance modules materialise on the GPU only (compat mode)
        vqb.L2Distance()(torch.zeros(2, 2), torch.zeros(2, 2))
    q = vqb.build_quantizer(VQGAN)
    with pytest.raises(vqb._lib.VQBError):          # CPU tensors: no fallback
        q(torch.zeros(4, 8), {})
    with pytest.raises(vqb._lib.VQBError):
        vqb.ops.pack_rows(torch.zeros(4, 8))


def test_key_packing_orders_like_the_device_keys():
    s = torch.tensor([-3.5, -0.0, 0.0, 1e-30, 0.25, 7.0, float('inf'), -float('inf')])
    idx = torch.arange(8)
    k = parallel.pack_keys_host(s, idx)
    # smaller key == better: compare as unsigned
    ku = [(int(v) + (1 << 64)) % (1 << 64) for v in k]
    order = sorted(range(8), key=lambda i: ku[i])
    assert [float(s[i]) for i in order][:2] == [float('inf'), 7.0] and order[-1] == 7
    # equal scores: lower index wins
    k2 = parallel.pack_keys_host(torch.tensor([1.0, 1.0]), torch.tensor([9, 3]))
    ku2 = [(int(v) + (1 << 64)) % (1 << 64) for v in k2]
    assert ku2[1] < ku2[0]
    sc, ix = parallel.unpack_keys_host(k)
    assert torch.equal(ix, idx) and torch.equal(sc[[0, 4, 5]], s[[0, 4, 5]])


def test_shard_range_covers_everything_once():
    for total, w in ((262144, 8), (1000, 3), (5, 8)):
        spans = [parallel.shard_range(total, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_token_dump_formats(tmp_path):
    """`.pth` Tokens dict of TokenizeCallback (runners/callbacks.py:40-53) and the LlamaGen `.npy` codes of shape
    (1, 10, -1) (tools/tokenize_llamagen.py:93-103); compact uint16 / int32 ids are widened to int64 on disk."""
    import numpy as np
    from vector_quantization_b200 import tokenizer
    quant = torch.randint(0, 60000, (2 * 4 * 4,)).to(torch.uint16)
    tokenizer.save_tokens(tmp_path / 't.pth', ['x', 'y'], torch.tensor([1, 2]), quant, (2, 8, 4, 4))
    d = tokenizer.load_tokens(tmp_path / 't.pth')
    assert d['tokens'].shape == (2, 4, 4) and d['tokens'].dtype == torch.int64
    assert torch.equal(d['tokens'].view(-1), quant.to(torch.int32).to(torch.int64)) and int(d['tokens'].max()) > 32767
    q10 = torch.randint(0, 16384, (10, 16, 16), dtype=torch.int32)
    tokenizer.save_llamagen_codes(tmp_path / 'c.npy', tmp_path / 'l.npy', q10, torch.tensor([7]))
    codes, labels = np.load(tmp_path / 'c.npy'), np.load(tmp_path / 'l.npy')
    assert codes.shape == (1, 10, 256) and codes.dtype == np.int64 and (codes[0] == q10.view(10, -1).numpy()).all()
    assert labels.tolist() == [7]


def test_oracle_caller_restatement_matches_reference_rearranges():
    """oracle.model_quantize follows BaseModel.quantize: token n = (b, h, w) row-major, channels last."""
    from oracle import oracle as O
    b, c, h, w, K = 2, 8, 3, 5, 16
    rows, E = O.synthetic_latents(b * h * w, K, c, seed=2)
    x = rows.view(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    spec = O.QuantizerSpec(distance='L2', losses={'l': dict(type='VQGANLoss')})
    z, loss, out = O.model_quantize(spec, x, E)
    ref = O.quantizer_forward(spec, [rows], E)
    assert torch.equal(out['quant'][0], ref['quant'][0]) and torch.equal(loss, ref['loss'][0])
    assert torch.equal(z, ref['z_ste'][0].view(b, h, w, c).permute(0, 3, 1, 2)) and z.is_contiguous()
    assert torch.equal(O.model_encode_to_quant('L2', x, E), ref['quant'][0].view(b, h, w))


@pytest.mark.parametrize('anchor_type', ['NearestAnchor', 'CachedAnchor'])
def test_cvqvae_callback_wiring_with_both_anchor_kinds(monkeypatch, anchor_type):
    """Host-side wiring of CVQVAECallback.after_encode (no GPU): the C-ABI wrappers are replaced by torch
    restatements so that the call sequence, the optional column keys and the anchor `gather` contract are exercised
    for NearestAnchor (needs the column arg-min pass) and CachedAnchor (samples rows; no column pass)."""
    import random
    from oracle import oracle as O
    from vector_quantization_b200 import ops
    from vector_quantization_b200.anchors import cached_rows_and_indices
    K, D, N = 16, 4, 40
    q = vqb.build_quantizer(dict(
        type='VQGANQuantizer', embedding=dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D),
        distance=dict(type='CosineDistance'), losses=dict(l=dict(type='CodebookLoss')),
        callbacks=[dict(type='CVQVAECallback', ema=dict(decay=0.9), anchor=dict(type=anchor_type))],
        init_weights=dict(type='vqgan')), training=True)
    cb = [c for c in q._callbacks._callbacks if type(c).__name__ == 'CVQVAECallback'][0]
    assert q._callbacks.needs_column_nearest == (anchor_type == 'NearestAnchor')

    def bincount(quant, counts, K_=None, total_slot=False):
        counts[:K] += torch.bincount(quant, minlength=K)
        if total_slot:
            counts[K] += quant.numel()
        return counts
    seen = {}
    monkeypatch.setattr(ops, 'bincount_accumulate', bincount)
    monkeypatch.setattr(ops, 'gather_rows_by_key', lambda rows, keys, off=0: rows.float()[(keys & 0xffffffff) - off])
    monkeypatch.setattr(ops, 'cvq_update', lambda W, anchors, prob, cnt, tot, **kw: seen.update(anchors=anchors.clone(), kw=kw))
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, D, generator=g)
    quant = torch.randint(0, K, (N,), generator=g)
    d = O.distance('Cosine', x, q.embedding.weight.data)
    memo = dict(encode=dict())
    if anchor_type == 'NearestAnchor':
        memo['encode']['column_keys'] = d.argmin(0)          # token index in the low key bits
    torch.manual_seed(5); random.seed(5)
    out = cb.after_encode(x, quant, memo)
    assert out is quant and seen['kw']['anchor_scale'] == 1.0 and seen['anchors'].shape == (K, D)
    if anchor_type == 'NearestAnchor':
        assert torch.equal(seen['anchors'], x[d.argmin(0)])
    else:
        torch.manual_seed(5); random.seed(5)
        rows, idx = cached_rows_and_indices(x, K, torch.empty(0))
        assert torch.equal(seen['anchors'], rows[idx]) and torch.equal(cb._anchor.cache, rows[idx])
        assert not any('_cache' in k for k in q.state_dict())   # the cache lives in a callback: not checkpointed


def test_operand_format_selection(monkeypatch):
    """Which plane format the host layer asks the pack kernel for (no GPU: ops.pack_rows is recorded):
    fp16 pair only for a normalised fp32 codebook matched against bf16 tokens at the default precision;
    exact bf16 planes for L2 / fp32 tokens / precision='exact'; fewer planes for 'high' / 'fast'."""
    from vector_quantization_b200 import functional as Fq, ops
    calls = []

    def fake_pack(src, **kw):
        calls.append(kw)
        fmt = kw.get('fmt', 'bf16')
        planes = {'f16x2': 2, 'f16': 1}.get(fmt, kw.get('planes') or (1 if src.dtype == torch.bfloat16 and not kw.get('normalize') else 3))
        return ops.Operand(torch.empty(0), src.shape[0], src.shape[1], planes, None, fmt=fmt)
    monkeypatch.setattr(ops, 'pack_rows', fake_pack)
    W = torch.randn(64, 32)
    xb, xf = torch.randn(10, 32).to(torch.bfloat16), torch.randn(10, 32)
    assert Fq.pack_codebook(W, 'Cosine', tokens=xb).fmt == 'f16x2'
    assert Fq.pack_codebook(W, 'Cosine', tokens=xf).fmt == 'bf16' and calls[-1]['planes'] == 3
    assert Fq.pack_codebook(W, 'Cosine').fmt == 'bf16'                                   # tokens unknown: exact planes
    assert Fq.pack_codebook(W, 'Cosine', tokens=xb, precision='exact').fmt == 'bf16' and calls[-1]['planes'] == 3
    assert Fq.pack_codebook(W, 'Cosine', tokens=xb, precision='high').fmt == 'bf16' and calls[-1]['planes'] == 2
    assert Fq.pack_codebook(W, 'Cosine', tokens=xb, precision='fast').fmt == 'bf16' and calls[-1]['planes'] == 1
    book = Fq.pack_codebook(W, 'L2', tokens=xb)                                          # un-normalised rows: no fp16
    assert book.fmt == 'bf16' and calls[-1]['planes'] == 3 and calls[-1]['want_half_sqnorm']
    assert Fq.pack_codebook(W, 'L2', tokens=xb, writeback_normalized=True).fmt == 'bf16'  # LlamaGen L2 stays on bf16 planes


def test_compat_components_build_from_reference_style_configs():
    """EntropyLoss / MultinomialAnchor / VQGAN_VQKDCallback / materialize_distance: registry names, flags that switch
    the quantizer into the distance-materialising compatibility mode, and the packed-key / raw-token gating."""
    emb = dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=64, embedding_dim=8)
    q = vqb.build_quantizer(dict(
        type='VQGANQuantizer', embedding=emb, distance=dict(type='CosineDistance'),
        callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='MultinomialAnchor'))],
        losses=dict(vqgan_loss=dict(type='VQGANLoss'), ent=dict(type='EntropyLoss', temperature=0.5)),
        init_weights=dict(type='vqgan')))
    assert q.wants_distance and q._callbacks.needs_distance and not q._callbacks.needs_column_nearest
    assert not q._loss_terms() and not q._losses['ent'].uses_mse4
    plain = vqb.build_quantizer(VQGAN)
    assert not plain.wants_distance
    plain.materialize_distance = True
    assert plain.wants_distance
    q2 = vqb.build_quantizer(dict(type='VQGANQuantizer', embedding=emb, distance=dict(type='L2Distance'),
                                  callbacks=[dict(type='VQGAN_VQKDCallback', ema=dict(decay=0.9))],
                                  losses=dict(l=dict(type='VQGANLoss')), init_weights=dict(type='vqgan')))
    cb = list(q2._callbacks)[0]
    assert type(cb).__name__ == 'VQGAN_VQKDCallback' and cb._ema.decay == 0.9
    assert q2._callbacks.lazy_normalize_ok() and q2._callbacks.packed_keys_ok() and len(q2._forward_pre_hooks) == 1
    # NormalizeCallback + CVQVAECallback (llamagen + cvqvae mixin): tokens must be normalised up front
    q3 = vqb.build_quantizer(dict(type='VQGANQuantizer', embedding=emb, distance=dict(type='L2Distance'),
                                  callbacks=[dict(type='NormalizeCallback'),
                                             dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor'))],
                                  losses=dict(l=dict(type='VQGANLoss')), init_weights=dict(type='vqgan')))
    assert not q3._callbacks.lazy_normalize_ok() and not q3._callbacks.packed_keys_ok()
    q3.eval()
    assert q3._callbacks.packed_keys_ok()


def test_token_stream_writer_on_host_tensors(tmp_path):
    from vector_quantization_b200 import tokenizer
    with tokenizer.TokenStreamWriter(tmp_path, rank=3) as w:
        for it in range(5):
            w.write(it, [f'a{it}', f'b{it}'], torch.tensor([1, 2]), torch.arange(32, dtype=torch.int32) + it, (2, 8, 4, 4))
    rec = tokenizer.load_tokens(tmp_path / '2_3.pth')
    assert rec['tokens'].shape == (2, 4, 4) and rec['tokens'].dtype == torch.int64 and int(rec['tokens'][0, 0, 0]) == 2
    assert rec['id_'] == ['a2', 'b2'] and len(list(tmp_path.glob('*.pth'))) == 5
