"""The callers either side of the quantizer: NCHW in / NCHW out around the
drop-in quantizer (`BaseModel.quantize`, models/base.py:116-129), the tokenise-only path
(`encode_to_quant`, base.py:131-146) and the token dump formats."""
import pathlib

import numpy as np
import pytest
import torch

import vector_quantization_b200 as vqb
from oracle import oracle as O
from vector_quantization_b200 import ops, tokenizer
from vector_quantization_b200 import functional as Fq

pytestmark = pytest.mark.gpu
GOLDEN = pathlib.Path(__file__).parent / 'golden'


def _golden_quantizer(name, dev):
    rec = torch.load(GOLDEN / f'{name}.pt', weights_only=False)
    q = vqb.build_quantizer(rec['config'], training=rec['training']).to(dev)
    q._forward_pre_hooks.clear()
    step = rec['steps'][0]
    with torch.no_grad():
        q.embedding.weight.copy_(step['W_before'])
    return rec, q, step


@pytest.mark.parametrize('shape,dtype', [((3, 32, 16, 16), torch.float32), ((2, 8, 5, 7), torch.bfloat16),
                                         ((1, 256, 16, 16), torch.bfloat16), ((5, 33, 3, 11), torch.float32),
                                         ((2, 4, 1, 1), torch.int64), ((2, 8, 16, 16), torch.bfloat16),
                                         ((3, 6, 4, 4), torch.int64), ((2, 100, 16, 24), torch.float32)])
def test_transpose_kernel_matches_permute(dev, shape, dtype):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 100).to(dtype).to(dev)
    rows = ops.transpose_last2(x.view(b, c, h * w))
    assert torch.equal(rows.view(b * h * w, c), x.permute(0, 2, 3, 1).reshape(-1, c))      # 'b c h w -> (b h w) c'
    back = ops.transpose_last2(rows)
    assert torch.equal(back.view(shape), x)                                                # '(b h w) c -> b c h w'


@pytest.mark.parametrize('name', ['vqgan_l2', 'vqkd_eval'])
def test_quantize_nchw_equals_rows_path(dev, name):
    """tokenizer.quantize on [b, c, h, w] == the quantizer on the rearranged rows, rearranged back; the
    straight-through gradient arrives in NCHW; memo['quantizer'] carries x_shape and quant."""
    rec, q, step = _golden_quantizer(name, dev)
    D = rec['D']
    rows = step['x'].to(dev)                                                  # [(b h w), c], b = 2, h = w = 16
    b, h, w = 2, 16, 16
    x = rows.view(b, h, w, D).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    gz = torch.randn(b, D, h, w, generator=torch.Generator().manual_seed(1)).to(dev)
    z, loss, memo = tokenizer.quantize(q, x, dict())
    assert z.shape == (b, D, h, w) and z.is_contiguous() and memo['quantizer']['x_shape'] == (b, D, h, w)
    (z * gz).sum().add(loss).backward()
    rows_in = rows.clone().requires_grad_(True)
    with torch.no_grad():   # normalising callbacks rewrite the codebook in place every forward: same start state
        q.embedding.weight.copy_(step['W_before'])
    z_ref, loss_ref, memo_ref = q(rows_in, dict())
    (z_ref * gz.permute(0, 2, 3, 1).reshape(-1, D)).sum().add(loss_ref).backward()
    assert torch.equal(memo['quantizer']['quant'], memo_ref['quant'])
    assert torch.equal(z.detach(), z_ref.detach().view(b, h, w, D).permute(0, 3, 1, 2))
    assert torch.equal(loss.detach(), loss_ref.detach())
    assert torch.equal(x.grad, rows_in.grad.view(b, h, w, D).permute(0, 3, 1, 2))
    # and against the golden indices of the reference's own source
    assert (memo['quantizer']['quant'].cpu() != step['quant']).float().mean() < 0.01


def test_encode_to_quant_and_token_dumps(dev, tmp_path):
    rec, q, step = _golden_quantizer('vqgan_l2', dev)
    q.eval()
    D = rec['D']
    b, h, w = 2, 16, 16
    x = step['x'].to(dev).view(b, h, w, D).permute(0, 3, 1, 2).contiguous()
    quant, memo = tokenizer.encode_to_quant(q, x, dict())
    assert quant.shape == (b, h, w) and quant.dtype == torch.int64
    assert torch.equal(memo['quantizer']['quant'], quant.view(-1)) and memo['quantizer']['x'].shape == (b * h * w, D)
    compact, memo_c = tokenizer.encode_to_quant(q, x, dict(), compact=True)
    assert compact.dtype == torch.uint16 and torch.equal(compact.to(torch.int64), quant)
    # .pth dump of TokenizeCallback: {id_, category, tokens [b, h, w] int64}
    tokenizer.save_tokens(tmp_path / '1_0.pth', ['a', 'b'], torch.tensor([3, 7]), compact, memo_c['quantizer']['x_shape'])
    dump = tokenizer.load_tokens(tmp_path / '1_0.pth')
    assert set(dump) == {'id_', 'category', 'tokens'} and dump['tokens'].dtype == torch.int64
    assert torch.equal(dump['tokens'], quant.cpu()) and dump['id_'] == ['a', 'b']
    # LlamaGen .npy codes (1, 10, -1): ten crops of one image
    x10 = x[:1].repeat(10, 1, 1, 1)
    q10, _ = tokenizer.encode_to_quant(q, x10, dict(), compact=True)
    tokenizer.save_llamagen_codes(tmp_path / '0.npy', tmp_path / '0_label.npy', q10, torch.tensor([5]))
    codes = np.load(tmp_path / '0.npy')
    assert codes.shape == (1, 10, h * w) and codes.dtype == np.int64
    assert (codes[0, 3] == quant[0].view(-1).cpu().numpy()).all()


def test_compact_tokens_int32_for_large_codebooks(dev):
    keys = ops.new_keys(1000, dev)
    x = torch.randn(1000, 16, device=dev)
    E = torch.randn(70000, 16, device=dev)
    ops.assign(ops.pack_rows(x), ops.pack_rows(E, want_half_sqnorm=True), keys, l2=True)
    idx = ops.unpack_keys(keys)
    c = ops.compact_tokens(keys, 70000)
    assert c.dtype == torch.int32 and torch.equal(c.to(torch.int64), idx) and int(idx.max()) > 65535


def test_quantize_nchw_vs_oracle_caller(dev):
    """tokenizer.quantize / encode_to_quant against the oracle restatement of BaseModel.quantize /
    encode_to_quant (einops rearranges + quantizer forward) on seeded NCHW latents."""
    b, c, h, w, K = 4, 32, 16, 16, 512
    rows, E = O.synthetic_latents(b * h * w, K, c, seed=21)
    x = rows.view(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    spec = O.QuantizerSpec(distance='L2', losses={'vqgan_loss': dict(type='VQGANLoss')})
    z_ref, loss_ref, out = O.model_quantize(spec, x, E)
    q = vqb.build_quantizer(dict(type='VQGANQuantizer',
                                 embedding=dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=c),
                                 distance=dict(type='L2Distance'), losses=dict(vqgan_loss=dict(type='VQGANLoss')),
                                 init_weights=dict(type='vqgan'))).to(dev)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    z, loss, memo = tokenizer.quantize(q, x.to(dev), dict())
    quant = memo['quantizer']['quant'].cpu()
    same = (quant == out['quant'][0])
    assert same.float().mean() > 0.999
    keep = same.view(b, 1, h, w).expand(b, c, h, w)
    assert torch.equal(z.detach().cpu()[keep], z_ref.detach()[keep])
    torch.testing.assert_close(loss.detach().cpu(), loss_ref.detach(), rtol=1e-5, atol=1e-7)
    tokens, _ = tokenizer.encode_to_quant(q, x.to(dev), dict())
    assert (tokens.cpu() == O.model_encode_to_quant('L2', x, E)).float().mean() > 0.999


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16], ids=['fp32', 'bf16'])
@pytest.mark.parametrize('c,hw,cfg_kind', [(32, (16, 16), 'vqkd'), (256, (8, 8), 'vqgan'), (8, (16, 16), 'llamagen'),
                                           (32, (8, 8), 'llamagen_cvq')])
def test_layout_fusion_launches_one_transpose_and_matches_the_rows_path(dev, dtype, c, hw, cfg_kind):
    """`tokenizer.quantize` folds the caller's rearranges into the kernels — ONE transposed copy of the
    latents on the way in, z written NCHW by the gather kernel, the backward reading / writing NCHW gradients: a
    single transpose launch per forward+backward instead of four.  Configurations whose callbacks need normalised
    tokens up front (NormalizeCallback + CVQVAECallback) keep the four-transpose path.  Either way the results equal
    the token-major path bit for bit."""
    h, w = hw
    b, K = 4, 256
    emb = dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=c)
    cfgs = {
        'vqkd': dict(type='VQKDQuantizer', distance=dict(type='CosineDistance'), callbacks=[dict(type='VQKDCallback', ema=dict())],
                     losses=dict(commitment_loss=dict(type='CommitmentLoss', mse=dict(norm=True)))),
        'vqgan': dict(type='VQGANQuantizer', distance=dict(type='L2Distance'), losses=dict(vqgan_loss=dict(type='VQGANLoss')),
                      init_weights=dict(type='vqgan')),
        'llamagen': dict(type='VQGANQuantizer', distance=dict(type='L2Distance'), callbacks=[dict(type='NormalizeCallback')],
                         losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan')),
        'llamagen_cvq': dict(type='VQGANQuantizer', distance=dict(type='L2Distance'),
                             callbacks=[dict(type='NormalizeCallback'),
                                        dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor'))],
                             losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan')),
    }
    q = vqb.build_quantizer(dict(cfgs[cfg_kind], embedding=emb), training=False).to(dev)
    q._forward_pre_hooks.clear()
    rows, E = O.synthetic_latents(b * h * w, K, c, seed=c + h, normalized_codebook=True)
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    x0 = rows.view(b, h, w, c).permute(0, 3, 1, 2).contiguous().to(dtype).to(dev)
    gz = torch.randn(b, c, h, w, generator=torch.Generator().manual_seed(2)).to(dev)
    x = x0.clone().requires_grad_(True)
    ops.PROFILE = []
    z, loss, memo = tokenizer.quantize(q, x, dict())
    (z * gz).sum().add(loss).backward()
    torch.cuda.synchronize()
    transposes = sum(name == 'vqb_transpose_last2' for name, _, _ in ops.PROFILE)
    ops.PROFILE = None
    assert q.can_fuse_nchw() == (cfg_kind != 'llamagen_cvq')
    assert transposes == (1 if q.can_fuse_nchw() else 4)      # unfused: in + out, and both again in the backward
    assert z.shape == (b, c, h, w) and z.is_contiguous() and x.grad.shape == x.shape and x.grad.dtype == dtype
    rows_in = x0.permute(0, 2, 3, 1).reshape(-1, c).contiguous().clone().requires_grad_(True)
    with torch.no_grad():   # normalising callbacks rewrite the codebook in place every forward: same start state
        q.embedding.weight.copy_(E)
    z_ref, loss_ref, memo_ref = q(rows_in, dict())
    (z_ref * gz.permute(0, 2, 3, 1).reshape(-1, c)).sum().add(loss_ref).backward()
    assert torch.equal(memo['quantizer']['quant'], memo_ref['quant'])
    assert torch.equal(z.detach(), z_ref.detach().view(b, h, w, c).permute(0, 3, 1, 2))
    assert torch.equal(loss.detach(), loss_ref.detach())
    assert torch.equal(x.grad, rows_in.grad.view(b, h, w, c).permute(0, 3, 1, 2))
