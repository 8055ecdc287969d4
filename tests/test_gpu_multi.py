"""2-GPU (NCCL) parity tests of the exchange steps against the multi-rank oracle: token-sharded VQ-KD
EMA update, CVQ-VAE anchors with sync=False (all-reduce mean) and sync=True (packed min-loc all-reduce
instead of the reference's all_gather of the N x K matrix), and the codebook-sharded assignment.
Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')]


def emb(K, D):
    return dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        out.put((rank, CASES[case](rank, world)))
    finally:
        dist.destroy_process_group()


def _spawn(case, world=2):
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(out.get() for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return [res[r] for r in range(world)]


def _quantizer_step(cfg, N, K, D, normalized, rank, world, steps=2):
    import vector_quantization_b200 as vqb
    from oracle import oracle as O
    dev = torch.device('cuda', rank)
    x_all, E = O.synthetic_latents(N * world * steps, K, D, seed=5, normalized_codebook=normalized)
    q = vqb.build_quantizer(dict(cfg, embedding=emb(K, D)), training=True).to(dev)
    q._forward_pre_hooks.clear()
    with torch.no_grad():
        q.embedding.weight.copy_(E)
    outs = []
    for s in range(steps):
        x = x_all[(s * world + rank) * N:(s * world + rank + 1) * N].to(dev).requires_grad_(True)
        z, loss, memo = q(x, dict())
        loss.backward()
        outs.append(dict(quant=memo['quant'].cpu(), loss=loss.detach().cpu(), W=q.embedding.weight.detach().cpu(),
                         prob=q.get_buffer('_probability').cpu() if hasattr(q, '_probability') else None))
    return outs


VQKD = dict(type='VQKDQuantizer', distance=dict(type='CosineDistance'), callbacks=[dict(type='VQKDCallback', ema=dict())],
            losses=dict(commitment_loss=dict(type='CommitmentLoss', mse=dict(norm=True))))
CVQ = dict(type='VQGANQuantizer', distance=dict(type='CosineDistance'),
           callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor'))],
           losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
CLUSTER = dict(type='VQGANQuantizer', distance=dict(type='CosineDistance'),
               callbacks=[dict(type='CVQVAECallback', ema=dict(), anchor=dict(type='NearestAnchor', sync=True))],
               losses=dict(vqgan_loss=dict(type='CodebookLoss')), init_weights=dict(type='vqgan'))


def _sharded(rank, world):
    from oracle import oracle as O
    from vector_quantization_b200 import parallel
    dev = torch.device('cuda', rank)
    N, K, D = 3000, 2048, 64
    x, E = O.synthetic_latents(N, K, D, seed=9)
    lo, hi = parallel.shard_range(K, rank, world)
    quant, keys = parallel.sharded_nearest_code(x.to(dev), E[lo:hi].to(dev), 'L2', shard_lo=lo)
    z = parallel.sharded_decode(keys, E[lo:hi].to(dev), lo)
    return quant.cpu(), z.cpu()


CASES = {
    'vqkd': lambda r, w: _quantizer_step(VQKD, 512, 128, 32, True, r, w),
    'cvq': lambda r, w: _quantizer_step(CVQ, 384, 96, 32, False, r, w),
    'cluster': lambda r, w: _quantizer_step(CLUSTER, 256, 64, 64, False, r, w),
    'sharded': _sharded,
}

SPECS = {
    'vqkd': (dict(distance='Cosine', callback='VQKDCallback',
                  losses={'commitment_loss': dict(type='CommitmentLoss', norm=True)}), 512, 128, 32, True),
    'cvq': (dict(distance='Cosine', callback='CVQVAECallback', losses={'vqgan_loss': dict(type='VQGANLoss')}),
            384, 96, 32, False),
    'cluster': (dict(distance='Cosine', callback='CVQVAECallback', anchor_sync=True,
                     losses={'vqgan_loss': dict(type='CodebookLoss')}), 256, 64, 64, False),
}


@pytest.mark.parametrize('case', ['vqkd', 'cvq', 'cluster'])
def test_token_sharded_training_steps_match_multi_rank_oracle(case):
    from oracle import oracle as O
    world, steps = 2, 2
    res = _spawn(case, world)
    spec_kw, N, K, D, normalized = SPECS[case]
    spec = O.QuantizerSpec(**spec_kw)
    x_all, W = O.synthetic_latents(N * world * steps, K, D, seed=5, normalized_codebook=normalized)
    prob = torch.zeros(K) if spec.callback == 'CVQVAECallback' else None
    for s in range(steps):
        xs = [x_all[(s * world + r) * N:(s * world + r + 1) * N] for r in range(world)]
        out = O.quantizer_forward(spec, xs, W, prob)
        mismatch = sum(int((res[r][s]['quant'] != out['quant'][r]).sum()) for r in range(world))
        assert mismatch <= 2, f'{case} step {s}: {mismatch} index mismatches'
        for r in range(world):
            torch.testing.assert_close(res[r][s]['loss'], out['loss'][r].detach(), rtol=2e-5, atol=1e-7)
        assert torch.equal(res[0][s]['W'], res[1][s]['W']), 'replicas diverged'
        if mismatch == 0:
            torch.testing.assert_close(res[0][s]['W'], out['weight'], rtol=1e-5, atol=1e-6)
            if prob is not None:
                torch.testing.assert_close(res[0][s]['prob'], out['prob'], rtol=1e-6, atol=1e-9)
        W, prob = out['weight'], out['prob']


def test_codebook_sharded_assignment_and_decode():
    from oracle import oracle as O
    res = _spawn('sharded')
    x, E = O.synthetic_latents(3000, 2048, 64, seed=9)
    q_ref, d = O.encode('L2', x, E)
    for quant, z in res:
        rows, gap = O.index_mismatch_report(d, q_ref, quant)
        assert (gap < 1e-5 * d[rows, q_ref[rows]].clamp_min(1)).all() and rows.numel() <= 3
        assert torch.equal(z, E[quant])
    assert torch.equal(res[0][0], res[1][0])
