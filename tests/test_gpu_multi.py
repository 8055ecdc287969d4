"""2-GPU parity tests of the exchange steps against the multi-rank oracle: token-sharded VQ-KD EMA update,
CVQ-VAE anchors with sync=False (mean over ranks) and sync=True (packed min-loc instead of the reference's
all_gather of the N x K matrix), the codebook-sharded assignment and the distributed k-means init, each with
the fused NVLink peer-memory exchange kernels (`VQB_COMM=p2p`, the default) AND with the torch.distributed
(NCCL) collectives (`VQB_COMM=nccl`).
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


def _worker(rank, world, port, case, out, comm, ack):
    # VQB_DRY_RUN: the reference's DRY_RUN `is_sync` assertions — every codebook update checks that all replicas agree
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), VQB_COMM=comm, VQB_DRY_RUN='1')
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        out.put((rank, CASES[case](rank, world)))
        ack.wait(120)            # stay alive until the parent has read the (shared-memory) tensors
    finally:
        dist.destroy_process_group()


def _spawn(case, world=2, comm='p2p'):
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = _free_port()
    ack = ctx.Event()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, out, comm, ack)) for r in range(world)]
    for p in procs:
        p.start()
    # never block on the queue: a worker that died (exception, trapped kernel) must fail the test, not hang it
    import time
    res, deadline = {}, time.time() + 240
    while len(res) < world:
        if not out.empty():
            r, v = out.get()
            res[r] = v
            continue
        dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
        if dead or time.time() > deadline or all(p.exitcode is not None for p in procs):
            for p in procs:
                if p.is_alive():
                    p.kill()
            raise AssertionError(f'worker(s) failed or timed out: exit codes {[p.exitcode for p in procs]}')
        time.sleep(0.05)
    ack.set()
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
    fused = [bool(c._region) for c in q._callbacks if hasattr(c, '_region')]
    outs[0]['fused'] = bool(fused and fused[0])
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
    quant2, _ = parallel.sharded_nearest_code(x.to(dev), E[lo:hi].to(dev), 'L2', shard_lo=lo)   # region reuse
    assert torch.equal(quant, quant2)
    z = parallel.sharded_decode(keys, E[lo:hi].to(dev), lo)
    return quant.cpu(), z.cpu()


def _sharded_certified(rank, world):
    """cfg-5-like: cosine, bf16 tokens, D = 256 -> the globally certified one-term pass."""
    from oracle import oracle as O
    from vector_quantization_b200 import functional as Fq
    from vector_quantization_b200 import parallel
    dev = torch.device('cuda', rank)
    N, K, D = 3000, 2048, 256
    x, E = O.synthetic_latents(N, K, D, seed=13, normalized_codebook=True, clustered=False)
    lo, hi = parallel.shard_range(K, rank, world)
    xb = x.to(torch.bfloat16).to(dev)
    quant, keys = parallel.sharded_nearest_code(xb, E[lo:hi].to(dev), 'Cosine', shard_lo=lo)
    flagged = int(Fq.LAST_CERTIFY['count'])
    quant2, _ = parallel.sharded_nearest_code(xb, E[lo:hi].to(dev), 'Cosine', shard_lo=lo, precision='exact')
    return quant.cpu(), quant2.cpu(), flagged


def _lazy_init(rank, world):
    """Distributed k-means init: every rank keeps ITS tokens; the result must equal the reference's rank-0 k-means
    over the concatenated tokens (oracle); rank 0 seeds `random` like the oracle run."""
    import random

    import vector_quantization_b200 as vqb
    from oracle import oracle as O
    dev = torch.device('cuda', rank)
    N, K, D = 1024, 64, 16
    x_all, _ = O.synthetic_latents(N * world, K, D, seed=11, normalized_codebook=True)
    cfg = dict(VQKD, embedding=emb(K, D), init_weights=dict(before_init_weights=dict(lazy_init_weights=dict(iters=10))))
    torch.manual_seed(0)
    q = vqb.build_quantizer(cfg, training=True).to(dev)
    seen = {}

    def record(module, args):
        seen.setdefault('W_init', module.embedding.weight.detach().clone())

    q.register_forward_pre_hook(record)
    random.seed(123)
    q(x_all[rank * N:(rank + 1) * N].to(dev), dict())
    return seen['W_init'].cpu()


CASES = {
    'vqkd': lambda r, w: _quantizer_step(VQKD, 512, 128, 32, True, r, w),
    'cvq': lambda r, w: _quantizer_step(CVQ, 384, 96, 32, False, r, w),
    'cluster': lambda r, w: _quantizer_step(CLUSTER, 256, 64, 64, False, r, w),
    'sharded': _sharded,
    'sharded_certified': _sharded_certified,
    'lazy_init': _lazy_init,
}

SPECS = {
    'vqkd': (dict(distance='Cosine', callback='VQKDCallback',
                  losses={'commitment_loss': dict(type='CommitmentLoss', norm=True)}), 512, 128, 32, True),
    'cvq': (dict(distance='Cosine', callback='CVQVAECallback', losses={'vqgan_loss': dict(type='VQGANLoss')}),
            384, 96, 32, False),
    'cluster': (dict(distance='Cosine', callback='CVQVAECallback', anchor_sync=True,
                     losses={'vqgan_loss': dict(type='CodebookLoss')}), 256, 64, 64, False),
}


@pytest.mark.parametrize('comm', ['p2p', 'nccl'])
@pytest.mark.parametrize('case', ['vqkd', 'cvq', 'cluster'])
def test_token_sharded_training_steps_match_multi_rank_oracle(case, comm):
    from oracle import oracle as O
    world, steps = 2, 2
    res = _spawn(case, world, comm)
    assert res[0][0]['fused'] == (comm == 'p2p'), 'the fused peer-memory exchange must be the path that ran'
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


def test_distributed_kmeans_init_matches_rank0_oracle():
    import random

    from oracle import oracle as O
    N, K, D, world = 1024, 64, 16, 2
    res = _spawn('lazy_init', world)
    assert torch.equal(res[0], res[1]), 'replicas diverged'
    x_all, _ = O.synthetic_latents(N * world, K, D, seed=11, normalized_codebook=True)
    random.seed(123)
    want = O.vqkd_lazy_init(x_all, torch.zeros(K, D), 10)
    close = torch.isclose(res[0], want, rtol=1e-4, atol=1e-5).all(1)
    assert close.float().mean() >= 0.95, f'{int((~close).sum())} of {K} centroids differ'


@pytest.mark.parametrize('comm', ['p2p', 'nccl'])
def test_codebook_sharded_assignment_and_decode(comm):
    from oracle import oracle as O
    res = _spawn('sharded', comm=comm)
    x, E = O.synthetic_latents(3000, 2048, 64, seed=9)
    q_ref, d = O.encode('L2', x, E)
    for quant, z in res:
        rows, gap = O.index_mismatch_report(d, q_ref, quant)
        assert (gap < 1e-5 * d[rows, q_ref[rows]].clamp_min(1)).all() and rows.numel() <= 3
        assert torch.equal(z, E[quant])
    assert torch.equal(res[0][0], res[1][0])


@pytest.mark.parametrize('comm', ['p2p', 'nccl'])
def test_codebook_sharded_assignment_globally_certified_one_term(comm):
    """Cosine, D = 256, bf16 tokens, codebook split over two ranks: the one-term pass certified against the GLOBAL
    runner-up equals the exact three-plane contraction index for index (and the oracle under the near-tie policy);
    on random data some rows fail the certificate and are re-run on every shard."""
    from oracle import oracle as O
    res = _spawn('sharded_certified', comm=comm)
    x, E = O.synthetic_latents(3000, 2048, 256, seed=13, normalized_codebook=True, clustered=False)
    q_ref, d = O.encode('Cosine', x.to(torch.bfloat16).float(), E)
    for quant, quant_exact, flagged in res:
        assert 0 < flagged < 1500
        rows, gap = O.index_mismatch_report(d, q_ref, quant)
        assert (gap < 1e-5).all() and rows.numel() <= 6
        assert (quant != quant_exact).sum() <= 2           # the pair's own 2^-22 operand error at fp32-level near-ties
    assert torch.equal(res[0][0], res[1][0])
