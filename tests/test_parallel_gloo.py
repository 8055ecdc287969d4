"""world_size-2 gloo tests (CPU) of the multi-GPU exchange steps in `vector_quantization_b200.parallel`.

The kernels cannot run here, so per-rank kernel OUTPUTS (packed keys, statistics buffers) are produced by
the oracle and the host-side key packing mirror; what is under test is the exchange logic itself: the
sign-flipped MIN all-reduce of packed uint64 keys (codebook-sharded arg-min and sync anchors), the fused
statistics all-reduce, and that the combined results equal the single-process oracle.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import oracle as O
from vector_quantization_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(rank, world, port, fn, out, ack):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        out.put((rank, fn(rank, world)))
        ack.wait(60)             # stay alive until the parent has read the (shared-memory) tensors
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    port = _free_port()
    ack = ctx.Event()
    procs = [ctx.Process(target=_run, args=(r, world, port, fn, out, ack)) for r in range(world)]
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


# ---- codebook-sharded assignment: per-shard keys -> all_reduce(MIN) -> global arg-min --------------
def _sharded_assign(rank, world):
    N, K, D = 512, 256, 16
    x, E = O.synthetic_latents(N, K, D, seed=7)
    lo, hi = parallel.shard_range(K, rank, world)
    score = x @ E[lo:hi].t() - 0.5 * (E[lo:hi] ** 2).sum(1)          # what vqb_assign maximises (L2)
    best, idx = score.max(1)
    # torch.max returns the first maximum -> lowest index inside the shard, like the kernel
    keys = parallel.pack_keys_host(best, idx + lo)                    # b_index_offset = lo
    parallel.all_reduce_min_keys_(keys)
    _, q = parallel.unpack_keys_host(keys)
    return q


def test_codebook_sharded_minloc_allreduce_equals_global_argmin():
    qs = _spawn(_sharded_assign)
    x, E = O.synthetic_latents(512, 256, 16, seed=7)
    q_ref, d = O.encode('L2', x, E)
    assert torch.equal(qs[0], qs[1])
    rows, gap = O.index_mismatch_report(d, q_ref, qs[0])
    assert (gap < 1e-5).all() and rows.numel() <= 2


# ---- sync=True anchors: global nearest token per code over the concatenated ranks -------------------
def _sync_anchor(rank, world):
    N, K, D = 300, 64, 16
    x_all, E = O.synthetic_latents(N * world, K, D, seed=11)
    x = x_all[rank * N:(rank + 1) * N]
    sim = F.normalize(E) @ F.normalize(x).t()                          # codes as rows, local tokens as columns
    best, tok = sim.max(1)
    keys = parallel.pack_keys_host(best, tok + rank * N)               # global token index = rank*N + n
    parallel.all_reduce_min_keys_(keys)
    _, gidx = parallel.unpack_keys_host(keys)
    local = gidx - rank * N
    mine = (local >= 0) & (local < N)
    anchors = torch.zeros(K, D)
    anchors[mine] = x[local[mine]]                                     # vqb_gather_rows_by_key: zero rows elsewhere
    parallel.all_reduce_sum_(anchors)
    return anchors, gidx


def test_sync_anchor_exchange_equals_allgather_semantics():
    res = _spawn(_sync_anchor)
    x_all, E = O.synthetic_latents(600, 64, 16, seed=11)
    d = O.cosine_distance(x_all, E)                                    # what the reference all_gathers (anchors.py:50-57)
    anchors_ref, idx_ref = O.nearest_anchor(x_all, d)
    for anchors, gidx in res:
        same = gidx == idx_ref
        assert same.float().mean() > 0.97                              # fp32 near-ties only
        assert torch.equal(anchors[same], anchors_ref[same])
    assert torch.equal(res[0][0], res[1][0])


# ---- token-sharded VQ-KD statistics: one fused [K*D | K] all-reduce ---------------------------------
def _vqkd_stats(rank, world):
    N, K, D = 400, 32, 8
    x_all, E = O.synthetic_latents(N * world, K, D, seed=3, normalized_codebook=True)
    x = F.normalize(x_all[rank * N:(rank + 1) * N])
    q, _ = O.encode('Cosine', x, E)
    stats = torch.zeros(K * D + K)                                      # layout of vqb_scatter_stats
    stats[:K * D].view(K, D).index_add_(0, q, x)
    stats[K * D:] += q.bincount(minlength=K).float()
    parallel.all_reduce_sum_(stats)
    cnt = stats[K * D:].unsqueeze(1)
    cent = torch.where(cnt > 0, stats[:K * D].view(K, D) / cnt.clamp_min(1), E)
    return F.normalize(O.ema(E, F.normalize(cent), 0.99)), q


def test_token_sharded_vqkd_update_equals_oracle():
    res = _spawn(_vqkd_stats)
    x_all, E = O.synthetic_latents(800, 32, 8, seed=3, normalized_codebook=True)
    xs = [x_all[:400], x_all[400:]]
    W_ref = O.vqkd_update(xs, [r[1] for r in res], E, 0.99)
    for W, _ in res:
        torch.testing.assert_close(W, W_ref, rtol=1e-5, atol=1e-6)


def test_single_process_helpers_are_noops():
    t = torch.arange(4.)
    assert parallel.world_size() == 1 and parallel.rank() == 0
    assert torch.equal(parallel.all_reduce_sum_(t.clone()), t)
    k = parallel.pack_keys_host(torch.tensor([0.5]), torch.tensor([3]))
    assert torch.equal(parallel.all_reduce_min_keys_(k.clone()), k)
