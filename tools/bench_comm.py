"""Microbenchmark of the fused peer-memory exchange kernels against NCCL all_reduce + update (N ranks, one node):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_comm.py
Back-to-back launches (no host skew between ranks), CUDA events, max over ranks."""
import json
import os
import pathlib
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from vector_quantization_b200 import ops, parallel  # noqa: E402


def timed(fn, iters=200, warm=20):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


def graphed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    return lambda: g.replay(), reps


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    out = {}
    for name, K, D in (('cfg2', 8192, 32), ('cfg3', 16384, 8), ('cfg4', 8192, 256)):
        region = parallel.PeerRegion((K * D * 2 + K) * 4 + 4096, dev)
        W = region.alloc('W', (K, D), torch.float32)
        stats = region.alloc('stats', (K * D + K,), torch.float32)
        W.copy_(torch.nn.functional.normalize(torch.randn(K, D, device=dev)))
        stats.copy_(torch.rand(K * D + K, device=dev))
        dist.barrier()
        res = {}
        W0, st0 = W.clone(), stats.clone()
        if (K * D + K) * 4 <= (4 << 20):
            # low-latency (flag-in-data) protocol in a second region; same inputs -> bit-identical result
            r2 = parallel.PeerRegion((K * D * 2 + K) * 4 + (K * D + K + 64 * world) * 16 + 8192, dev)
            W2 = r2.alloc('W', (K, D), torch.float32)
            s2 = r2.alloc('stats', (K * D + K,), torch.float32)
            for nm, shp, dt in ops.comm_ll_layout(K, D, world):
                r2.alloc(nm, shp, dt)
            W2.copy_(W0); s2.copy_(st0)
            dist.barrier()
            ops.comm_kmeans_ema_update(region, K, D, 0.99)
            ops.comm_kmeans_ema_update(r2, K, D, 0.99)
            torch.cuda.synchronize()
            res['ll_equals_barrier'] = bool(torch.equal(W, W2))
            ref = st0.clone(); dist.all_reduce(ref); Wr = W0.clone(); ops.kmeans_ema_update(ref, Wr, 0.99)
            res['ll_max_abs_diff_vs_nccl'] = float((W2 - Wr).abs().max())
            res['ll_eager_us'] = timed(lambda: ops.comm_kmeans_ema_update(r2, K, D, 0.99))
            fn, reps = graphed(lambda: ops.comm_kmeans_ema_update(r2, K, D, 0.99))
            res['ll_graph_us'] = timed(fn, iters=20, warm=3) / reps
        res['fused_eager_us'] = timed(lambda: ops.comm_kmeans_ema_update(region, K, D, 0.99))
        fn, reps = graphed(lambda: ops.comm_kmeans_ema_update(region, K, D, 0.99))
        res['fused_graph_us'] = timed(fn, iters=20, warm=3) / reps
        Wn = W.clone()
        st2 = stats.clone()

        def nccl():
            dist.all_reduce(st2)
            ops.kmeans_ema_update(st2, Wn, 0.99)
        res['nccl_eager_us'] = timed(nccl)
        fn, reps = graphed(nccl)
        res['nccl_graph_us'] = timed(fn, iters=20, warm=3) / reps
        res['sum_f32_graph_us'] = None
        fn, reps = graphed(lambda: ops.comm_allreduce_sum_f32(region, K * D + K, 'stats'))
        res['sum_f32_graph_us'] = timed(fn, iters=20, warm=3) / reps
        res['bytes'] = (K * D + K) * 4
        out[name] = res
    if rank == 0:
        print(json.dumps(dict(world=world, blocks_per_sm=os.environ.get('VQB_COMM_BLOCKS_PER_SM', '4'), results=out)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
