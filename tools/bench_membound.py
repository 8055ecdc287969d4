"""Dev micro-benchmark of the HBM-bound kernels: achieved GB/s on ALGORITHMIC bytes against
the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  CUDA events, 1 GiB memset between launches
(flushes L2 and keeps the GPU busy while the host enqueues)."""
import json, pathlib, sys
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from vector_quantization_b200 import ops

dev = torch.device('cuda', 0)
peak = json.load(open(pathlib.Path(__file__).resolve().parents[1] / 'MEASURED_PEAKS.json')).get('hbm_gbs', 6548.2) \
    if (pathlib.Path(__file__).resolve().parents[1] / 'MEASURED_PEAKS.json').exists() else 6548.2
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)


def timeit(fn, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = []


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    rec = dict(kernel=name, ms=round(ms, 4), algorithmic_MB=round(nbytes / 1e6, 2), GBps=round(gbs, 1),
               frac_of_measured_hbm=round(gbs / peak, 3))
    print(json.dumps(rec), flush=True)
    out.append(rec)


for N, K, D, dt in ((1 << 20, 8192, 32, torch.bfloat16), (1 << 20, 8192, 32, torch.float32),
                    (1 << 18, 8192, 256, torch.bfloat16), (1 << 22, 16384, 8, torch.bfloat16)):
    sx = 2 if dt == torch.bfloat16 else 4
    x = torch.randn(N, D, device=dev).to(dt)
    W = torch.nn.functional.normalize(torch.randn(K, D, device=dev))
    q = torch.randint(0, K, (N,), device=dev)
    gz = torch.randn(N, D, device=dev)
    g4 = torch.tensor([0., 0., 0., 1.], device=dev)
    tag = f'N={N} D={D} {"bf16" if sx == 2 else "fp32"}'
    ms = timeit(lambda: ops.gather_ste_loss(x, W, quant=q, normalize_x=True, want_norm=True))
    report(f'quantize_forward (norm) {tag}', ms, N * (D * (sx + 4 + 4) + 8))          # x, W row, z, idx
    ms = timeit(lambda: ops.quantize_backward(gz, x, W, q, g4, normalize_x=True, want_norm=True, need_gW=False))
    report(f'quantize_backward (norm) {tag}', ms, N * (D * (sx + 4 + 4 + sx) + 8))    # x, W row, gz, gx, idx
    stats = torch.zeros(K * D + K, device=dev)
    ms = timeit(lambda: ops.scatter_stats(x, q, K, normalize_x=True, out=stats))
    report(f'scatter_stats {tag}', ms, N * (D * sx + 8))
    ms = timeit(lambda: ops.pack_rows(x, normalize=True, planes=3, want_half_sqnorm=True))
    report(f'pack_rows 3 planes {tag}', ms, N * (D * sx + 3 * ops.operand_shape(1, D)[1] * 2 + 4))
    del x, W, q, gz, stats

# caller-side layout kernels: NCHW <-> token-major transposes, fp16 token plane, compact token ids
for B, C, HW, dt in ((4096, 32, 256, torch.bfloat16), (4096, 32, 256, torch.float32), (1024, 256, 256, torch.bfloat16),
                     (16384, 8, 256, torch.float32)):
    sx = 2 if dt == torch.bfloat16 else 4
    x = torch.randn(B, C, HW, device=dev).to(dt)
    ms = timeit(lambda: ops.transpose_last2(x))
    report(f'transpose nchw->rows B={B} C={C} HW={HW} {"bf16" if sx == 2 else "fp32"}', ms, 2 * x.numel() * sx)
    rows = ops.transpose_last2(x)
    ms = timeit(lambda: ops.transpose_last2(rows))
    report(f'transpose rows->nchw B={B} C={C} HW={HW} {"bf16" if sx == 2 else "fp32"}', ms, 2 * x.numel() * sx)
    del x, rows
xb = torch.randn(1 << 20, 32, device=dev).to(torch.bfloat16)
ms = timeit(lambda: ops.pack_rows(xb, fmt='f16'))
report('pack_rows one fp16 plane N=1048576 D=32 bf16', ms, xb.numel() * 4)
Wn = torch.randn(1 << 18, 32, device=dev)
ms = timeit(lambda: ops.pack_rows(Wn, normalize=True, fmt='f16x2'))
report('pack_rows fp16 pair N=262144 D=32 fp32', ms, Wn.numel() * (4 + 4))
keys = torch.randint(0, 8192, (1 << 22,), device=dev)
ms = timeit(lambda: ops.compact_tokens(keys, 8192))
report('compact_tokens N=4194304 uint16', ms, keys.numel() * 10)
del xb, Wn, keys

for levels in ([8, 8, 5, 5, 5], [8, 8, 8, 5, 5, 5]):
    N = 1 << 22
    D = len(levels)
    p = ops.fsq_params(levels, 1e-3)
    for dt in (torch.float32, torch.bfloat16):
        sx = 2 if dt == torch.bfloat16 else 4
        x = (1.5 * torch.randn(N, D, device=dev)).to(dt)
        ms = timeit(lambda: ops.fsq_forward(x, p))
        report(f'fsq_forward N={N} D={D} {"bf16" if sx == 2 else "fp32"}', ms, N * (D * 2 * sx + 4))
json.dump(out, open('gpurun_out/bench_membound.json', 'w'), indent=1)
