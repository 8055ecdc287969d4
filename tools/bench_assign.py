"""Dev micro-benchmark of vqb_assign (CUDA events, L2 flushed between iterations)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from vector_quantization_b200 import ops

dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


cases = [  # name, N, K, D, metric, x dtype, planes_x, planes_e
    ('cfg2 cos bf16x/3p', 65536, 8192, 32, 'cos', torch.bfloat16, 1, 3),
    ('cfg2 cos bf16x/pair', 65536, 8192, 32, 'cos', torch.bfloat16, 1, 'pair'),
    ('cfg2 cos bf16x/2p', 65536, 8192, 32, 'cos', torch.bfloat16, 1, 2),
    ('cfg2 cos bf16x/1p', 65536, 8192, 32, 'cos', torch.bfloat16, 1, 1),
    ('cfg3 l2 D8 3x3', 65536, 16384, 8, 'l2', torch.float32, 3, 3),
    ('cfg3 l2 D8 1x1', 65536, 16384, 8, 'l2', torch.bfloat16, 1, 1),
    ('cfg4 cos D256 1x3', 16384, 8192, 256, 'cos', torch.bfloat16, 1, 3),
    ('cfg4 cos D256 1xpair', 16384, 8192, 256, 'cos', torch.bfloat16, 1, 'pair'),
    ('cfg4 cos D256 1x1', 16384, 8192, 256, 'cos', torch.bfloat16, 1, 1),
    ('cfg1 l2 D256 3x3', 16384, 8192, 256, 'l2', torch.float32, 3, 3),
    ('cfg5/8 cos D768 1x1', 65536, 32768, 768, 'cos', torch.bfloat16, 1, 1),
    ('cfg5/8 cos D768 1xpair', 65536, 32768, 768, 'cos', torch.bfloat16, 1, 'pair'),
]
out = []
flt = sys.argv[1] if len(sys.argv) > 1 else ''
for name, N, K, D, metric, dt, px, pe in cases:
    if flt not in name:
        continue
    x = torch.randn(N, D, device=dev).to(dt)
    E = torch.randn(K, D, device=dev)
    cos = metric == 'cos'
    a = ops.pack_rows(x, planes=px) if pe != 'pair' else ops.pack_rows(x, fmt='f16')
    pair = pe == 'pair'
    b = ops.pack_rows(E, normalize=cos, planes=None if pair else pe, want_half_sqnorm=not cos, fmt='f16x2' if pair else 'bf16')
    keys = ops.new_keys(N, dev)
    med, best = timeit(lambda: ops.assign(a, b, keys, l2=not cos))
    terms = {(1, 1): 1, (1, 2): 2, (1, 3): 3, (3, 3): 6, (2, 2): 3, (1, 'pair'): 2}[(px, pe)]
    flops = 2.0 * N * K * D
    rec = dict(case=name, ms_median=round(med, 4), ms_best=round(best, 4), terms=terms,
               algo_tflops=round(flops / med / 1e9, 1), mma_tflops=round(flops * terms * (ops.operand_shape(1, D)[1] / D) / med / 1e9, 1),
               gelem_per_s=round(N * K / med / 1e6, 1), mtok_per_s=round(N / med / 1e3, 1))
    if not cos and ops.can_fold_l2(D):
        # the same assignment with the side terms folded into the spare operand columns (vqb_fold_l2_side)
        af = ops.fold_l2_side(ops.pack_rows(x, planes=px), 'tokens')
        bf = ops.fold_l2_side(ops.pack_rows(E, planes=pe, want_half_sqnorm=True), 'codes')
        m_f, b_f = timeit(lambda: ops.assign(af, bf, keys, l2=True))
        rec['folded_ms_median'], rec['folded_ms_best'] = round(m_f, 4), round(b_f, 4)
        rec['fold_launch_ms'] = round(timeit(lambda: ops._call('vqb_fold_l2_side', ops._lib.load().vqb_fold_l2_side,
                                                                 bf.planes.device, ops._p(bf.planes), bf.nplanes, bf.rows, bf.dim,
                                                                 ops._p(bf.half_sqnorm), 1, ops._S), iters=5)[0], 4)
    pk = timeit(lambda: ops.pack_rows(E, normalize=cos, planes=None if pair else pe, want_half_sqnorm=not cos, fmt='f16x2' if pair else 'bf16'), iters=5)[0]
    rec['pack_codebook_ms'] = round(pk, 4)
    if pair and ops.operand_shape(1, D)[1] >= 128:
        # certified one-term pass (hi plane + certificate + exact re-run of the uncertified rows), random AND clustered
        from vector_quantization_b200 import functional as Fq
        b2 = ops.pack_rows(E, normalize=True, fmt='f16x2', want_lo_norm=True)
        for tag, xs in (('random', x), ('clustered', (torch.nn.functional.normalize(E)[torch.randint(0, K, (N,), device=dev)]
                                                      + 0.5 / D ** 0.5 * torch.randn(N, D, device=dev)).to(dt))):
            a2 = ops.pack_rows(xs, fmt='f16')
            a2.inv_norm = ops.row_inv_norm(xs, f16_rows=True)

            def cert():
                keys.fill_(-1)
                Fq.certified_assign(a2, b2, keys, a_inv_norm=a2.inv_norm)

            def plain():
                keys.fill_(-1)
                ops.assign(a2, b2, keys, l2=False)
            m_c, _ = timeit(cert)
            frac = float(Fq.LAST_CERTIFY['count']) / N
            m_p, _ = timeit(plain)
            rec[f'certified_{tag}_ms'] = round(m_c, 4)
            rec[f'two_term_{tag}_ms'] = round(m_p, 4)
            rec[f'uncertified_fraction_{tag}'] = round(frac, 4)
            rec[f'certified_{tag}_algo_tflops'] = round(flops / m_c / 1e9, 1)
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del x, E, a, b, keys
json.dump(out, open('gpurun_out/bench_assign.json', 'w'), indent=1)
