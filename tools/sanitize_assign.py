"""Small launches of every assign_tc_kernel mode for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_assign.py
Covers: whole-tile pipeline (D <= 64) with resident A slots, in-kernel bf16->fp16 conversion, the k-blocked ring
(D >= 128), L2 side term, column scale, the certified one-term epilogue and the device-side row count."""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from oracle import oracle as O  # noqa: E402
from vector_quantization_b200 import functional as Fq  # noqa: E402
from vector_quantization_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)


def check(name, got, x, E, metric):
    q_ref, d = O.encode(metric, x, E)
    rows, gap = O.index_mismatch_report(d, q_ref, got.cpu())
    ok = bool((gap < 1e-4).all())
    print(f'{name}: {"ok" if ok else "MISMATCH"} ({rows.numel()} near-tie rows)')
    assert ok


# 1. whole-tile, cosine, zero-copy bf16 tokens x fp16 pair (the bench path), partial last tiles
x, E = O.synthetic_latents(700, 1100, 32, seed=1, normalized_codebook=True)
xb = x.to(torch.bfloat16).to(dev)
book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2')
keys = ops.new_keys(700, dev)
ops.assign(ops.as_operand(xb), book, keys, l2=False)
check('whole/pair/convert', ops.unpack_keys(keys), xb.float().cpu(), E, 'Cosine')
# 2. whole-tile, L2, 3x3 exact planes, D = 8
x, E = O.synthetic_latents(600, 900, 8, seed=2)
keys = ops.new_keys(600, dev)
ops.assign(ops.pack_rows(x.to(dev)), ops.pack_rows(E.to(dev), want_half_sqnorm=True), keys, l2=True)
check('whole/l2/3x3', ops.unpack_keys(keys), x, E, 'L2')
# 3. k-blocked ring, D = 256, pair, plain + certified (row and column pass)
x, E = O.synthetic_latents(500, 700, 256, seed=3, normalized_codebook=True, clustered=False)
xb = x.to(torch.bfloat16).to(dev)
book = ops.pack_rows(E.to(dev), normalize=True, fmt='f16x2', want_lo_norm=True)
toks = ops.pack_rows(xb, fmt='f16')
toks.inv_norm = ops.row_inv_norm(xb, f16_rows=True)
keys = ops.new_keys(500, dev)
ops.assign(toks, book, keys, l2=False)
check('ring/pair', ops.unpack_keys(keys), xb.float().cpu(), E, 'Cosine')
keys = ops.new_keys(500, dev)
Fq.certified_assign(toks, book, keys, a_inv_norm=toks.inv_norm)
check('ring/certified rows', ops.unpack_keys(keys), xb.float().cpu(), E, 'Cosine')
ck = ops.new_keys(700, dev)
Fq.certified_assign(book, toks, ck, scale_columns=True)
torch.cuda.synchronize()
print('ring/certified columns: launched, flagged', int(Fq.LAST_CERTIFY['count']))
print('done')
