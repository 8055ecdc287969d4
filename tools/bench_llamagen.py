"""The reference's LlamaGen tokenizer quantizer (configs/llamagen/vqgan.py:10-20: 16384 x 8 codebook, L2 distance on
l2-normalised tokens and codes, VQGAN loss) as a module step: forward + backward of N tokens (default 262 144: at 65 536 the eager step is bound by the host-side launches), CUDA events, with
and without the L2 side terms folded into the contraction.
    python tools/bench_llamagen.py"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch

import vector_quantization_b200 as vqb
from vector_quantization_b200 import functional as Fq

dev = torch.device('cuda', 0)
N, K, D = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 16384, 8
cfg = dict(type='VQGANQuantizer', embedding=dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=K, embedding_dim=D),
           distance=dict(type='L2Distance'), callbacks=[dict(type='NormalizeCallback')],
           losses=dict(vqgan_loss=dict(type='VQGANLoss')), init_weights=dict(type='vqgan'))
q = vqb.build_quantizer(cfg, training=True).to(dev)
with torch.no_grad():
    q.embedding.weight.copy_(torch.randn(K, D, device=dev))
x0 = torch.randn(N, D, device=dev, dtype=torch.bfloat16)
gz = torch.randn(N, D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step():
    x = x0.clone().requires_grad_(True)
    z, loss, _ = q(x, {})
    (loss + (z * gz).sum()).backward()
    q.embedding.weight.grad = None
    return loss


out = {}
for fold in (True, False, True, False):
    Fq.FOLD_L2 = fold
    for _ in range(5):
        step()
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    out.setdefault('folded' if fold else 'side_mode_1', []).append(round(ts[len(ts) // 2], 4))
print(json.dumps(dict(case=f'LlamaGen quantizer module step (eager launches), {N} x 16384 x 8, L2 + NormalizeCallback + VQGANLoss',
                      ms_median=out)))
