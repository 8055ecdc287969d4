"""Host-side cost of the EAGER (no CUDA graph) training step of cfg 2: cProfile over 300 steps, top entries by
self time.  python tools/profile_eager.py"""
import cProfile
import io
import pathlib
import pstats
import sys
import time

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch

import bench
import vector_quantization_b200 as vqb

dev = torch.device('cuda', 0)
wl = bench.WORKLOADS['cfg2']
N, K, D = wl['N'], wl['K'], wl['D']
q = vqb.build_quantizer(dict(wl['config'], embedding=bench.emb(K, D)), training=True).to(dev)
q._forward_pre_hooks.clear()
q.requires_grad_(False)
x0, E, gz0 = bench.synth(N, K, D, bench.SEED)
with torch.no_grad():
    q.embedding.weight.copy_(E)
x = x0.to(dev).requires_grad_(True)
gz = gz0.to(dev)
one = torch.ones([], device=dev)


def step():
    z, loss, memo = q(x, dict())
    (gx,) = torch.autograd.grad((z, loss), (x,), (gz, one))
    return gx


for _ in range(20):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(300):
    step()
host = (time.perf_counter() - t) / 300
torch.cuda.synchronize()
total = (time.perf_counter() - t) / 300
print(f'host time per step {host * 1e3:.3f} ms, wall incl. drain {total * 1e3:.3f} ms')
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(38)
print('\n'.join(line[:150] for line in s.getvalue().splitlines()[:60]))
