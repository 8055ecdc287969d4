import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from oracle import oracle as O
from vector_quantization_b200 import ops
dev = torch.device('cuda', 0)
N, K, D = 300, 130, 8
x, E = O.synthetic_latents(N, K, D, seed=N + K)
d64 = torch.cdist(x.double(), E.double())
d_cpu = torch.cdist(x, E)
d_gpu = ops.distance_matrix(x.to(dev), E.to(dev), 'L2').cpu()
d_gpu_t = torch.cdist(x.to(dev), E.to(dev)).cpu()
for name, d in (('cpu cdist', d_cpu), ('kernel', d_gpu), ('gpu cdist', d_gpu_t)):
    err = (d.double() - d64).abs()
    print(name, 'max', err.max().item(), 'frac>2e-5', (err > 2e-5).float().mean().item())
err = (d_gpu.double() - d64).abs()
bad = (err > 2e-5).nonzero()
print('bad rows hist', torch.bincount(bad[:, 0] % 64, minlength=64).tolist())
print('bad cols hist', torch.bincount(bad[:, 1] % 64, minlength=64).tolist())
print(torch.get_float32_matmul_precision(), torch.__config__.parallel_info()[:200])
