import sys, pathlib, ctypes
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
from vector_quantization_b200 import ops, _lib
dev = torch.device('cuda', 0)
N, K, D = 65536, 8192, 32
x = torch.randn(N, D, device=dev).to(torch.bfloat16)
E = torch.randn(K, D, device=dev)
pe = sys.argv[1] if len(sys.argv) > 1 else '1'
if pe == 'pair':     # bench.py default: zero-copy bf16 tokens x fp16-pair codebook
    a = ops.as_operand(x)
    b = ops.pack_rows(E, normalize=True, fmt='f16x2')
else:
    a = ops.pack_rows(x, planes=1)
    b = ops.pack_rows(E, normalize=True, planes=int(pe))
keys = ops.new_keys(N, dev)
for _ in range(3):
    ops.assign(a, b, keys, l2=False)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (3 * 64 * 8))()
lib.vqb_debug_timeline.argtypes = [ctypes.c_void_p]
print('rc', lib.vqb_debug_timeline(buf))
ts = torch.tensor(list(buf)).view(3, 64, 8)
t0 = int(ts[ts > 0].min())
rel = (ts - t0).clamp_min(-1)
print('kernel start', int(rel[0, 63, 3]))
print('tile | producer: wait_start got_empty tma_issued | mma: start got_tmem_empty got_full loop_end elected mma_issued commit1 commit2 | epi: got_full tile_done flushed')
for t in range(0, 64):
    print(t, rel[0, t, :3].tolist(), rel[1, t].tolist(), rel[2, t, :4].tolist())
