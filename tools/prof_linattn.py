"""Micro-driver: coarse linear attention (fp16 storage, mma.sync kernels) at bench size: 32 samples x 4800 tokens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_init(dev)
n, l, h, d = 32, 4800, 8, 32
c = h * d
qkv = (torch.rand(n * l, 3 * c, device=dev) + 0.5).half()
def timeit(name, fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/k*1e3:9.1f} us", flush=True)
timeit("linattn reduce+apply (fp16, 32x4800)", lambda: ops.linattn(qkv, 3 * c, qkv[:, c:], 3 * c, qkv[:, 2 * c:], 3 * c, n, l, l, h, d))
