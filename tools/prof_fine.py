"""Micro-driver: fine-level GEMM shapes + window attention + fine match at bench size (131K windows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
m = 131072
x = torch.randn(m * 25, 128, device=dev, generator=g)
wqkv = torch.randn(384, 128, device=dev, generator=g) / 11
w1 = torch.randn(256, 256, device=dev, generator=g) / 16
w2 = torch.randn(128, 256, device=dev, generator=g) / 16
wm = torch.randn(128, 128, device=dev, generator=g) / 11
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
def timeit(name, fn, k=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/k*1e3:9.1f} us", flush=True)
    return out
qkv = timeit("qkv 384x128 elu", lambda: ops.linear(x, wqkv, epi=ops.EPI_ELU1, act_cols=256))
att = timeit("linattn_window", lambda: ops.linattn_window(qkv, 384, qkv[:, 128:], 384, qkv[:, 256:], 384, m, 25, 8, 16))
m1 = timeit("merge 128x128 LN", lambda: ops.linear(att, wm, epi=ops.EPI_LN, gamma=gam, beta=bet))
h = timeit("mlp1 256x(128+128) relu", lambda: ops.linear(x, w1, a2=m1, epi=ops.EPI_RELU))
y = timeit("mlp2 128x256 LN+res", lambda: ops.linear(h, w2, epi=ops.EPI_LN, gamma=gam, beta=bet, residual=x))
f0 = y.view(m, 25, 128)[: m // 2].contiguous(); f1 = y.view(m, 25, 128)[m // 2:].contiguous()
k0 = torch.zeros(m // 2, 2, device=dev); b = torch.zeros(m // 2, device=dev, dtype=torch.int64)
timeit("fine_match 65K", lambda: ops.fine_match(f0, f1, 0.1, 0.1, k0, k0, b, 5, 8.0, 4.0, 2.0))
