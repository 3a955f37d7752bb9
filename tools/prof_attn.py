"""Micro-driver: geo self-attention (tensor-core path) and the n=16 similarity at bench sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops

dev = torch.device("cuda:0")
ops.ensure_init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
n, l, c, h, d, cnt = 16, 4800, 256, 4, 64, int(os.environ.get("CNT", "4000"))
qkv = torch.randn(n * l, 3 * c, device=dev, generator=g)
aidx = torch.stack([torch.sort(torch.randperm(l, device=dev, generator=g)[:cnt])[0] for _ in range(n)]).int()
acnt = torch.full((n,), cnt, device=dev, dtype=torch.int32)
f0 = torch.randn(n, l, c, device=dev, generator=g) * 3 + 1.5
f1 = torch.randn(n, l, c, device=dev, generator=g) * 3 + 1.5

def timeit(name, fn, k=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1)/k*1e3:9.1f} us", flush=True)

for mode in ("tf32", "tf32_mat"):
    timeit(f"geo self-attention {mode} (n=16, {cnt} anchors)", lambda: ops.geo_self_attention(qkv, 3*c, qkv[:, c:], 3*c, qkv[:, 2*c:], 3*c, n, l, h, d, aidx, acnt, max_cnt=cnt, impl=mode))
ops.PROFILE.enable("*")
ops.geo_self_attention(qkv, 3*c, qkv[:, c:], 3*c, qkv[:, 2*c:], 3*c, n, l, h, d, aidx, acnt, max_cnt=cnt, impl="tf32")
ops.similarity(f0, f1, 0.1)
sim = ops.similarity(f0, f1, 0.1)
ops.dual_softmax_(sim)
for k, (cn, ms) in ops.PROFILE.summary().items():
    print(f"  {k:34s} x{cn:3d} {ms/cn*1e3:9.1f} us each")
ops.PROFILE.disable()
