"""Micro-driver for the ncu --set full capture of the headline kernels at bench size (16 pairs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops
from geoformer_b200.engine import pack_conv3x3
dev = torch.device("cuda:0"); ops.ensure_init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
f0 = torch.randn(16, 4800, 256, device=dev, generator=g) * 3 + 1.5
f1 = torch.randn(16, 4800, 256, device=dev, generator=g) * 3 + 1.5
x = torch.randn(32, 240, 320, 128, device=dev, generator=g).half()
wt, bias = pack_conv3x3(torch.randn(128, 128, 3, 3) * 0.03, torch.zeros(128), 128, 128, dev)
a = torch.randn(153600, 256, device=dev, generator=g)
w = torch.randn(512, 512, device=dev, generator=g) / 22
n, l, c, hh, dd, cnt = 16, 4800, 256, 4, 64, 4000
qkv = torch.randn(n * l, 3 * c, device=dev, generator=g)
aidx = torch.stack([torch.sort(torch.randperm(l, device=dev, generator=g)[:cnt])[0] for _ in range(n)]).int()
acnt = torch.full((n,), cnt, device=dev, dtype=torch.int32)
for _ in range(3):
    att = ops.geo_self_attention(qkv, 3 * c, qkv[:, c:], 3 * c, qkv[:, 2 * c:], 3 * c, n, l, hh, dd, aidx, acnt, max_cnt=cnt, impl="tf32")
    sim = ops.similarity(f0, f1, 0.1)
    mres, mcounts = ops.coarse_match_fused(f0, f1, 0.1, 0.0, 0, (60, 80), (60, 80), 8.0)
    y = ops.conv3x3(x, wt, bias, None, 1)
    h = ops.linear(a, w, a2=a, epi=ops.EPI_RELU)
torch.cuda.synchronize()
