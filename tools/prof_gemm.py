"""Micro-driver for ncu: a handful of GEMM launches at bench sizes (linear tf32 + split-fp16 similarity)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops

dev = torch.device("cuda:0")
ops.ensure_init(dev)
M = 153600
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, 256, device=dev, generator=g)
x2 = torch.randn(M, 256, device=dev, generator=g)
w256 = torch.randn(256, 256, device=dev, generator=g) / 16
w768 = torch.randn(768, 256, device=dev, generator=g) / 16
w512 = torch.randn(512, 512, device=dev, generator=g) / 22
gam, bet = torch.ones(256, device=dev), torch.zeros(256, device=dev)
f0 = torch.randn(4, 4800, 256, device=dev, generator=g) * 3 + 1.5
f1 = torch.randn(4, 4800, 256, device=dev, generator=g) * 3 + 1.5

def timeit(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n*1e3:9.1f} us", flush=True)

timeit("linear 256x256 plain  M=153600", lambda: ops.linear(x, w256))
timeit("linear 256x256 LN+res M=153600", lambda: ops.linear(x, w256, epi=ops.EPI_LN, gamma=gam, beta=bet, residual=x2))
timeit("linear 768x256 elu    M=153600", lambda: ops.linear(x, w768, epi=ops.EPI_ELU1, act_cols=512))
timeit("linear 512x512 relu   M=153600", lambda: ops.linear(x, w512, a2=x2, epi=ops.EPI_RELU))
timeit("similarity n=4 (incl. pack)", lambda: ops.similarity(f0, f1, 0.1))
timeit("torch matmul fp32(tf32 off) 256x256", lambda: x @ w256.T)
