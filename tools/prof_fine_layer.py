"""Micro-driver: fused fine-level layer (gf_fine_layer) at bench size; GF_FL_DEBUG bits isolate the phases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops, engine
dev = torch.device("cuda:0"); ops.ensure_init(dev)
g = torch.Generator().manual_seed(0)
r = lambda *s, sc=1.0: torch.randn(*s, generator=g) * sc
wq, wk, wv, wm = (r(128, 128, sc=128 ** -0.5) for _ in range(4))
w1, w2 = r(256, 256, sc=1 / 16), r(128, 256, sc=1 / 16)
wp = engine.pack_fine_layer(wq, wk, wv, wm, w1, w2, dev)
gam, bet = torch.ones(128, device=dev), torch.zeros(128, device=dev)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 112000
x = torch.randn(m, 25, 128, device=dev); s = torch.randn(m, 25, 128, device=dev)
def timeit(name, fn, k=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / k
    tiles = (m + 4) // 5
    print(f"{name:28s} {t*1e3:9.1f} us   {t*1e3/ (tiles/148):7.2f} us/tile/SM", flush=True)
tag = os.environ.get("GF_FL_DEBUG", "0")
timeit(f"self  debug={tag}", lambda: ops.fine_layer_fused(x, x, wp, gam, bet, gam, bet))
timeit(f"cross debug={tag}", lambda: ops.fine_layer_fused(x, s, wp, gam, bet, gam, bet))
