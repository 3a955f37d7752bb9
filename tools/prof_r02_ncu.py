"""r02 ncu driver: one or two launches of every dominant kernel at bench size (16 pairs, 480x640) in the SHIPPED
variants (fp16 intermediates, fp16 backbone), for `ncu --set full -k regex:gf:: ...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops, engine
from geoformer_b200.engine import pack_conv3x3
from geoformer_b200.ops import EPI_ELU1, EPI_LN, EPI_RELU
dev = torch.device("cuda:0"); ops.ensure_init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
R = lambda *s: torch.randn(*s, device=dev, generator=g)
M, C = 32 * 4800, 256
x, m1 = R(M, C), R(M, C)
wqkv, wm16, w1, w2_16 = R(3 * C, C) / 16, (R(C, C) / 16).half(), R(2 * C, 2 * C) / 22, (R(C, 2 * C) / 22).half()
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
f0, f1 = R(16, 4800, C) * 3 + 1.5, R(16, 4800, C) * 3 + 1.5
n, l, hh, dd, cnt = 16, 4800, 4, 64, 4000
qkv16 = R(n * l, 3 * C).half()
aidx = torch.stack([torch.sort(torch.randperm(l, device=dev, generator=g)[:cnt])[0] for _ in range(n)]).int()
acnt = torch.full((n,), cnt, device=dev, dtype=torch.int32)
wins = 130898
wq, wk, wv, wmf = (torch.randn(128, 128) * 128 ** -0.5 for _ in range(4))
wp = engine.pack_fine_layer(wq, wk, wv, wmf, torch.randn(256, 256) / 16, torch.randn(128, 256) / 16, dev)
g1, b1 = torch.ones(128, device=dev), torch.zeros(128, device=dev)
xf, sf_ = R(wins, 25, 128), R(wins, 25, 128)
convs = []
for (cin, cout, cpi, cpo, h, w) in [(128, 128, 128, 128, 240, 320), (196, 196, 200, 200, 240, 320), (256, 256, 256, 256, 120, 160)]:
    convs.append((R(32, h, w, cpi).half(), *pack_conv3x3(torch.randn(cout, cin, 3, 3) * 0.03, torch.zeros(cout), cpi, cpo, dev)))
for _ in range(2):
    # projection GEMMs of one coarse layer, shipped variants
    q = ops.linear(x, wqkv, epi=EPI_ELU1, act_cols=2 * C, out_f16=True)                    # 768x256 > h
    msg = q[:, :C].contiguous()
    n1 = ops.linear(msg, wm16, epi=EPI_LN, gamma=gam, beta=bet)                             # 256x256 h, LayerNorm
    hid = ops.linear(x, w1, a2=n1, epi=EPI_RELU, out_f16=True)                              # 512x512 > h
    y = ops.linear(hid, w2_16, epi=EPI_LN, gamma=gam, beta=bet, residual=x)                 # 256x512 h, LayerNorm + residual
    ops.coarse_match_fused(f0, f1, 0.1, 0.0, 0, (60, 80), (60, 80), 8.0)
    ops.geo_self_attention(qkv16, 3 * C, qkv16[:, C:], 3 * C, qkv16[:, 2 * C:], 3 * C, n, l, hh, dd, aidx, acnt, max_cnt=cnt, impl="tf32")
    ops.fine_layer_fused(xf, xf, wp, g1, b1, g1, b1)
    ops.fine_layer_fused(xf, sf_, wp, g1, b1, g1, b1)
    for cx, wt, bias in convs:
        ops.conv3x3(cx, wt, bias, None, 1)
torch.cuda.synchronize()
