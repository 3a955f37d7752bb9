"""Host <-> kernel layout contracts that need no GPU: the operand-block stream of the fused fine layer
(geoformer_b200.engine.pack_fine_layer <-> csrc/fine_layer.cu) and the tap-major convolution weights
(pack_conv3x3 <-> csrc/conv_tc.cu), plus BN folding (eval-mode BatchNorm of resnet_fpn.py:32-40 folded into the conv)."""
import numpy as np
import torch
import torch.nn.functional as F

from geoformer_b200 import engine


def _rnd(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def test_pack_fine_layer_block_stream():
    wq, wk, wv, wm = (_rnd(128, 128, seed=s) for s in range(4))
    w1, w2 = _rnd(256, 256, seed=4), _rnd(128, 256, seed=5)
    pk = engine.pack_fine_layer(wq, wk, wv, wm, w1, w2, "cpu")
    assert pk.dtype == torch.uint8 and tuple(pk.shape) == (30, 128, 128)         # 30 blocks of 128 rows x 128 B
    f32 = lambda b: pk[b].contiguous().view(torch.float32)                       # [128, 32]
    f16 = lambda b: pk[b].contiguous().view(torch.float16)                       # [128, 64]
    wqkv = torch.cat([wq, wk, wv], 0)
    blk = 0
    for nc in range(3):                       # GEMM0: q | k | v, tf32, 4 k-blocks of 32 floats
        for kb in range(4):
            assert torch.equal(f32(blk), wqkv[nc * 128:(nc + 1) * 128, kb * 32:(kb + 1) * 32]); blk += 1
    for nc in range(2):                       # MLP up, x half (tf32)
        for kb in range(4):
            assert torch.equal(f32(blk), w1[nc * 128:(nc + 1) * 128, kb * 32:(kb + 1) * 32]); blk += 1
    for kb in range(2):                       # merge (fp16, 64 halves per k-block)
        assert torch.equal(f16(blk), wm[:, kb * 64:(kb + 1) * 64].half()); blk += 1
    for nc in range(2):                       # MLP up, message half (fp16)
        for kb in range(2):
            assert torch.equal(f16(blk), w1[nc * 128:(nc + 1) * 128, 128 + kb * 64:128 + (kb + 1) * 64].half()); blk += 1
    for kb in range(4):                       # MLP down (fp16)
        assert torch.equal(f16(blk), w2[:, kb * 64:(kb + 1) * 64].half()); blk += 1
    assert blk == 30


def test_pack_conv_weights_tap_major_and_padding():
    for k, cin, cout, cin_p, cout_p in ((3, 196, 196, 200, 200), (1, 128, 196, 128, 200), (3, 128, 128, 128, 128)):
        w, b = _rnd(cout, cin, k, k, seed=7), _rnd(cout, seed=8)
        wt, bias = engine.pack_conv3x3(w, b, cin_p, cout_p, "cpu")
        cin_k = (cin_p + 63) // 64 * 64
        assert wt.dtype == torch.float16 and tuple(wt.shape) == (cout_p, k * k, cin_k) and tuple(bias.shape) == (cout_p,)
        for tap in range(k * k):
            dy, dx = divmod(tap, k)
            assert torch.equal(wt[:cout, tap, :cin], w[:, :, dy, dx].half())
        assert (wt[cout:] == 0).all() and (wt[:, :, cin:] == 0).all() and (bias[cout:] == 0).all()
        assert torch.equal(bias[:cout], b)


def test_bn_folding_equals_conv_then_eval_batchnorm():
    w = _rnd(16, 8, 3, 3, seed=1)
    sd = {"bn.weight": _rnd(16, seed=2).abs() + 0.5, "bn.bias": _rnd(16, seed=3), "bn.running_mean": _rnd(16, seed=4),
          "bn.running_var": _rnd(16, seed=5).abs() + 0.1}
    x = _rnd(2, 8, 9, 11, seed=6)
    wf, bf = engine._fold_bn(w, sd, "bn")
    want = F.batch_norm(F.conv2d(x, w, None, 1, 1), sd["bn.running_mean"], sd["bn.running_var"], sd["bn.weight"],
                        sd["bn.bias"], False, 0.0, 1e-5)
    got = F.conv2d(x, wf, bf, 1, 1)
    assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


def test_synth_mixed_regime_layout():
    from geoformer_b200 import synth
    a, b = synth.make_pairs(3, 32, 48, "mixed", 5)
    assert torch.equal(a[0], b[0]) and not torch.equal(a[1], b[1])
    assert torch.equal(b[2], torch.roll(a[2], (8, 16), (1, 2)))
