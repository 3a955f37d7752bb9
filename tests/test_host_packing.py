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


def test_backbone_host_wiring_matches_oracle_on_cpu(monkeypatch):
    """Both backbone host paths (accurate: fp32 [taps][cin][cout] packs; product: tap-major fp16 packs with the 196-wide
    stage zero-padded to 200 channels) reproduce the oracle's resnet_fpn restatement when the kernels are replaced by
    torch emulations of their contracts: fp32 path to round-off, fp16 path within fp16 storage error."""
    from geoformer_b200 import synth
    from oracle import geoformer_oracle as O
    from tests import emu_ops
    emu_ops.install(monkeypatch)        # torch emulations of the operators' contracts (tests/emu_ops.py)
    sd = synth.make_state_dict(7, randomize_norm=True)
    img = torch.cat(synth.make_pairs(1, 64, 96, "shift", 3), 0)
    with torch.no_grad():
        want_c, want_f = O.backbone(sd, img)
    want_c, want_f = want_c.permute(0, 2, 3, 1), want_f.permute(0, 2, 3, 1)
    rel = lambda a, b: ((a.float() - b).abs().max() / b.abs().max()).item()
    c, f = engine.backbone_forward(engine.PackedWeights(sd, torch.device("cpu"), torch.float32), img)
    assert c.shape == want_c.shape and f.shape == want_f.shape
    assert rel(c, want_c) <= 2e-5 and rel(f, want_f) <= 2e-5
    pw = engine.PackedWeights(sd, torch.device("cpu"), torch.float16)
    assert pw.bb_tc["layer2.0.conv1"][0].shape == (200, 9, 128) and pw.bb_tc["layer2.1.conv1"][0].shape == (200, 9, 256)
    c16, f16 = engine.backbone_forward(pw, img)
    assert c16.dtype == torch.float32 and f16.dtype == torch.float16 and f16.shape == want_f.shape
    assert rel(c16, want_c) <= 1e-2 and rel(f16, want_f) <= 1e-2
