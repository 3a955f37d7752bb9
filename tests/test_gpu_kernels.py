"""GPU parity tests, kernel by kernel, through the C ABI (geoformer_b200.ops -> libgeoformer_sm100.so).
Each CUDA kernel is compared with the CPU oracle (oracle/geoformer_oracle.py) on identical seeded inputs.
Tolerances are stated per test; integer / index outputs must be bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import geoformer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from geoformer_b200 import ops as _ops
    assert torch.cuda.is_available(), "GPU tests need a B200"
    _ops.ensure_init(torch.device("cuda:0"))
    return _ops


def dev(t):
    return t.cuda().contiguous()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------------------------------- linear
LIN_CASES = [
    # (M, N, K1, K2, epi, act_cols, bias, rowbias_group, ln, residual)
    (300, 256, 256, 0, 0, 0, False, 0, False, False),
    (1000, 768, 256, 0, 4, 512, False, 0, False, False),      # fused QKV, elu+1 on q,k
    (77, 512, 256, 256, 1, 0, False, 0, False, False),        # cat[x,msg] -> relu
    (640, 512, 256, 256, 2, 0, False, 0, False, False),       # tanh
    (515, 256, 512, 0, 8, 0, False, 0, True, True),           # LN + residual
    (129, 256, 256, 0, 8, 0, False, 0, True, False),          # LN only
    (250, 128, 128, 0, 0, 0, True, 25, False, False),         # fine merge: bias + per-window row bias
    (375, 384, 128, 0, 4, 256, False, 0, False, False),       # fine QKV
    (200, 128, 256, 0, 8, 0, False, 0, True, True),           # fine LN
    (128, 128, 256, 0, 0, 0, True, 0, False, False),          # down_proj
]


def _lin_inputs(case, seed):
    M, N, K1, K2, epi, act_cols, bias, rbg, ln, res = case
    a = rnd(M, K1, seed=seed)
    a2 = rnd(M, K2, seed=seed + 1) if K2 else None
    w = rnd(N, K1 + K2, seed=seed + 2, scale=(K1 + K2) ** -0.5)
    b = rnd(N, seed=seed + 3) if bias else None
    rb = rnd((M + rbg - 1) // rbg, N, seed=seed + 4) if rbg else None
    gamma = 1 + 0.1 * rnd(N, seed=seed + 5) if ln else None
    beta = 0.1 * rnd(N, seed=seed + 6) if ln else None
    r = rnd(M, N, seed=seed + 7) if res else None
    return a, a2, w, b, rb, gamma, beta, r


def _lin_torch(case, a, a2, w, b, rb, gamma, beta, r, dtype=torch.float64):
    M, N, K1, K2, epi, act_cols, bias, rbg, ln, res = case
    x = a.to(dtype) if a2 is None else torch.cat([a, a2], 1).to(dtype)
    y = x @ w.to(dtype).T
    if b is not None:
        y = y + b.to(dtype)
    if rb is not None:
        y = y + rb.to(dtype)[torch.arange(M) // rbg]
    if epi & 1:
        y = F.relu(y)
    if epi & 2:
        y = torch.tanh(y)
    if epi & 4:
        y[:, :act_cols] = F.elu(y[:, :act_cols]) + 1
    if epi & 8:
        y = F.layer_norm(y, (N,), gamma.to(dtype), beta.to(dtype), 1e-5)
    if r is not None:
        y = y + r.to(dtype)
    return y.float()


@pytest.mark.parametrize("case", LIN_CASES)
@pytest.mark.parametrize("impl,tol", [("ref", 2e-5), ("tf32", 8e-3)])
def test_linear(ops, case, impl, tol):
    """tolerance: fp32 FFMA kernel 2e-5; tcgen05 kind::tf32 (10-bit mantissa operands, fp32 accumulate)
    8e-3 max-abs on O(1..4) outputs (K <= 512; operands are truncated, not rounded, to tf32)."""
    M, N, K1, K2, epi, act_cols, bias, rbg, ln, res = case
    a, a2, w, b, rb, gamma, beta, r = _lin_inputs(case, 11)
    want = _lin_torch(case, a, a2, w, b, rb, gamma, beta, r)
    d = lambda t: None if t is None else dev(t)
    got = ops.linear(d(a), d(w), a2=d(a2), epi=epi, act_cols=act_cols, bias=d(b), rowbias=d(rb), rowbias_group=rbg,
                     gamma=d(gamma), beta=d(beta), residual=d(r), impl=impl).cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err <= tol, f"{impl} {case}: max-abs {err}"


def test_linear_tf32_large_multi_tile(ops):
    """More tiles than SMs (persistent scheduler wraps), both accumulator buffers and all smem stages cycle."""
    M, N, K = 128 * 301 + 5, 512, 512
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    got = ops.linear(dev(a), dev(w), impl="tf32")
    ref = ops.linear(dev(a), dev(w), impl="ref")
    torch.cuda.synchronize()
    assert (got - ref).abs().max().item() <= 8e-3


F16_OUT_CASES = [c for c in LIN_CASES if not c[8] and not c[9] and not c[7]]          # lean epilogue only
F16_IN_CASES = [c for c in LIN_CASES if c[2] % 64 == 0 and c[3] % 64 == 0]


@pytest.mark.parametrize("case", F16_OUT_CASES)
def test_linear_tf32_fp16_output(ops, case):
    """gf_linear_mixed, fp32 operands (kind::tf32) -> fp16 result: tf32 tolerance 8e-3 + one fp16 rounding of an
    O(1..8) value (2^-11 relative)."""
    M, N, K1, K2, epi, act_cols, bias, rbg, ln, res = case
    a, a2, w, b, rb, gamma, beta, r = _lin_inputs(case, 21)
    want = _lin_torch(case, a, a2, w, b, rb, gamma, beta, r)
    d = lambda t: None if t is None else dev(t)
    got = ops.linear(d(a), d(w), a2=d(a2), epi=epi, act_cols=act_cols, bias=d(b), impl="tf32", out_f16=True)
    assert got.dtype == torch.float16
    got = got.float().cpu()
    err = (got - want).abs().max().item()
    assert err <= 8e-3 + want.abs().max().item() * 2.0 ** -11, f"{case}: max-abs {err}"


@pytest.mark.parametrize("case", F16_IN_CASES)
def test_linear_fp16_operands(ops, case):
    """gf_linear_mixed with fp16 A / A2 / W (kind::f16, fp32 accumulate and epilogue, fp32 result) against fp64 on the
    SAME fp16-rounded operands: only accumulation-order error remains (1e-4)."""
    M, N, K1, K2, epi, act_cols, bias, rbg, ln, res = case
    a, a2, w, b, rb, gamma, beta, r = _lin_inputs(case, 31)
    h = lambda t: None if t is None else t.half()
    a, a2, w = h(a), h(a2), h(w)
    want = _lin_torch(case, a.float(), None if a2 is None else a2.float(), w.float(), b, rb, gamma, beta, r)
    d = lambda t: None if t is None else dev(t)
    got = ops.linear(d(a), d(w), a2=d(a2), epi=epi, act_cols=act_cols, bias=d(b), rowbias=d(rb), rowbias_group=rbg,
                     gamma=d(gamma), beta=d(beta), residual=d(r), impl="tf32")
    assert got.dtype == torch.float32
    err = (got.cpu() - want).abs().max().item()
    assert err <= 1e-4, f"{case}: max-abs {err}"


def test_linear_fp16_large_multi_tile(ops):
    M, N, K = 128 * 301 + 5, 768, 256
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    got = ops.linear(dev(a), dev(w), impl="tf32", out_f16=True).float()
    ref = ops.linear(dev(a), dev(w), impl="ref")
    torch.cuda.synchronize()
    assert (got - ref).abs().max().item() <= 8e-3 + ref.abs().max().item() * 2.0 ** -11


# ------------------------------------------------------------------------------------------- similarity / conf
def _feat(n, l, c, seed, rms=3.0, offset=1.5):
    # coarse features under random init have rms ~3 and a common offset -> logits ~40-95
    return rnd(n, l, c, seed=seed, scale=rms) + offset


@pytest.mark.parametrize("shape", [(2, 300, 333), (1, 4800, 4800), (1, 128, 256)])
def test_similarity_f16x3_max_abs_1e3(ops, shape):
    """North-star tolerance: max-abs <= 1e-3 on the similarity logits (vs fp64)."""
    n, l, s = shape
    f0, f1 = _feat(n, l, 256, 3), _feat(n, s, 256, 4)
    want = torch.einsum("nlc,nsc->nls", f0.double() / 16, f1.double() / 16) / 0.1
    got = ops.similarity(dev(f0), dev(f1), 0.1, impl="f16x3").cpu().double()
    assert want.abs().max() > 20
    err = (got - want).abs().max().item()
    assert err <= 1e-3, err
    ref = ops.similarity(dev(f0), dev(f1), 0.1, impl="ref").cpu().double()
    assert (ref - want).abs().max().item() <= 1e-3


def test_dual_softmax_conf(ops):
    n, l, s = 2, 300, 333
    sim = rnd(n, l, s, seed=5, scale=3.0) + 50
    want = F.softmax(sim, 1) * F.softmax(sim, 2)
    conf, crmax, ccmax = ops.dual_softmax_(dev(sim))
    conf = conf.cpu()
    assert (conf - want).abs().max().item() <= 1e-6 + 1e-5 * want.max().item()
    assert torch.equal(crmax.cpu(), conf.max(dim=2)[0])
    assert torch.equal(ccmax.cpu(), conf.max(dim=1)[0])


def _mnn_gpu(ops, conf, thr, border, hw0c, hw1c):
    c = dev(conf)
    crmax, ccmax = ops.conf_row_col_max(c)
    m, counts = ops.mutual_nearest(c, crmax, ccmax, thr, border, hw0c, hw1c, 8.0)
    return {k: v.cpu() for k, v in m.items()}, counts


@pytest.mark.parametrize("thr,border", [(0.0, 0), (0.2, 0), (0.01, 2)])
def test_mnn_bit_exact_on_golden_conf(ops, golden_dir, thr, border):
    """Given identical confidence inputs, MNN + threshold + border + compaction are bit-exact (north star)."""
    z = np.load(os.path.join(golden_dir, "small_dense.npz"))
    conf = torch.from_numpy(z["conf"])
    hw = (12, 16)
    want = O.coarse_match(conf, thr, (96, 128), hw, hw, border)
    got, counts = _mnn_gpu(ops, conf, thr, border, hw, hw)
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(got[k], want[k]), k
    assert torch.equal(got["mconf"], want["mconf"])
    assert torch.equal(got["mkpts0_c"], want["mkpts0_c"]) and torch.equal(got["mkpts1_c"], want["mkpts1_c"])
    assert counts.tolist() == [int((want["b_ids"] == b).sum()) for b in range(conf.shape[0])]


def test_mnn_ties_and_rectangular(ops):
    """Exact ties (first j wins, as torch CPU max on bool), all-equal matrix, rectangular L != S, empty result."""
    g = torch.Generator().manual_seed(9)
    conf = torch.rand(3, 6 * 7, 5 * 9, generator=g)
    conf[0, 3, 10] = conf[0, 3, 20] = 2.0            # two row-maxima that are both column maxima -> j = 10
    conf[1] = 0.25                                   # uniform: every row matches j = 0
    conf[2, 5, :] = 0.0
    want = O.coarse_match(conf, 0.1, (48, 56), (6, 7), (5, 9), 0)
    got, _ = _mnn_gpu(ops, conf, 0.1, 0, (6, 7), (5, 9))
    for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_c", "mkpts1_c"):
        assert torch.equal(got[k], want[k]), k
    got, counts = _mnn_gpu(ops, conf, 5.0, 0, (6, 7), (5, 9))
    assert got["b_ids"].numel() == 0 and counts.sum() == 0


def test_mnn_random_full_size(ops):
    """4800x4800 random confidences: sortedness + agreement with the oracle."""
    g = torch.Generator().manual_seed(10)
    conf = torch.rand(1, 4800, 4800, generator=g)
    want = O.coarse_match(conf, 0.5, (480, 640), (60, 80), (60, 80), 0)
    got, _ = _mnn_gpu(ops, conf, 0.5, 0, (60, 80), (60, 80))
    assert torch.equal(got["i_ids"], want["i_ids"]) and torch.equal(got["j_ids"], want["j_ids"])
    assert (got["i_ids"][1:] > got["i_ids"][:-1]).all()


# ------------------------------------------------------------------------------------------- linear attention
def test_linear_attention(ops):
    n, l, s, h, d = 2, 500, 300, 8, 32
    q, k, v = rnd(n, l, h, d, seed=1), rnd(n, s, h, d, seed=2), rnd(n, s, h, d, seed=3)
    want = O.linear_attention(q, k, v).reshape(n * l, h * d)
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    got = ops.linattn(dev(Q.reshape(n * l, h * d)), h * d, dev(K.reshape(n * s, h * d)), h * d,
                      dev(v.reshape(n * s, h * d)), h * d, n, l, s, h, d).cpu()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_linear_attention_window(ops):
    m, t, h, d = 37, 25, 8, 16
    q, k, v = rnd(m, t, h, d, seed=1), rnd(m, t, h, d, seed=2), rnd(m, t, h, d, seed=3)
    want = O.linear_attention(q, k, v).reshape(m * t, h * d)
    Q, K = F.elu(q) + 1, F.elu(k) + 1
    got = ops.linattn_window(dev(Q.reshape(m * t, h * d)), h * d, dev(K.reshape(m * t, h * d)), h * d,
                             dev(v.reshape(m * t, h * d)), h * d, m, t, h, d).cpu()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_linear_attention_fp16_storage(ops):
    """fp16-storage variants (Q/K/V and message fp16, fp32 accumulation; the 8x32 coarse shape runs on mma.sync with
    the KV / Ksum blocks as fp16 B fragments) against fp64 on the same fp16-rounded inputs: fp32-kernel tolerance + one
    fp16 rounding each of KV, Ksum and the output (3 * 2^-11 of the largest magnitude)."""
    n, l, s, h, d = 2, 500, 300, 8, 32
    q, k, v = rnd(n, l, h, d, seed=1), rnd(n, s, h, d, seed=2), rnd(n, s, h, d, seed=3)
    Q, K, V = (F.elu(q) + 1).half(), (F.elu(k) + 1).half(), v.half()
    # oracle applies elu+1 itself: feed it inverse-mapped values is not possible, so restate the two lines here
    Qf, Kf, Vf = Q.double(), K.double(), V.double()
    KV = torch.einsum("nshd,nshv->nhdv", Kf, Vf / s)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Qf, Kf.sum(1)) + 1e-6)
    want = (torch.einsum("nlhd,nhdv,nlh->nlhv", Qf, KV, Z) * s).reshape(n * l, h * d).float()
    got = ops.linattn(dev(Q.reshape(n * l, h * d)), h * d, dev(K.reshape(n * s, h * d)), h * d,
                      dev(V.reshape(n * s, h * d)), h * d, n, l, s, h, d)
    assert got.dtype == torch.float16
    assert (got.float().cpu() - want).abs().max().item() <= (2e-5 + 3 * 2.0 ** -11) * max(1.0, want.abs().max().item())
    # strided Q|K|V buffer, window kernel
    m, t, h, d = 37, 25, 8, 16
    c = h * d
    qkv = torch.cat([F.elu(rnd(m * t, c, seed=4)) + 1, F.elu(rnd(m * t, c, seed=5)) + 1, rnd(m * t, c, seed=6)], 1).half()
    Qf, Kf, Vf = (qkv[:, i * c:(i + 1) * c].double().reshape(m, t, h, d) for i in range(3))
    KV = torch.einsum("nshd,nshv->nhdv", Kf, Vf / t)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Qf, Kf.sum(1)) + 1e-6)
    want = (torch.einsum("nlhd,nhdv,nlh->nlhv", Qf, KV, Z) * t).reshape(m * t, c).float()
    g = dev(qkv)
    got = ops.linattn_window(g, 3 * c, g[:, c:], 3 * c, g[:, 2 * c:], 3 * c, m, t, h, d)
    assert got.dtype == torch.float16
    assert (got.float().cpu() - want).abs().max().item() <= (2e-5 + 2.0 ** -11) * max(1.0, want.abs().max().item())


# ------------------------------------------------------------------------------------------- geo attention
def _homographies():
    eye = np.eye(3)
    tr = np.array([[1, 0, 16.0], [0, 1, 8.0], [0, 0, 1]])
    g = np.random.RandomState(0)
    rand = np.array([[1.02, 0.03, -5.3], [-0.02, 0.97, 7.9], [1e-5, -2e-5, 1.0]]) + g.randn(3, 3) * 1e-6
    return [eye, tr, rand]


def test_geo_window_table(ops):
    """Index table vs the oracle.  Exact for integer-valued warps; for a general homography the fp32
    projective warp may round differently from the CPU BLAS 3x3 bmm in the last ulp, which can move a
    sample across a cell boundary: require >= 99.9% identical entries there."""
    hw_i, hw_c = (96, 128), (12, 16)
    Hs = _homographies()
    n = len(Hs)
    hm = torch.tensor(np.stack(Hs).reshape(n, 9), dtype=torch.float32)
    has = torch.tensor([1, 1, 1], dtype=torch.int32)
    got = ops.geo_window_table(dev(hm), dev(has), n, hw_c, hw_i, hw_c[1], 8, 5).cpu()
    for b, Hm in enumerate(Hs):
        c = O.warp_points(O.grid_keypoints(hw_i[0], hw_i[1], 8), torch.from_numpy(Hm).float())
        win, mask = O.window_table(c, hw_i, 5, 8)
        want = torch.where(mask, O.window_token_index(win, hw_c[1]), torch.full_like(mask, -1, dtype=torch.long))
        same = (got[b].long() == want).float().mean().item()
        assert same == 1.0 if b < 2 else same >= 0.999, (b, same)
    none = ops.geo_window_table(dev(hm), dev(torch.zeros(3, dtype=torch.int32)), n, hw_c, hw_i, hw_c[1], 8, 5).cpu()
    assert (none == -1).all()


@pytest.mark.parametrize("impl,tol", [("ref", 2e-5), ("tf32", 5e-3), ("tf32_mat", 5e-3)])
def test_geo_self_attention(ops, impl, tol):
    """'ref': fp32 flash-style kernel; 'tf32': per-head tcgen05 GEMMs (Q K^T, P V) + masked row softmax —
    tf32 operand rounding on the logits gives ~1e-3 abs error on O(1) outputs."""
    n, l, h, d = 3, 200, 4, 64
    c = h * d
    qkv = rnd(n * l, 3 * c, seed=4)
    cnts = [0, 1, 130]
    g = torch.Generator().manual_seed(1)
    aidx = torch.zeros(n, 130, dtype=torch.int32)
    for b, cnt in enumerate(cnts):
        aidx[b, :cnt] = torch.sort(torch.randperm(l, generator=g)[:cnt])[0].int()
    dq = dev(qkv)
    got = ops.geo_self_attention(dq, 3 * c, dq[:, c:], 3 * c, dq[:, 2 * c:], 3 * c, n, l, h, d, dev(aidx),
                                 dev(torch.tensor(cnts, dtype=torch.int32)), max_cnt=max(cnts), impl=impl).cpu().view(n, l, c)
    x = qkv.view(n, l, 3 * c)
    for b, cnt in enumerate(cnts):
        if cnt == 0:
            assert (got[b] == 0).all()
            continue
        sel = aidx[b, :cnt].long()
        q = x[b, :, :c].view(1, l, h, d)
        k = x[b, sel, c:2 * c].view(1, cnt, h, d)
        v = x[b, sel, 2 * c:].view(1, cnt, h, d)
        want = O.softmax_attention(q, k, v).view(l, c)
        assert (got[b] - want).abs().max().item() <= tol


def test_geo_self_attention_fp16_qkv(ops):
    """Product path of the geo self layers: Q|K|V already stored fp16 (strided buffer), anchors gathered by
    gf_gather_anchor_kv_h16 (incl. the smem-transposed V), queries read in place.  Reference: the oracle's softmax
    attention on the same fp16-rounded values; fp16 P operand inside the kernel -> 5e-3 like the tf32 bar."""
    n, l, h, d = 3, 333, 4, 64
    c = h * d
    qkv = rnd(n * l, 3 * c, seed=4).half()
    cnts = [0, 70, 200]
    g = torch.Generator().manual_seed(1)
    aidx = torch.zeros(n, 256, dtype=torch.int32)
    for b, cnt in enumerate(cnts):
        aidx[b, :cnt] = torch.sort(torch.randperm(l, generator=g)[:cnt])[0].int()
    dq = dev(qkv)
    got = ops.geo_self_attention(dq, 3 * c, dq[:, c:], 3 * c, dq[:, 2 * c:], 3 * c, n, l, h, d, dev(aidx),
                                 dev(torch.tensor(cnts, dtype=torch.int32)), max_cnt=max(cnts), impl="tf32").cpu().view(n, l, c)
    x = qkv.float().view(n, l, 3 * c)
    for b, cnt in enumerate(cnts):
        if cnt == 0:
            assert (got[b] == 0).all()
            continue
        sel = aidx[b, :cnt].long()
        want = O.softmax_attention(x[b, :, :c].view(1, l, h, d), x[b, sel, c:2 * c].view(1, cnt, h, d),
                                   x[b, sel, 2 * c:].view(1, cnt, h, d)).view(l, c)
        assert (got[b] - want).abs().max().item() <= 5e-3


def test_geo_cross_attention(ops):
    n, l, s, h, d = 2, 150, 170, 4, 64
    c = h * d
    q, kp, vp = rnd(n * l, c, seed=1), rnd(n * s, c, seed=2), rnd(n * s, c, seed=3)
    g = torch.Generator().manual_seed(2)
    widx = torch.randint(-1, s, (n, l, 25), generator=g).int()
    widx[0, 5] = -1                                       # fully masked row -> zeros
    widx[1, 7, 1:] = -1                                   # single live key
    got = ops.geo_cross_attention(dev(q), c, dev(kp), c, dev(vp), c, n, l, s, h, d, dev(widx)).cpu().view(n, l, c)
    for b in range(n):
        idx = widx[b].long().clamp(min=0)
        mask = widx[b] >= 0
        kk = kp.view(n, s, c)[b][idx].view(l, 25, h, d)
        vv = vp.view(n, s, c)[b][idx].view(l, 25, h, d)
        want = O.softmax_attention(q.view(n, l, c)[b].view(l, 1, h, d), kk, vv, kv_mask=mask).view(l, c)
        assert (got[b] - want).abs().max().item() <= 2e-5
    assert (got[0, 5] == 0).all()


def test_geo_cross_attention_fp16_storage(ops):
    """fp16-stored Q / K / V (strided Q|K|V buffers) and fp16 message, fp32 arithmetic: the oracle on the same
    fp16-rounded inputs + one fp16 rounding of the O(1) output."""
    n, l, s, h, d = 2, 150, 170, 4, 64
    c = h * d
    qkv0, qkv1 = rnd(n * l, 3 * c, seed=1).half(), rnd(n * s, 3 * c, seed=2).half()
    g = torch.Generator().manual_seed(2)
    widx = torch.randint(-1, s, (n, l, 25), generator=g).int()
    widx[0, 5] = -1
    widx[1, 7, 1:] = -1
    d0, d1 = dev(qkv0), dev(qkv1)
    got = ops.geo_cross_attention(d0, 3 * c, d1[:, c:], 3 * c, d1[:, 2 * c:], 3 * c, n, l, s, h, d, dev(widx))
    assert got.dtype == torch.float16
    got = got.float().cpu().view(n, l, c)
    q, kp, vp = qkv0[:, :c].float(), qkv1[:, c:2 * c].float(), qkv1[:, 2 * c:].float()
    for b in range(n):
        idx = widx[b].long().clamp(min=0)
        mask = widx[b] >= 0
        kk = kp.view(n, s, c)[b][idx].view(l, 25, h, d)
        vv = vp.view(n, s, c)[b][idx].view(l, 25, h, d)
        want = O.softmax_attention(q.view(n, l, c)[b].view(l, 1, h, d), kk, vv, kv_mask=mask).view(l, c)
        assert (got[b] - want).abs().max().item() <= 2e-5 + 2.0 ** -11 * max(1.0, want.abs().max().item())
    assert (got[0, 5] == 0).all()


# ------------------------------------------------------------------------------------------- fine level
def test_fine_gather_exact(ops):
    n, c, hf, wf, wc = 2, 128, 48, 64, 16
    fmap = rnd(n, c, hf, wf, seed=6)
    g = torch.Generator().manual_seed(3)
    m = 77
    b_ids = torch.randint(0, n, (m,), generator=g)
    tok = torch.randint(0, 12 * 16, (m,), generator=g)
    tok[:4] = torch.tensor([0, 15, 176, 191])           # corners: zero padding
    want = O.fine_windows(fmap, b_ids, tok, 5, 4)
    got = ops.fine_gather(dev(fmap.permute(0, 2, 3, 1)), dev(b_ids), dev(tok), wc, 4, 5).cpu()
    assert torch.equal(got, want)


def test_fine_match(ops, golden_dir):
    """Window features from the reference run (golden) -> identical fine cells / coordinates; confidences 1e-5."""
    z = np.load(os.path.join(golden_dir, "small_dense.npz"))
    f0, f1 = torch.from_numpy(z["fine_out0"]), torch.from_numpy(z["fine_out1"])
    m = f0.shape[0]
    k0c, k1c = torch.from_numpy(z["mkpts0_c"][:m]), torch.from_numpy(z["mkpts1_c"][:m])
    b_ids = torch.from_numpy(z["b_ids"][:m])
    conf = O.dual_softmax_conf(f0, f1, 0.1)
    want = O.fine_match(conf, 0.1, k0c, k1c, b_ids, (96, 128), (12, 16), (48, 64), 5)
    out, fmat, raw = ops.fine_match(dev(f0), dev(f1), 0.1, 0.1, dev(k0c), dev(k1c), dev(b_ids), 5, 8.0, 4.0, 2.0, True)
    assert (fmat.cpu() - conf).abs().max().item() <= 2e-4    # peaky conf ~1 on logits of a few hundred: fp32 round-off
    assert torch.equal(out["mkpts0_f"].cpu(), want["mkpts0_f"]) and torch.equal(out["mkpts1_f"].cpu(), want["mkpts1_f"])
    assert torch.equal(out["m_bids"].cpu(), want["m_bids"])
    assert (out["mconf"].cpu() - want["mconf"]).abs().max().item() <= 2e-4


def test_fine_match_threshold_and_random(ops):
    m = 500
    f0, f1 = rnd(m, 25, 128, seed=7, scale=1.0), rnd(m, 25, 128, seed=8, scale=1.0)
    k0c = torch.randint(0, 60, (m, 2), generator=torch.Generator().manual_seed(4)).float() * 8
    k1c = torch.randint(0, 60, (m, 2), generator=torch.Generator().manual_seed(5)).float() * 8
    b_ids = torch.zeros(m, dtype=torch.long)
    conf = O.dual_softmax_conf(f0, f1, 0.1)
    want = O.fine_match(conf, 0.1, k0c, k1c, b_ids, (480, 640), (60, 80), (240, 320), 5)
    out, _, raw = ops.fine_match(dev(f0), dev(f1), 0.1, 0.1, dev(k0c), dev(k1c), dev(b_ids), 5, 8.0, 4.0, 2.0)
    assert 0 < want["mkpts0_f"].shape[0] < m              # the threshold bites
    assert torch.equal(out["mkpts0_f"].cpu(), want["mkpts0_f"]) and torch.equal(out["mkpts1_f"].cpu(), want["mkpts1_f"])
    assert (out["mkpts0_f"].cpu() % 2 == 0).all()


# ------------------------------------------------------------------------------------------- backbone conv
@pytest.mark.parametrize("cin,cout,cin_p,cout_p,res,act", [
    (128, 128, 128, 128, False, 1), (128, 128, 128, 128, True, 1), (196, 196, 200, 200, True, 1),
    (256, 196, 256, 200, False, 0), (196, 128, 200, 128, False, 0), (256, 256, 256, 256, False, 2),
])
def test_conv3x3_tcgen05(ops, cin, cout, cin_p, cout_p, res, act):
    """tcgen05 implicit-GEMM 3x3 conv (NHWC fp16, zero-padded channels, partial 8x16 tiles at the borders) vs
    F.conv2d in fp32 on the same fp16-rounded operands.  Tolerance: fp16 output rounding (2^-11) on top of
    fp32 accumulation -> 1.5e-3 of the output range."""
    from geoformer_b200.engine import pack_conv3x3
    b, h, w = 2, 20, 40
    x = rnd(b, cin, h, w, seed=1).half()
    wgt = (rnd(cout, cin, 3, 3, seed=2) * (cin * 9) ** -0.5).half()
    bias = rnd(cout, seed=3) * 0.1
    r = rnd(b, cout, h, w, seed=4).half() if res else None
    want = F.conv2d(x.float(), wgt.float(), bias, 1, 1)
    if res:
        want = want + r.float()
    want = F.relu(want) if act == 1 else (F.leaky_relu(want, 0.01) if act == 2 else want)
    xp = torch.zeros(b, h, w, cin_p, dtype=torch.float16); xp[..., :cin] = x.permute(0, 2, 3, 1)
    rp = None
    if res:
        rp = torch.zeros(b, h, w, cout_p, dtype=torch.float16); rp[..., :cout] = r.permute(0, 2, 3, 1)
    wt, bp = pack_conv3x3(wgt.float(), bias, cin_p, cout_p, "cuda")
    y = ops.conv3x3(dev(xp), wt, bp, None if rp is None else dev(rp), act).cpu().float()
    got = y[..., :cout].permute(0, 3, 1, 2)
    assert (y[..., cout:] == 0).all()                       # padded channels stay exactly zero
    err = (got - want).abs().max().item()
    assert err <= 1.5e-3 * want.abs().max().item(), err


@pytest.mark.parametrize("h,w,cin,cin_p,res", [(41, 40, 128, 128, True), (48, 17, 196, 200, False), (33, 64, 128, 128, False)])
def test_conv3x3_two_pixel_tiles_per_weight_box(ops, h, w, cin, cin_p, res):
    """N = 128 layers on images of >= 32 rows take the MT = 2 variant (two stacked 8 x 16 pixel tiles per weight box,
    two accumulators); odd heights leave the second tile partially or completely outside the image."""
    from geoformer_b200.engine import pack_conv3x3
    b, cout = 2, 128
    x = rnd(b, cin, h, w, seed=1).half()
    wgt = (rnd(cout, cin, 3, 3, seed=2) * (cin * 9) ** -0.5).half()
    bias = rnd(cout, seed=3) * 0.1
    r = rnd(b, cout, h, w, seed=4).half() if res else None
    want = F.conv2d(x.float(), wgt.float(), bias, 1, 1)
    want = F.relu(want + r.float()) if res else F.relu(want)
    xp = torch.zeros(b, h, w, cin_p, dtype=torch.float16); xp[..., :cin] = x.permute(0, 2, 3, 1)
    rp = dev(r.permute(0, 2, 3, 1)) if res else None
    wt, bp = pack_conv3x3(wgt.float(), bias, cin_p, cout, "cuda")
    got = ops.conv3x3(dev(xp), wt, bp, rp, 1).cpu().float().permute(0, 3, 1, 2)
    err = (got - want).abs().max().item()
    assert err <= 1.5e-3 * want.abs().max().item(), err


@pytest.mark.parametrize("cin,cout,cin_p,cout_p,ksize,stride,hw,act", [
    (128, 196, 128, 200, 3, 2, (20, 40), 1),      # layer2.0.conv1: 3x3 / stride 2 + ReLU
    (128, 196, 128, 200, 1, 2, (20, 40), 0),      # layer2.0.downsample: 1x1 / stride 2
    (196, 256, 200, 256, 3, 2, (21, 37), 1),      # odd sizes: (h-1)//2+1 outputs, partial tiles
    (196, 256, 200, 256, 1, 2, (21, 37), 0),
    (256, 256, 256, 256, 1, 1, (15, 20), 0),      # layer3_outconv (1x1 lateral, no bias in the reference)
    (128, 196, 128, 200, 1, 1, (24, 33), 0),      # layer1_outconv
])
def test_conv_strided_and_1x1_tcgen05(ops, cin, cout, cin_p, cout_p, ksize, stride, hw, act):
    """The generalised implicit-GEMM conv (TMA element stride 2 on W/H for stride-2 layers; a single tap for 1x1)
    vs F.conv2d in fp32 on the same fp16-rounded operands; same tolerance as the 3x3 / stride-1 test."""
    from geoformer_b200.engine import pack_conv3x3
    b, (h, w) = 2, hw
    x = rnd(b, cin, h, w, seed=1).half()
    wgt = (rnd(cout, cin, ksize, ksize, seed=2) * (cin * ksize * ksize) ** -0.5).half()
    bias = rnd(cout, seed=3) * 0.1
    want = F.conv2d(x.float(), wgt.float(), bias, stride, ksize // 2)
    want = F.relu(want) if act == 1 else want
    xp = torch.zeros(b, h, w, cin_p, dtype=torch.float16); xp[..., :cin] = x.permute(0, 2, 3, 1)
    wt, bp = pack_conv3x3(wgt.float(), bias, cin_p, cout_p, "cuda")
    y = ops.conv(dev(xp), wt, bp, None, act, stride).cpu().float()
    assert tuple(y.shape) == (b, want.shape[2], want.shape[3], cout_p)
    got = y[..., :cout].permute(0, 3, 1, 2)
    assert (y[..., cout:] == 0).all()
    err = (got - want).abs().max().item()
    assert err <= 1.5e-3 * want.abs().max().item(), err


# ------------------------------------------------------------------------------------------- fused coarse matching
def _match_lists(d):
    return torch.stack([d["b_ids"].cpu(), d["i_ids"].cpu(), d["j_ids"].cpu()], 1)


@pytest.mark.parametrize("shape,thr", [((2, 192, 192), 0.0), ((2, 300, 333), 0.0), ((1, 4800, 4800), 0.0), ((2, 192, 192), 0.2), ((2, 300, 333), 1e-4)])
def test_fused_coarse_matching_equals_materialised(ops, golden_dir, shape, thr):
    """The non-materialising two-pass tcgen05 path must select the same matches as the materialised kernels
    (same split-fp16 logits) and as the CPU oracle; confidences within 1e-4 relative."""
    n, l, s = shape
    if l == 192:
        z = np.load(os.path.join(golden_dir, "small_dense.npz"))
        f0, f1 = torch.from_numpy(z["geo0"]), torch.from_numpy(z["geo1"])
        hw0, hw1 = (12, 16), (12, 16)
    else:
        f0, f1 = _feat(n, l, 256, 31, rms=1.0, offset=0.3), _feat(n, s, 256, 32, rms=1.0, offset=0.3)
        hw0, hw1 = ((60, 80), (60, 80)) if l == 4800 else ((15, 20), (9, 37))
    fused, counts = ops.coarse_match_fused(dev(f0), dev(f1), 0.1, thr, 0, hw0, hw1, 8.0)
    sim = ops.similarity(dev(f0), dev(f1), 0.1)
    conf, crmax, ccmax = ops.dual_softmax_(sim)
    mat, counts2 = ops.mutual_nearest(conf, crmax, ccmax, thr, 0, hw0, hw1, 8.0)
    a, b = _match_lists(fused), _match_lists(mat)
    assert a.shape[0] > 0 or thr > 0
    assert torch.equal(a, b)
    assert counts.tolist() == counts2.tolist()
    if a.shape[0]:
        rel = ((fused["mconf"] - mat["mconf"]).abs() / mat["mconf"].abs().clamp(min=1e-30)).max().item()
        assert rel <= 1e-4, rel
    if l <= 333:
        want = O.coarse_match(O.dual_softmax_conf(f0, f1, 0.1), thr, (hw0[0] * 8, hw0[1] * 8), hw0, hw1, 0)
        assert torch.equal(a, torch.stack([want["b_ids"], want["i_ids"], want["j_ids"]], 1))


@pytest.mark.parametrize("border,thr", [(0, 0.0), (2, 0.0), (2, 1e-4), (1, 0.0)])
def test_fused_coarse_matching_exact_ties_and_border(ops, border, thr):
    """Reference tie order in the PRODUCT matcher (coarse_matching.py:176-188: first j with conf == row max AND conf ==
    column max AND inside the border).  Duplicated rows of f1 give exactly equal logits, hence exactly equal confidences,
    in two columns; the earlier copy sits in the border strip, so with border > 0 the reference takes the LATER copy while
    the smallest-j candidate of pass 1 is rejected -> the exact re-scan (sim_fused_kernel<2>) must recover it.
    Duplicated rows of f0 (two rows sharing a column maximum) are in the mix too.  Bit-exact (b, i, j) vs the CPU
    oracle and vs the materialised kernels; L != S; batch 2."""
    n, hw0, hw1 = 2, (12, 16), (11, 19)
    l, s = hw0[0] * hw0[1], hw1[0] * hw1[1]
    f0, f1 = _feat(n, l, 256, 41, rms=1.0, offset=0.3), _feat(n, s, 256, 42, rms=1.0, offset=0.3)
    w0, w1 = hw0[1], hw1[1]
    pairs = []                                                  # (i, j_border, j_inside): f0[i] looks like both copies
    for k, (i, ja, jb) in enumerate([(3 * w0 + 4, 0 * w1 + 5, 4 * w1 + 7), (5 * w0 + 9, 3 * w1 + 0, 6 * w1 + 11),
                                     (7 * w0 + 6, 1 * w1 + 18, 8 * w1 + 3), (8 * w0 + 8, 5 * w1 + 5, 9 * w1 + 9)]):
        for b in range(n):
            f1[b, jb] = f1[b, ja]                               # exact duplicate -> identical columns of the logit matrix
            f0[b, i] = 3.0 * f1[b, ja]                          # strongly aligned: (i, ja) and (i, jb) are mutual maxima
        pairs.append((i, ja, jb))
    f0[:, 2 * w0 + 3] = f0[:, 2 * w0 + 2]                       # duplicated query rows
    f0[1, 6 * w0 + 1] = f0[1, 3 * w0 + 4]                       # a second row tied on the same (duplicated) columns
    conf = O.dual_softmax_conf(f0, f1, 0.1)
    want = O.coarse_match(conf, thr, (hw0[0] * 8, hw0[1] * 8), hw0, hw1, border)
    want_l = torch.stack([want["b_ids"], want["i_ids"], want["j_ids"]], 1)
    for i, ja, jb in pairs[:3]:                                 # the constructed rows really tie and really match
        assert conf[0, i, ja] == conf[0, i, jb] == conf[0, i].max()
        hit = want_l[(want_l[:, 0] == 0) & (want_l[:, 1] == i)]
        assert hit.shape[0] == 1 and int(hit[0, 2]) == (ja if border == 0 else jb)      # ja lies in the border strip
    fused, counts = ops.coarse_match_fused(dev(f0), dev(f1), 0.1, thr, border, hw0, hw1, 8.0)
    got = _match_lists(fused)
    assert torch.equal(got, want_l), (got.shape, want_l.shape)
    assert torch.equal(fused["mkpts1_c"].cpu(), want["mkpts1_c"]) and torch.equal(fused["mkpts0_c"].cpu(), want["mkpts0_c"])
    sim = ops.similarity(dev(f0), dev(f1), 0.1)
    cf, crmax, ccmax = ops.dual_softmax_(sim)
    mat, counts2 = ops.mutual_nearest(cf, crmax, ccmax, thr, border, hw0, hw1, 8.0)
    assert torch.equal(got, _match_lists(mat)) and counts.tolist() == counts2.tolist()
    rel = ((fused["mconf"].cpu() - want["mconf"]).abs() / want["mconf"].abs().clamp(min=1e-30)).max().item()
    assert rel <= 2e-3, rel


def test_stem_conv7x7(ops):
    """Stem 7x7/s2 conv + folded BN + ReLU (FFMA kernel, fp32 image -> NHWC fp16) vs F.conv2d; odd sizes hit the
    partial-tile and zero-padding paths.  Tolerance = fp16 output rounding."""
    b, h, w = 2, 70, 100
    img = torch.rand(b, 1, h, w, generator=torch.Generator().manual_seed(1))
    wgt = rnd(128, 1, 7, 7, seed=2) * 0.2
    bias = rnd(128, seed=3) * 0.1
    want = F.relu(F.conv2d(img, wgt, bias, 2, 3))
    wt7 = wgt.reshape(128, 49).t().contiguous()
    got = ops.stem_conv(dev(img), dev(wt7), dev(bias)).cpu().float().permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-3 * want.abs().max().item()


def test_upsample_add(ops):
    """FPN top-down merge: lateral + bilinear x2 (align_corners=True), NHWC fp16."""
    b, hs, ws, c = 2, 15, 20, 200
    lat = rnd(b, c, 2 * hs, 2 * ws, seed=1).half()
    src = rnd(b, c, hs, ws, seed=2).half()
    want = lat.float() + F.interpolate(src.float(), size=(2 * hs, 2 * ws), mode="bilinear", align_corners=True)
    got = ops.upsample_add(dev(lat.permute(0, 2, 3, 1)), dev(src.permute(0, 2, 3, 1))).cpu().float().permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() <= 1.5e-3 * want.abs().max().item()


# ------------------------------------------------------------------------------------------- fused fine layer
@pytest.mark.parametrize("windows", [3, 37, 1003])
@pytest.mark.parametrize("cross", [False, True])
def test_fine_layer_fused(ops, windows, cross):
    """gf_fine_layer (whole LoFTR layer of the fine level in one tcgen05 kernel) against the per-op fp32 FFMA kernels
    of the same library.  Operands are tf32 / fp16 (10-bit mantissa) with fp32 accumulation: 8e-3 max-abs on the
    LayerNorm'd O(1) output (same bar as gf_linear_tf32), partial last tile and ragged window counts included."""
    from geoformer_b200 import engine
    g = lambda *sh, seed, scale=1.0: rnd(*sh, seed=seed, scale=scale)
    wq, wk, wv, wm = (g(128, 128, seed=s, scale=128 ** -0.5) for s in (1, 2, 3, 4))
    w1, w2 = g(256, 256, seed=5, scale=256 ** -0.5), g(128, 256, seed=6, scale=256 ** -0.5)
    lw = dict(wq=dev(wq), wkv=dev(torch.cat([wk, wv], 0)), wqkv=dev(torch.cat([wq, wk, wv], 0)), wm=dev(wm), w1=dev(w1), w2=dev(w2),
              wm16=dev(wm.half()), w2_16=dev(w2.half()),
              n1w=dev(1 + 0.1 * g(128, seed=7)), n1b=dev(0.1 * g(128, seed=8)),
              n2w=dev(1 + 0.1 * g(128, seed=9)), n2b=dev(0.1 * g(128, seed=10)))
    lw["wpack"] = engine.pack_fine_layer(wq, wk, wv, wm, w1, w2, "cuda:0")
    x = dev(g(windows, 25, 128, seed=11))
    src = dev(g(windows, 25, 128, seed=12)) if cross else x
    try:
        ops.set_precision(linear="ref")
        engine.FUSED_FINE_LAYER = False
        want = engine.fine_layer(lw, x, src, 8)
    finally:
        ops.set_precision(linear="tf32")
        engine.FUSED_FINE_LAYER = True
    got = engine.fine_layer(lw, x, src, 8)
    torch.cuda.synchronize()
    assert got.shape == want.shape and torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err <= 8e-3, f"windows={windows} cross={cross}: max-abs {err}"
