"""torch (CPU) emulations of the `geoformer_b200.ops` front end, one per operator, with the SAME argument contract
(packed weights, row strides, in-place semantics, capacity / count conventions of include/geoformer_b200.h).

TEST INFRASTRUCTURE ONLY.  `install(monkeypatch)` swaps these in for the C-ABI launchers so that the HOST side of the
path - geoformer_b200.engine and GeoFormer._forward: which operator runs when, on which buffer views, with which
weights, masks, counts and skip flags - executes on a machine without a GPU and can be held against the reference
goldens.  The arithmetic of the real kernels is what the `-m gpu` tests check; nothing in the product imports this file.
Emulations compute in fp32 (fp16 only where the contract stores fp16)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPI_RELU, EPI_TANH, EPI_ELU1, EPI_LN = 1, 2, 4, 8
CALLS: dict = {}            # operator name -> number of calls since the last install() (lets tests assert the route taken)


def _count(name):
    CALLS[name] = CALLS.get(name, 0) + 1


# ------------------------------------------------------------------------------------------------ linear layers
def linear(a, w, a2=None, epi=0, act_cols=0, bias=None, rowbias=None, rowbias_group=0, gamma=None, beta=None,
           residual=None, out=None, impl=None, out_f16=False):
    """gf_linear_tf32 / gf_linear_ref / gf_linear_mixed: acc -> +bias -> +rowbias[row / group] -> activation -> LayerNorm
    -> +residual (header, epilogue flags)."""
    _count("linear")
    assert a.is_contiguous() and w.is_contiguous() and a.dtype == w.dtype
    x = a.float() if a2 is None else torch.cat([a.float(), a2.float()], 1)
    assert x.shape[1] == w.shape[1]
    y = x @ w.float().t()
    if bias is not None:
        y = y + bias
    if rowbias is not None:
        y = y + rowbias.repeat_interleave(rowbias_group, 0)
    if epi & EPI_RELU:
        y = F.relu(y)
    if epi & EPI_TANH:
        y = torch.tanh(y)
    if epi & EPI_ELU1:
        y = torch.cat([F.elu(y[:, :act_cols]) + 1, y[:, act_cols:]], 1)
    if epi & EPI_LN:
        y = F.layer_norm(y, (y.shape[1],), gamma, beta, 1e-5)
    if residual is not None:
        y = y + residual
    y = y.half() if out_f16 else y
    if out is not None:
        assert out.shape == y.shape and out.dtype == y.dtype
        out.copy_(y)
        return out
    return y.contiguous()


# ------------------------------------------------------------------------------------------------ backbone
def _act(y, act):
    return F.relu(y) if act == 1 else (F.leaky_relu(y, 0.01) if act == 2 else y)


def conv_ref(x, wt, bias, residual=None, act=0, stride=1):                 # wt [k*k, cin, cout] fp32, x NHWC fp32
    _count("conv_ref")
    k = {1: 1, 9: 3, 49: 7}[wt.shape[0]]
    w = wt.reshape(k, k, wt.shape[1], wt.shape[2]).permute(3, 2, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, stride, k // 2).permute(0, 2, 3, 1)
    return _act(y if residual is None else y + residual, act).contiguous()


def conv(x, wt, bias, residual=None, act=0, stride=1):                      # wt [cout_p, taps, cin_k] fp16, x NHWC fp16
    _count("conv")
    taps, cin_p = wt.shape[1], x.shape[-1]
    k = 3 if taps == 9 else 1
    w = wt[:, :, :cin_p].float().reshape(wt.shape[0], k, k, cin_p).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, bias, stride, k // 2).permute(0, 2, 3, 1)
    return _act(y if residual is None else y + residual.float(), act).half().contiguous()


def stem_conv(img, wperm, bias):                                           # wperm [49, 128]
    _count("stem_conv")
    w = wperm.t().reshape(128, 1, 7, 7)
    return F.relu(F.conv2d(img, w, bias, 2, 3)).permute(0, 2, 3, 1).half().contiguous()


def upsample_add(lateral, src):
    _count("upsample_add")
    up = F.interpolate(src.float().permute(0, 3, 1, 2), size=lateral.shape[1:3], mode="bilinear", align_corners=True)
    return (lateral.float() + up.permute(0, 2, 3, 1)).to(lateral.dtype).contiguous()


def add_posenc(x, pe):
    _count("add_posenc")
    assert pe.shape == x.shape[1:]
    return x + pe[None]


# ------------------------------------------------------------------------------------------------ padding masks
def token_mask(mask, n, tokens):
    if mask.shape[0] != n or mask[0].numel() != tokens:
        raise ValueError(f"padding mask of shape {tuple(mask.shape)} does not cover {n} x {tokens} coarse tokens")
    m = mask.reshape(n, tokens)
    return (m if m.dtype == torch.bool else m != 0).contiguous().view(torch.uint8)


def mask_rows_(buf, mask, col0=0, cols=None):
    """gf_mask_rows: rows with mask == 0 get buf[r, col0:col0+cols] cleared, in place."""
    _count("mask_rows_")
    assert buf.dim() == 2 and buf.is_contiguous() and mask.dtype == torch.uint8 and mask.numel() == buf.shape[0]
    cols = buf.shape[1] - col0 if cols is None else cols
    assert (buf.shape[1] * buf.element_size()) % 4 == 0 and (col0 * buf.element_size()) % 4 == 0 and (cols * buf.element_size()) % 4 == 0
    buf[mask.reshape(-1) == 0, col0:col0 + cols] = 0
    return buf


def mask_fill_sim_(sim, mask0, mask1, fill=-1e9):
    _count("mask_fill_sim_")
    n, l, s = sim.shape
    assert tuple(mask0.shape) == (n, l) and tuple(mask1.shape) == (n, s) and mask0.dtype == mask1.dtype == torch.uint8
    keep = (mask0[:, :, None] != 0) & (mask1[:, None, :] != 0)
    sim[~keep] = fill
    return sim


# ------------------------------------------------------------------------------------------------ linear attention
def _linattn(q, k, v, n, l, s, heads, dim, out_dtype):
    c = heads * dim
    Q = q[:, :c].float().reshape(n, l, heads, dim)          # Q / K arrive mapped through elu+1 (projection epilogue)
    K = k[:, :c].float().reshape(n, s, heads, dim)
    V = v[:, :c].float().reshape(n, s, heads, dim)
    kv = torch.einsum("nshd,nshv->nhdv", K, V / s)
    z = 1.0 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(1)) + 1e-6)
    out = torch.einsum("nlhd,nhdv,nlh->nlhv", Q, kv, z) * s
    return out.reshape(n * l, c).to(out_dtype).contiguous()


def linattn(q, ldq, k, ldk, v, ldv, n, l, s, heads, dim):
    _count("linattn")
    assert q.stride(0) == ldq and k.stride(0) == ldk and v.stride(0) == ldv and q.dtype == k.dtype == v.dtype
    return _linattn(q, k, v, n, l, s, heads, dim, q.dtype)


def linattn_window(q, ldq, k, ldk, v, ldv, n_windows, tokens, heads, dim):
    _count("linattn_window")
    assert q.stride(0) == ldq and k.stride(0) == ldk and v.stride(0) == ldv
    return _linattn(q, k, v, n_windows, tokens, tokens, heads, dim, q.dtype)


def unpack_fine_layer(wpack):
    """Inverse of engine.pack_fine_layer: the 30 operand blocks [30][128][128 B] -> (wqkv [384,128], wm, w1, w2) fp32."""
    assert wpack.dtype == torch.uint8 and tuple(wpack.shape) == (30, 128, 128)
    f32 = lambda b: wpack[b].contiguous().view(torch.float32)          # [128, 32]
    f16 = lambda b: wpack[b].contiguous().view(torch.float16).float()  # [128, 64]
    blk = 0
    wqkv = torch.zeros(384, 128)
    for nc in range(3):
        for kb in range(4):
            wqkv[nc * 128:(nc + 1) * 128, kb * 32:(kb + 1) * 32] = f32(blk); blk += 1
    w1 = torch.zeros(256, 256)
    for nc in range(2):
        for kb in range(4):
            w1[nc * 128:(nc + 1) * 128, kb * 32:(kb + 1) * 32] = f32(blk); blk += 1
    wm = torch.zeros(128, 128)
    for kb in range(2):
        wm[:, kb * 64:(kb + 1) * 64] = f16(blk); blk += 1
    for nc in range(2):
        for kb in range(2):
            w1[nc * 128:(nc + 1) * 128, 128 + kb * 64:128 + (kb + 1) * 64] = f16(blk); blk += 1
    w2 = torch.zeros(128, 256)
    for kb in range(4):
        w2[:, kb * 64:(kb + 1) * 64] = f16(blk); blk += 1
    assert blk == 30
    return wqkv, wm, w1, w2


def fine_layer_fused(x, src, wpack, n1w, n1b, n2w, n2b):
    """gf_fine_layer: one fine-level LoFTR layer (transformer.py:28-60), x / src [m, 25, 128] fp32."""
    _count("fine_layer_fused")
    m, t, c = x.shape
    assert (t, c) == (25, 128) and src.shape == x.shape and x.is_contiguous() and src.is_contiguous()
    wqkv, wm, w1, w2 = unpack_fine_layer(wpack)
    x2, s2 = x.reshape(m * t, c), src.reshape(m * t, c)
    q = F.elu(x2 @ wqkv[:128].t()) + 1
    k = F.elu(s2 @ wqkv[128:256].t()) + 1
    v = s2 @ wqkv[256:].t()
    msg = _linattn(q, k, v, m, t, t, 8, 16, torch.float32)
    msg = F.layer_norm(msg @ wm.t(), (c,), n1w, n1b, 1e-5)
    h = F.relu(torch.cat([x2, msg], 1) @ w1.t())
    y = F.layer_norm(h @ w2.t(), (c,), n2w, n2b, 1e-5)
    return (x2 + y).reshape(m, t, c).contiguous()


# ------------------------------------------------------------------------------------------------ coarse matching
def similarity(f0, f1, temperature, impl=None):
    _count("similarity")
    c = f0.shape[-1]
    return (torch.einsum("nlc,nsc->nls", f0 / c ** 0.5, f1 / c ** 0.5) / temperature).contiguous()


def dual_softmax_(sim):
    """in place sim -> conf; returns (conf, per-row max of conf, per-column max of conf)."""
    _count("dual_softmax_")
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    sim.copy_(conf)
    return sim, sim.max(dim=2)[0], sim.max(dim=1)[0]


def conf_row_col_max(conf):
    _count("conf_row_col_max")
    return conf.max(dim=2)[0], conf.max(dim=1)[0]


def _compact(mj, mc, hw0c, hw1c, scale):
    """gf_compact_coarse: ordered (b, i) compaction of match_j >= 0; per-sample counts as an int32 host tensor."""
    n, l = mj.shape
    b_ids, i_ids = torch.where(mj >= 0)
    j_ids = mj[b_ids, i_ids].long()
    k0 = torch.stack([i_ids % hw0c[1], i_ids // hw0c[1]], 1).float() * float(scale)
    k1 = torch.stack([j_ids % hw1c[1], j_ids // hw1c[1]], 1).float() * float(scale)
    counts = torch.bincount(b_ids, minlength=n).to(torch.int32)
    return dict(b_ids=b_ids, i_ids=i_ids, j_ids=j_ids, m_bids=b_ids, mconf=mc[b_ids, i_ids], mkpts0_c=k0,
                mkpts1_c=k1), counts


def mutual_nearest(conf, crmax, ccmax, thr, border, hw0c, hw1c, scale):
    """gf_mnn_select + gf_compact_coarse: first j with conf > thr, == row max, == column max, inside the border."""
    _count("mutual_nearest")
    n, l, s = conf.shape
    hit = (conf > thr) & (conf == crmax[:, :, None]) & (conf == ccmax[:, None, :])
    if border > 0:
        ok0 = torch.zeros(hw0c, dtype=torch.bool); ok0[border:hw0c[0] - border, border:hw0c[1] - border] = True
        ok1 = torch.zeros(hw1c, dtype=torch.bool); ok1[border:hw1c[0] - border, border:hw1c[1] - border] = True
        hit = hit & ok0.reshape(1, l, 1) & ok1.reshape(1, 1, s)
    any_, first = hit.max(dim=2)
    mj = torch.where(any_, first, torch.full_like(first, -1))
    mc = torch.where(any_, crmax, torch.zeros_like(crmax))
    return _compact(mj, mc, hw0c, hw1c, scale)


def coarse_match_fused(f0, f1, temperature, thr, border, hw0c, hw1c, scale):
    _count("coarse_match_fused")
    conf, crmax, ccmax = dual_softmax_(similarity(f0, f1, temperature))
    CALLS["similarity"] -= 1; CALLS["dual_softmax_"] -= 1
    out = mutual_nearest(conf, crmax, ccmax, thr, border, hw0c, hw1c, scale)
    CALLS["mutual_nearest"] -= 1
    return out


# ------------------------------------------------------------------------------------------------ geo attention
def geo_window_table(hmat, has_h, n, hw_src_c, hw_dst_px, w_dst_c, scale, window):
    """gf_geo_window_table: widx[n, l, window^2] = token of the other image, -1 outside it or when has_h[n] == 0."""
    _count("geo_window_table")
    hs, ws = hw_src_c
    l = hs * ws
    ys, xs = torch.meshgrid(torch.arange(hs), torch.arange(ws), indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(l, dtype=torch.long)], 1).float()
    pts[:, :2] *= scale
    half = window // 2
    r = torch.arange(window).reshape(-1, 1).repeat(1, window)
    offs = torch.stack([r.T, r], -1).float().sub(half).mul(scale).reshape(1, window * window, 2)
    widx = torch.full((n, l, window * window), -1, dtype=torch.int32)
    for b in range(n):
        if not int(has_h[b]):
            continue
        wp = pts @ hmat[b].reshape(3, 3).float().t()
        sc = wp[:, 2:].clone()
        sc[sc == 0] = 1e-6
        k = (wp[:, :2] / sc)[:, None, :] + offs
        oob = (k[..., 0] < 0) | (k[..., 1] < 0) | (k[..., 0] >= hw_dst_px[1]) | (k[..., 1] >= hw_dst_px[0]) | torch.isnan(k).any(-1)
        cell = torch.div(k.clamp_min(0).long(), scale, rounding_mode="floor")
        tok = (cell[..., 1] * w_dst_c + cell[..., 0]).to(torch.int32)
        widx[b] = torch.where(oob, torch.full_like(tok, -1), tok)
    return widx


def geo_self_attention(q, ldq, k, ldk, v, ldv, n, l, heads, dim, anchor_idx, anchor_cnt, max_cnt=0, impl=None):
    """All l tokens of a sample against its anchor tokens (full softmax, scale 1/sqrt(dim)); samples without anchors
    get zeros (the caller restores the layer input with select_rows_)."""
    _count("geo_self_attention")
    c = heads * dim
    assert q.stride(0) == ldq and k.stride(0) == ldk and v.stride(0) == ldv
    out = torch.zeros(n * l, c)
    for b in range(n):
        cnt = int(anchor_cnt[b])
        if cnt == 0:
            continue
        idx = anchor_idx[b, :cnt].long() + b * l
        Q = q[b * l:(b + 1) * l, :c].float().reshape(l, heads, dim)
        K = k[idx, :c].float().reshape(cnt, heads, dim)
        V = v[idx, :c].float().reshape(cnt, heads, dim)
        p = torch.softmax(torch.einsum("lhd,shd->lsh", Q, K) / math.sqrt(dim), dim=1)
        out[b * l:(b + 1) * l] = torch.einsum("lsh,shd->lhd", p, V).reshape(l, c)
    return out


def geo_cross_attention(q, ldq, kp, ldk, vp, ldv, n, l, s, heads, dim, widx):
    """One query token against its window of the other image's projected K / V rows; entries with widx < 0 are masked
    (-1e8 fill), rows without a valid entry give 0 (geo_attention.py:72-100)."""
    _count("geo_cross_attention")
    c = heads * dim
    assert q.stride(0) == ldq and kp.stride(0) == ldk and vp.stride(0) == ldv
    ww = widx.shape[2]
    out = torch.zeros(n * l, c)
    for b in range(n):
        wi = widx[b].long()                                    # [l, ww]
        valid = wi >= 0
        rows = wi.clamp_min(0) + b * s
        Q = q[b * l:(b + 1) * l, :c].float().reshape(l, heads, dim)
        K = kp[rows.reshape(-1), :c].float().reshape(l, ww, heads, dim)
        V = vp[rows.reshape(-1), :c].float().reshape(l, ww, heads, dim)
        sc = torch.einsum("lhd,lwhd->lwh", Q, K) / math.sqrt(dim)
        sc = sc.masked_fill(~valid[:, :, None], -1e8)
        o = torch.einsum("lwh,lwhd->lhd", torch.softmax(sc, dim=1), V).reshape(l, c)
        o[valid.sum(1) == 0] = 0
        out[b * l:(b + 1) * l] = o
    return out.to(q.dtype)


def select_rows_(dst, src, flag, n, l, c):
    _count("select_rows_")
    keep = (flag.reshape(-1)[:n] == 0)
    d, s_ = dst.view(n, l * c), src.reshape(n, l * c)
    d[keep] = s_[keep].to(d.dtype)


# ------------------------------------------------------------------------------------------------ fine level
def fine_gather(fine_nhwc, b_ids, tok_ids, wc, stride, window, out=None):
    """window x window (stride `stride`, zero padding window // 2) patches of the NHWC fine map around coarse tokens."""
    _count("fine_gather")
    nb, hf, wf, c = fine_nhwc.shape
    m = b_ids.shape[0]
    half = window // 2
    pad = F.pad(fine_nhwc.float(), (0, 0, half, half, half, half))
    cy, cx = (tok_ids // wc) * stride, (tok_ids % wc) * stride            # top-left of the window in the padded map
    r = torch.arange(window)
    yy = (cy[:, None, None] + r[None, :, None]).expand(m, window, window)
    xx = (cx[:, None, None] + r[None, None, :]).expand(m, window, window)
    win = pad[b_ids[:, None, None], yy, xx].reshape(m, window * window, c)
    if out is None:
        return win.contiguous()
    out.copy_(win.to(out.dtype))
    return out


def gather_rows(feat, b_ids, tok_ids, out=None):
    _count("gather_rows")
    rows = feat[b_ids, tok_ids].float()
    if out is None:
        return rows.contiguous()
    out.copy_(rows)
    return out


def fine_match(f0, f1, temperature, thr, mkpts0_c, mkpts1_c, b_ids, window, coarse_scale, c2f, fine_scale,
               want_matrix=False):
    """gf_fine_match + gf_compact_fine (fine_matching2.py:52-124): per match the global arg-max cell of the 25 x 25 dual
    softmax (first index on ties), kept if > thr."""
    _count("fine_match")
    m, ww, c = f0.shape
    sim = torch.einsum("mlc,msc->mls", f0 / c ** 0.5, f1 / c ** 0.5) / temperature
    conf = F.softmax(sim, 1) * F.softmax(sim, 2)
    top = conf.reshape(m, ww * ww).argmax(1)
    fconf = conf.reshape(m, ww * ww)[torch.arange(m), top]
    fi, fj = (top // ww), (top % ww)
    sel = fconf > thr
    half = window // 2
    c0 = mkpts0_c / coarse_scale * c2f
    c1 = mkpts1_c / coarse_scale * c2f
    k0 = (torch.stack([fi % window - half, fi // window - half], 1).float() + c0) * fine_scale
    k1 = (torch.stack([fj % window - half, fj // window - half], 1).float() + c1) * fine_scale
    out = dict(mkpts0_f=k0[sel], mkpts1_f=k1[sel], mconf=fconf[sel], m_bids=b_ids[sel])
    return out, (conf if want_matrix else None), (sel.to(torch.int32), fi.to(torch.int32), fj.to(torch.int32), fconf)


def resize_gray_u8(src, dst):
    """gf_resize_gray_u8: cv2.resize (INTER_LINEAR, uint8) + / 255 (the GPU kernel is bit-exact to cv2, test_next_rows)."""
    import cv2
    _count("resize_gray_u8")
    dst.copy_(torch.from_numpy(cv2.resize(src.numpy(), (dst.shape[1], dst.shape[0]))).float().div(255))
    return dst


def ransac_homography(k0, k1, b_ids, counts, n, hw0_c, hw1_c, scale, thr=8.0, hyps=1024, seed=0):
    """gf_ransac_homography's OUTPUT CONTRACT (hm [2,n,9] = H and H^-1 as fp32, has_h [n], inlier [m], ascending anchor
    lists aidx [2,n,cap] + acnt [2,n]; all first-pass matches are anchors when a sample has no homography) with
    OpenCV as the estimator: exercises the host glue of GeoFormer.ransac == 'gpu' (the device estimator itself is not
    bit-identical to OpenCV and is judged geometrically on the GPU, tests/test_gpu_ransac.py)."""
    import cv2
    import numpy as np
    _count("ransac_homography")
    m = int(k0.shape[0])
    l0, l1 = hw0_c[0] * hw0_c[1], hw1_c[0] * hw1_c[1]
    cap = max(l0, l1)
    hm = torch.zeros(2, n, 9)
    has_h = torch.zeros(n, dtype=torch.int32)
    inlier = torch.ones(max(m, 1), dtype=torch.int32)
    aidx = torch.zeros(2, n, cap, dtype=torch.int32)
    acnt = torch.zeros(2, n, dtype=torch.int32)
    offs = np.concatenate([[0], np.cumsum(counts.numpy())])
    for b in range(n):
        a = k0[offs[b]:offs[b + 1]].numpy().astype(np.int64)
        c = k1[offs[b]:offs[b + 1]].numpy().astype(np.int64)
        keep = np.ones(len(a), dtype=bool)
        if len(a) > 8:
            M, mask = cv2.findHomography(a, c, cv2.RANSAC, thr)
            if M is not None:
                has_h[b] = 1
                keep = mask[:, 0] == 1
                hm[0, b] = torch.from_numpy(M.astype(np.float32).reshape(9))
                hm[1, b] = torch.inverse(torch.from_numpy(M)).to(torch.float32).reshape(9)
        inlier[offs[b]:offs[b + 1]] = torch.from_numpy(keep.astype(np.int32))
        for side, (pts, wc) in enumerate(((a[keep], hw0_c[1]), (c[keep], hw1_c[1]))):
            tok = np.unique((pts[:, 1] // scale) * wc + pts[:, 0] // scale)
            acnt[side, b] = len(tok)
            aidx[side, b, :len(tok)] = torch.from_numpy(tok.astype(np.int32))
    return hm, has_h, inlier[:m], aidx, acnt


# ------------------------------------------------------------------------------------------------ plumbing
class _NoStream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass


OPS = ("linear", "conv", "conv_ref", "stem_conv", "upsample_add", "add_posenc", "token_mask", "mask_rows_",
       "mask_fill_sim_", "linattn", "linattn_window", "fine_layer_fused", "similarity", "dual_softmax_", "mutual_nearest",
       "coarse_match_fused", "geo_window_table", "geo_self_attention", "geo_cross_attention", "select_rows_", "fine_gather",
       "gather_rows", "fine_match", "resize_gray_u8", "conf_row_col_max",
       "ransac_homography")


def install(monkeypatch):
    """Route every operator engine.py / full_model.py call through the emulations above (and make the two
    `torch.cuda.current_stream(...).synchronize()` calls of the weight-packing code no-ops)."""
    from geoformer_b200 import ops
    g = globals()
    for name in OPS:
        assert hasattr(ops, name), name
        monkeypatch.setattr(ops, name, g[name])
    monkeypatch.setattr(ops, "upsample_add_ref", upsample_add)
    monkeypatch.setattr(ops, "ensure_init", lambda device: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _NoStream())
    CALLS.clear()
