"""CPU oracle for the GeoFormer matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``geoformer_b200/`` may import this
module; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the *checker*
(and as the timed CPU arm), never as part of the product path.

What it is: a functional (no ``nn.Module``) fp32 restatement of the reference
algorithm over a flat parameter dict that uses the reference's checkpoint key
names.  Every function cites the reference ``file:line`` it follows (paths are
relative to the reference repository root).

Parity pinning: the reference ships no tests / golden vectors,
so the oracle is pinned against *outputs of the reference itself* run in the
build container: ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (with arithmetic-free stub modules for the
absent ``yacs``/``kornia``/...), runs it on the deterministic synthetic weights
of ``geoformer_b200.synth`` and commits the resulting vectors under
``tests/golden/``.  ``tests/test_oracle_golden.py`` replays them (bit-exact on
the integer outputs, <=1e-5 on floats).

Third-party arithmetic on the path that is not under the reference tree:
torch/ATen (reference pins torch==1.8.1; here 2.11) and OpenCV
``cv2.findHomography`` (reference pins 4.6.0.66; here 4.13.0).  Both oracle and
product call the *same* cv2 in-process, so RANSAC parity holds by construction.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# Resolved inference configuration (model/geo_config.py:9-19 and
# model/loftr_src/loftr/utils/cvpr_ds_config.py:10-50).
DEFAULT_CFG = dict(
    coarse_thr=0.2, fine_thr=0.1, coarse_temperature=0.1, fine_temperature=0.1,
    border_rm=0,            # hard-overridden: coarse_matching.py:31-32
    coarse_nhead=8, fine_nhead=8, geo_nhead=4,
    coarse_layers=("self", "cross") * 4, fine_layers=("self", "cross"),
    geo_layers=("self", "cross") * 2,
    window=5, fine_window=5, ransac_thr=8.0, min_ransac_pts=8,
)


# ----------------------------------------------------------------------------
# backbone: model/loftr_src/loftr/backbone/resnet_fpn.py:15-40, 43-118
# ----------------------------------------------------------------------------
def _bn(P: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"],
                        P[pre + ".weight"], P[pre + ".bias"], False, 0.0, 1e-5)


def _basic_block(P: Params, pre: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    # resnet_fpn.py:32-40
    y = F.conv2d(x, P[pre + ".conv1.weight"], None, stride, 1)
    y = F.relu(_bn(P, pre + ".bn1", y))
    y = _bn(P, pre + ".bn2", F.conv2d(y, P[pre + ".conv2.weight"], None, 1, 1))
    if stride != 1:
        x = _bn(P, pre + ".downsample.1",
                F.conv2d(x, P[pre + ".downsample.0.weight"], None, stride, 0))
    return F.relu(x + y)


def _fpn_head(P: Params, pre: str, x: torch.Tensor) -> torch.Tensor:
    # nn.Sequential(conv3x3, BN, LeakyReLU(0.01), conv3x3): resnet_fpn.py:70-82
    x = F.conv2d(x, P[pre + ".0.weight"], None, 1, 1)
    x = F.leaky_relu(_bn(P, pre + ".1", x), 0.01)
    return F.conv2d(x, P[pre + ".3.weight"], None, 1, 1)


def backbone(P: Params, img: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B,1,H,W] -> coarse [B,256,H/8,W/8], fine [B,128,H/2,W/2] (resnet_fpn.py:100-118)."""
    b = "backbone."
    x0 = F.relu(_bn(P, b + "bn1", F.conv2d(img, P[b + "conv1.weight"], None, 2, 3)))
    x1 = _basic_block(P, b + "layer1.1", _basic_block(P, b + "layer1.0", x0, 1), 1)
    x2 = _basic_block(P, b + "layer2.1", _basic_block(P, b + "layer2.0", x1, 2), 1)
    x3 = _basic_block(P, b + "layer3.1", _basic_block(P, b + "layer3.0", x2, 2), 1)
    x3o = F.conv2d(x3, P[b + "layer3_outconv.weight"])
    x2o = F.conv2d(x2, P[b + "layer2_outconv.weight"])
    up3 = F.interpolate(x3o, size=x2o.shape[2:], mode="bilinear", align_corners=True)
    x2o = _fpn_head(P, b + "layer2_outconv2", x2o + up3)
    x1o = F.conv2d(x1, P[b + "layer1_outconv.weight"])
    up2 = F.interpolate(x2o, size=x1o.shape[2:], mode="bilinear", align_corners=True)
    x1o = _fpn_head(P, b + "layer1_outconv2", x1o + up2)
    return x3o, x1o


# ----------------------------------------------------------------------------
# position encoding: model/loftr_src/loftr/utils/position_encoding.py:11-42
# ----------------------------------------------------------------------------
def position_encoding(d_model: int, h: int, w: int) -> torch.Tensor:
    """[d_model,h,w] table.  Bug-compatible branch (temp_bug_fix=False, line 28):
    ``-ln(1e4) / d_model // 2`` parses as ``(-ln(1e4)/d_model) // 2 = -1.0`` so
    div_term[k] = exp(-2k), k = 0..d_model/4-1."""
    ypos = torch.ones(h, w).cumsum(0).float().unsqueeze(0)
    xpos = torch.ones(h, w).cumsum(1).float().unsqueeze(0)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float()
                    * (-math.log(10000.0) / d_model // 2))[:, None, None]
    pe = torch.zeros(d_model, h, w)
    pe[0::4] = torch.sin(xpos * div)
    pe[1::4] = torch.cos(xpos * div)
    pe[2::4] = torch.sin(ypos * div)
    pe[3::4] = torch.cos(ypos * div)
    return pe


def add_pe_flatten(fmap: torch.Tensor) -> torch.Tensor:
    """[N,C,h,w] -> [N,h*w,C] with PE added (full_model.py:69-77, geo_module.py:28-29)."""
    n, c, h, w = fmap.shape
    x = fmap + position_encoding(c, h, w).unsqueeze(0)
    return x.permute(0, 2, 3, 1).reshape(n, h * w, c)


# ----------------------------------------------------------------------------
# LoFTR encoder layer + linear attention
# model/loftr_src/loftr/loftr_module/transformer.py:37-60, linear_attention.py:21-51
# ----------------------------------------------------------------------------
def linear_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, eps: float = 1e-6,
                     q_mask: Optional[torch.Tensor] = None, kv_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [N,L,H,D], k/v [N,S,H,D] -> [N,L,H,D] (linear_attention.py:33-51).  Optional padding masks q_mask [N,L],
    kv_mask [N,S] zero the feature-mapped Q / K and the values of padded tokens (lines 37-43); the normaliser
    v_length stays the full S."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        v = v * kv_mask[:, :, None, None]
    s_len = v.size(1)
    v = v / s_len
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * s_len).contiguous()


def softmax_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor,
                      kv_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Full attention of the geo transformer (model/geo_transformer/geo_attention.py:72-101).
    q [N,L,H,D], k/v [N,S,H,D], kv_mask [N,S] bool. Masked logits are *filled* with
    -1e8 (not -inf); rows whose mask is all-False are zeroed."""
    qk = torch.einsum("nlhd,nshd->nlsh", q, k)
    if kv_mask is not None:
        qk.masked_fill_(~kv_mask[:, None, :, None], float(-1e8))
    a = torch.softmax(qk * (1.0 / q.size(3) ** 0.5), dim=2)
    out = torch.einsum("nlsh,nshd->nlhd", a, v)
    if kv_mask is not None:
        dead = kv_mask.sum(-1) == 0
        out[dead] = 0 * out[dead]
    return out.contiguous()


def encoder_layer(P: Params, pre: str, x: torch.Tensor, src: torch.Tensor, nhead: int,
                  attention: str = "linear", act: str = "relu",
                  kv_mask: Optional[torch.Tensor] = None, q_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One LoFTR-style layer.  ``act``: 'relu' for LoFTR (transformer.py:27-31),
    'tanh' for the geo transformer (model/geo_transformer/transformer.py:29-33)."""
    n, _, c = x.shape
    d = c // nhead
    q = F.linear(x, P[pre + ".q_proj.weight"]).view(n, -1, nhead, d)
    k = F.linear(src, P[pre + ".k_proj.weight"]).view(n, -1, nhead, d)
    v = F.linear(src, P[pre + ".v_proj.weight"]).view(n, -1, nhead, d)
    if attention == "linear":
        m = linear_attention(q, k, v, q_mask=q_mask, kv_mask=kv_mask)
    else:
        m = softmax_attention(q, k, v, kv_mask)
    m = F.linear(m.view(n, -1, c), P[pre + ".merge.weight"])
    m = F.layer_norm(m, (c,), P[pre + ".norm1.weight"], P[pre + ".norm1.bias"], 1e-5)
    m = F.linear(torch.cat([x, m], dim=2), P[pre + ".mlp.0.weight"])
    m = F.relu(m) if act == "relu" else torch.tanh(m)
    m = F.linear(m, P[pre + ".mlp.2.weight"])
    m = F.layer_norm(m, (c,), P[pre + ".norm2.weight"], P[pre + ".norm2.bias"], 1e-5)
    return x + m


def local_feature_transformer(P: Params, pre: str, f0: torch.Tensor, f1: torch.Tensor,
                              layer_names, nhead: int, mask0: Optional[torch.Tensor] = None,
                              mask1: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """loftr_module/transformer.py:82-104.  NB: the cross update of f1 sees the
    *already updated* f0 (lines 99-100).  mask0/mask1 [N,L]/[N,S]: optional padding masks (training collation)."""
    for i, name in enumerate(layer_names):
        lp = f"{pre}.layers.{i}"
        if name == "self":
            f0 = encoder_layer(P, lp, f0, f0, nhead, q_mask=mask0, kv_mask=mask0)
            f1 = encoder_layer(P, lp, f1, f1, nhead, q_mask=mask1, kv_mask=mask1)
        else:
            f0 = encoder_layer(P, lp, f0, f1, nhead, q_mask=mask0, kv_mask=mask1)
            f1 = encoder_layer(P, lp, f1, f0, nhead, q_mask=mask1, kv_mask=mask0)
    return f0, f1


# ----------------------------------------------------------------------------
# coarse matching: model/loftr_src/loftr/utils/coarse_matching.py:90-212
# ----------------------------------------------------------------------------
def dual_softmax_conf(f0: torch.Tensor, f1: torch.Tensor, temperature: float, mask0: Optional[torch.Tensor] = None,
                      mask1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conf [N,L,S] (coarse_matching.py:110-125; fine_matching2.py:52-60).  With padding masks the logits of padded
    rows / columns are filled with -1e9 before the two softmaxes (coarse_matching.py:120-124, INF = 1e9)."""
    c = f0.shape[-1]
    a, b = f0 / c ** 0.5, f1 / c ** 0.5
    sim = torch.einsum("nlc,nsc->nls", a, b) / temperature
    if mask0 is not None:
        sim = sim.masked_fill(~(mask0[..., None] * mask1[:, None]).bool(), -1e9)
    return F.softmax(sim, 1) * F.softmax(sim, 2)


def similarity(f0: torch.Tensor, f1: torch.Tensor, temperature: float) -> torch.Tensor:
    c = f0.shape[-1]
    return torch.einsum("nlc,nsc->nls", f0 / c ** 0.5, f1 / c ** 0.5) / temperature


def _mask_border(mask5: torch.Tensor, b: int) -> None:
    # coarse_matching.py:71-88 (mask is [N,h0,w0,h1,w1]); no-op for b<=0
    if b <= 0:
        return
    mask5[:, :b] = False
    mask5[:, :, :b] = False
    mask5[:, :, :, :b] = False
    mask5[:, :, :, :, :b] = False
    mask5[:, -b:] = False
    mask5[:, :, -b:] = False
    mask5[:, :, :, -b:] = False
    mask5[:, :, :, :, -b:] = False


def mutual_nearest(conf: torch.Tensor, thr: float, hw0c: Tuple[int, int], hw1c: Tuple[int, int],
                   border_rm: int = 0):
    """(b_ids, i_ids, j_ids, mconf) in row-major (b,i) order (coarse_matching.py:161-190).
    First ``True`` per row wins on exact ties (torch CPU ``max`` on bool)."""
    n = conf.shape[0]
    mask = conf > thr
    m5 = mask.reshape(n, hw0c[0], hw0c[1], hw1c[0], hw1c[1])
    _mask_border(m5, border_rm)
    mask = m5.reshape(n, hw0c[0] * hw0c[1], hw1c[0] * hw1c[1])
    mask = mask * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    row_any, first_j = mask.max(dim=2)
    b_ids, i_ids = torch.where(row_any)
    j_ids = first_j[b_ids, i_ids]
    return b_ids, i_ids, j_ids, conf[b_ids, i_ids, j_ids]


def coarse_match(conf: torch.Tensor, thr: float, hw0_i, hw0_c, hw1_c, border_rm: int = 0) -> Dict[str, torch.Tensor]:
    """coarse_matching.py:132-212 (inference branch: no 'dataset_name', no scale0/1)."""
    b_ids, i_ids, j_ids, mconf = mutual_nearest(conf, thr, tuple(hw0_c), tuple(hw1_c), border_rm)
    scale = torch.tensor(hw0_i[0]) / torch.tensor(hw0_c[0])       # 0-dim fp32 (int/int -> true div)
    w0, w1 = torch.tensor(hw0_c[1]), torch.tensor(hw1_c[1])
    k0 = torch.stack([i_ids % w0, i_ids // w0], dim=1) * scale
    k1 = torch.stack([j_ids % w1, j_ids // w1], dim=1) * scale
    return dict(b_ids=b_ids, i_ids=i_ids, j_ids=j_ids, m_bids=b_ids, mkpts0_c=k0, mkpts1_c=k1, mconf=mconf)


# ----------------------------------------------------------------------------
# geo module: model/geo_module.py:23-116, utils/common_utils.py:65-91,137-144,166-181,
#             utils/homography.py:86-105
# ----------------------------------------------------------------------------
def grid_keypoints(h: int, w: int, scale: int = 8) -> torch.Tensor:
    """[L,2] int64 (x,y) = (col*scale, row*scale), row-major (common_utils.py:137-144)."""
    ys, xs = torch.meshgrid(torch.arange(h // scale), torch.arange(w // scale), indexing="ij")
    return torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).long() * scale


def warp_points(pts_xy: torch.Tensor, Hm: torch.Tensor) -> torch.Tensor:
    """fp32 projective warp of [L,2] points with a [3,3] matrix (homography.py:86-105)."""
    l = pts_xy.shape[0]
    hom = torch.cat([pts_xy.to(Hm.dtype), torch.ones(l, 1, dtype=Hm.dtype)], -1)      # [L,3]
    wp = torch.bmm(Hm[None], hom[None].permute(0, 2, 1)).permute(0, 2, 1)[0]          # [L,3]
    sc = wp[:, 2:]
    sc[sc == 0] = 1e-6
    return wp[:, :2] / sc


def window_table(centers_xy: torch.Tensor, img_hw: Tuple[int, int], window: int = 5,
                 scale: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """25 window coordinates per centre + validity (common_utils.py:65-91).
    Window slot w = r*window + c carries offset (dx,dy) = ((c-half)*scale, (r-half)*scale).
    Out-of-image float coordinates -> coordinate (0,0), mask False; then ``.long()``."""
    h, w = img_hw
    half = window // 2
    r = torch.arange(window).reshape(-1, 1).repeat(1, window)         # y offset index
    c = r.T                                                           # x offset index
    offs = torch.stack([c, r], -1).float().sub(half).mul(scale).reshape(1, window * window, 2)
    k = centers_xy[:, None, :] + offs                                 # [L,25,2] fp32
    oob = (k[..., 0] < 0) | (k[..., 1] < 0) | (k[..., 0] >= w) | (k[..., 1] >= h)
    k = k.clone()
    k[oob] = 0
    return k.long(), ~oob


def window_token_index(win_xy: torch.Tensor, w_c: int, scale: int = 8) -> torch.Tensor:
    """pixel window coords [L,25,2] int64 -> token index in the other image
    (common_utils.py:171-179: ``keypoints.float() // s`` then ``.long()``)."""
    cell = (win_xy.float() // scale).long()
    return cell[..., 1] * w_c + cell[..., 0]


def ransac_homography(kp0: np.ndarray, kp1: np.ndarray, thr: float = 8.0):
    """geo_module.py:45-48: host OpenCV RANSAC; returns (M float64 [3,3] | None, inlier mask | None)."""
    import cv2
    if len(kp0) <= 8:
        return None, None
    M, mask = cv2.findHomography(kp0, kp1, cv2.RANSAC, thr)
    if M is None:
        return None, None
    return M, mask[:, 0] == 1


def geo_prepare(data: Dict, n: int, hw0_i, hw1_i, hw0_c, hw1_c, cfg) -> List[Dict]:
    """Per-sample geometry (geo_module.py:39-94): inlier sets, cross window tables, anchor maps."""
    scale = int(hw0_i[0] // hw0_c[0])
    out = []
    for b in range(n):
        sel = data["m_bids"] == b
        kp0 = data["mkpts0_c"][sel].long()
        kp1 = data["mkpts1_c"][sel].long()
        M, inl = ransac_homography(kp0.numpy(), kp1.numpy(), cfg["ransac_thr"])
        g = dict(H=M, win0=None, win1=None, wmask0=None, wmask1=None)
        if M is not None:
            inl_t = torch.from_numpy(inl)
            kp0, kp1 = kp0[inl_t], kp1[inl_t]
            Hf = torch.from_numpy(M).to(torch.float32)
            c1 = warp_points(grid_keypoints(hw0_i[0], hw0_i[1], scale), Hf)            # image0 grid -> image1
            g["win1"], g["wmask1"] = window_table(c1, tuple(hw1_i), cfg["window"], scale)
            Hinv = torch.inverse(torch.from_numpy(M)[None])[0].to(torch.float32)     # fp64 inverse, then cast
            c0 = warp_points(grid_keypoints(hw1_i[0], hw1_i[1], scale), Hinv)          # image1 grid -> image0
            g["win0"], g["wmask0"] = window_table(c0, tuple(hw0_i), cfg["window"], scale)
        a0 = torch.zeros(hw0_c[0], hw0_c[1], dtype=torch.bool)
        a1 = torch.zeros(hw1_c[0], hw1_c[1], dtype=torch.bool)
        a0[kp0[:, 1] // scale, kp0[:, 0] // scale] = True
        a1[kp1[:, 1] // scale, kp1[:, 0] // scale] = True
        g["anchor0"], g["anchor1"] = a0.reshape(-1), a1.reshape(-1)
        g["n_inliers"] = int(kp0.shape[0])
        out.append(g)
    return out


def geo_transformer(P: Params, f0: torch.Tensor, f1: torch.Tensor, geo: List[Dict], hw0_c, hw1_c, cfg):
    """model/geo_transformer/transformer.py:89-146.  f0 [N,L,C], f1 [N,S,C] (PE'd CNN maps)."""
    pre = "geo_module.des_transformer"
    nh = cfg["geo_nhead"]
    f0, f1 = f0.clone(), f1.clone()
    for li, name in enumerate(cfg["geo_layers"]):
        lp = f"{pre}.layers.{li}"
        if name == "self":
            for b, g in enumerate(geo):
                x0, x1 = f0[b], f1[b]
                if g["anchor0"].sum() > 0:
                    x0 = encoder_layer(P, lp, x0[None], x0[g["anchor0"]][None], nh, "full", "tanh")[0]
                if g["anchor1"].sum() > 0:
                    x1 = encoder_layer(P, lp, x1[None], x1[g["anchor1"]][None], nh, "full", "tanh")[0]
                f0[b], f1[b] = x0, x1
        else:
            # windows of BOTH images are gathered from the pre-layer features (lines 126-129)
            snap = []
            for b, g in enumerate(geo):
                if g["win0"] is None:
                    snap.append(None)
                    continue
                idx0 = window_token_index(g["win0"], hw0_c[1])     # tokens of image0 seen from image1 grid
                idx1 = window_token_index(g["win1"], hw1_c[1])
                snap.append((f0[b][idx0], f1[b][idx1]))            # [S,25,C], [L,25,C]
            for b, g in enumerate(geo):
                if snap[b] is None:
                    continue
                w0, w1 = snap[b]
                x0 = encoder_layer(P, lp, f0[b][:, None], w1, nh, "full", "tanh", kv_mask=g["wmask1"])
                x1 = encoder_layer(P, lp, f1[b][:, None], w0, nh, "full", "tanh", kv_mask=g["wmask0"])
                f0[b], f1[b] = x0[:, 0], x1[:, 0]
    return f0, f1


# ----------------------------------------------------------------------------
# fine path: fine_preprocess.py:30-74, model/fine_matching2.py:21-126
# ----------------------------------------------------------------------------
def fine_windows(fmap: torch.Tensor, b_ids: torch.Tensor, tok: torch.Tensor, window: int, stride: int) -> torch.Tensor:
    """[N,C,Hf,Wf] -> [M,window^2,C] windows centred on coarse token ``tok`` (fine_preprocess.py:41-56)."""
    n, c = fmap.shape[:2]
    u = F.unfold(fmap, kernel_size=(window, window), stride=stride, padding=window // 2)
    u = u.reshape(n, c, window * window, -1).permute(0, 3, 2, 1)
    return u[b_ids, tok]


def fine_preprocess(P: Params, ff0, ff1, fc0, fc1, b_ids, i_ids, j_ids, window: int, stride: int):
    m = b_ids.shape[0]
    if m == 0:
        e = torch.empty(0, window * window, ff0.shape[1])
        return e, e.clone()
    w0 = fine_windows(ff0, b_ids, i_ids, window, stride)
    w1 = fine_windows(ff1, b_ids, j_ids, window, stride)
    cw = F.linear(torch.cat([fc0[b_ids, i_ids], fc1[b_ids, j_ids]], 0),
                  P["fine_preprocess.down_proj.weight"], P["fine_preprocess.down_proj.bias"])
    merged = F.linear(torch.cat([torch.cat([w0, w1], 0), cw.unsqueeze(1).repeat(1, window * window, 1)], -1),
                      P["fine_preprocess.merge_feat.weight"], P["fine_preprocess.merge_feat.bias"])
    a, b = torch.chunk(merged, 2, dim=0)
    return a, b


def fine_match(conf: torch.Tensor, thr: float, mkpts0_c, mkpts1_c, b_ids, hw0_i, hw0_c, hw0_f, window: int):
    """model/fine_matching2.py:65-126: keep, per coarse match, only the global arg-max cell of
    the WWxWW dual-softmax matrix if it is > thr (mutual-NN holds trivially for a global max)."""
    m = conf.shape[0]
    ww = window * window
    mask = (conf > thr) * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    top = conf.view(m, ww * ww).argmax(1)
    only = torch.zeros(m, ww * ww, dtype=torch.bool)
    only[torch.arange(m), top] = True
    mask = mask * only.view(m, ww, ww)
    row_any, first_j = mask.max(dim=2)
    mi, ii = torch.where(row_any)
    jj = first_j[mi, ii]
    mconf = conf[mi, ii, jj]
    c2f = torch.tensor(hw0_f[0]) / torch.tensor(hw0_c[0])
    coarse_scale = torch.tensor(hw0_i[0]) / torch.tensor(hw0_c[0])
    fine_scale = torch.tensor(hw0_i[0]) / torch.tensor(hw0_f[0])
    c0 = mkpts0_c / coarse_scale * c2f
    c1 = mkpts1_c / coarse_scale * c2f
    half = window // 2
    k0 = (torch.stack([ii % window - half, ii // window - half], 1) + c0[mi]) * fine_scale
    k1 = (torch.stack([jj % window - half, jj // window - half], 1) + c1[mi]) * fine_scale
    return dict(mkpts0_f=k0, mkpts1_f=k1, mconf=mconf, m_bids=b_ids[mi], fine_sel=mi, fine_i=ii, fine_j=jj)


# ----------------------------------------------------------------------------
# full forward: model/full_model.py:39-123
# ----------------------------------------------------------------------------
def forward(P: Params, image0: torch.Tensor, image1: torch.Tensor, cfg: Optional[dict] = None,
            capture: Optional[dict] = None, mask0: Optional[torch.Tensor] = None,
            mask1: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """full_model.py:39-123.  mask0/mask1 [N,h_c,w_c] bool: the optional padding masks of the MegaDepth training
    collation (full_model.py:80-84); they reach the coarse transformer and both coarse-matching calls only (border
    masking with padding is a no-op at border_rm == 0, coarse_matching.py:54-56).  The product refuses masks."""
    cfg = {**DEFAULT_CFG, **(cfg or {})}
    mc0 = None if mask0 is None else mask0.flatten(-2)
    mc1 = None if mask1 is None else mask1.flatten(-2)
    assert (mc0 is None) == (mc1 is None) and (mc0 is None or cfg["border_rm"] == 0)
    cap = capture if capture is not None else {}
    n = image0.shape[0]
    hw0_i, hw1_i = tuple(image0.shape[2:]), tuple(image1.shape[2:])
    if hw0_i == hw1_i:
        fc, ff = backbone(P, torch.cat([image0, image1], 0))
        (c0, c1), (ff0, ff1) = fc.split(n), ff.split(n)
    else:
        (c0, ff0), (c1, ff1) = backbone(P, image0), backbone(P, image1)
    hw0_c, hw1_c = tuple(c0.shape[2:]), tuple(c1.shape[2:])
    hw0_f = tuple(ff0.shape[2:])
    cap.update(cnn_c0=c0, cnn_c1=c1, fine0=ff0, fine1=ff1)

    x0, x1 = add_pe_flatten(c0), add_pe_flatten(c1)
    cap.update(pe0=x0, pe1=x1)
    t0, t1 = local_feature_transformer(P, "loftr_coarse", x0, x1, cfg["coarse_layers"], cfg["coarse_nhead"], mc0, mc1)
    cap.update(coarse0=t0, coarse1=t1)

    conf1 = dual_softmax_conf(t0, t1, cfg["coarse_temperature"], mc0, mc1)
    m1 = coarse_match(conf1, cfg["coarse_thr"], hw0_i, hw0_c, hw1_c, cfg["border_rm"])
    cap.update(conf_first=conf1, first=m1)

    geo = geo_prepare(m1, n, hw0_i, hw1_i, hw0_c, hw1_c, cfg)
    g0, g1 = geo_transformer(P, x0, x1, geo, hw0_c, hw1_c, cfg)   # geo module's own PE == same table
    cap.update(geo=geo, geo0=g0, geo1=g1)

    conf2 = dual_softmax_conf(g0, g1, cfg["coarse_temperature"], mc0, mc1)
    m2 = coarse_match(conf2, cfg["coarse_thr"], hw0_i, hw0_c, hw1_c, cfg["border_rm"])
    cap.update(conf_second=conf2)

    stride = hw0_f[0] // hw0_c[0]
    w0, w1 = fine_preprocess(P, ff0, ff1, g0, g1, m2["b_ids"], m2["i_ids"], m2["j_ids"], cfg["fine_window"], stride)
    cap.update(fine_in0=w0, fine_in1=w1)
    out = dict(m2)
    out.update(hw0_i=hw0_i, hw1_i=hw1_i, hw0_c=hw0_c, hw1_c=hw1_c, hw0_f=hw0_f,
               first_b_ids=m1["b_ids"], first_i_ids=m1["i_ids"], first_j_ids=m1["j_ids"])
    if w0.shape[0] == 0:
        out.update(mkpts0_f=m2["mkpts0_c"], mkpts1_f=m2["mkpts1_c"],
                   fine_matrix=torch.empty(0, cfg["fine_window"] ** 2, cfg["fine_window"] ** 2))
        return out
    w0, w1 = local_feature_transformer(P, "loftr_fine", w0, w1, cfg["fine_layers"], cfg["fine_nhead"])
    cap.update(fine_out0=w0, fine_out1=w1)
    fconf = dual_softmax_conf(w0, w1, cfg["fine_temperature"])
    out["fine_matrix"] = fconf
    out.update(fine_match(fconf, cfg["fine_thr"], m2["mkpts0_c"], m2["mkpts1_c"], m2["b_ids"],
                          hw0_i, hw0_c, hw0_f, cfg["fine_window"]))
    return out
