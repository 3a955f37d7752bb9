"""Parity of the SHIPPED configuration - the one bench.py times: fp16 tcgen05 backbone, tf32 / fp16 tensor-core
projections with fp16 intermediates, fused (non-materialising) coarse matcher, fused fine-layer kernel — against
reference runs at the benchmarked shape (480 x 640), on pairs whose geometry is NOT the identity:

* full_shift_rn_480x640 : image1 = image0 rolled by (16, 8) px, randomised-norm weights -> peaky confidences (max 0.81)
* full_warp_rn_480x640  : image1 = cv2.warpPerspective(image0, H), same weights (max 0.86)
* full_shift_480x640    : same shift, default-norm weights -> flat confidences (max 1e-4): the worst case for match-set
                          identity (SURVEY.md hard part 8), kept to show the spread
* full_dense_480x640    : the bench workload itself (image1 == image0)

The fixtures were produced by the unmodified reference (tests/golden/make_golden.py) and the CPU oracle reproduces
them exactly (tests/test_oracle_golden.py).  Tolerances are stated next to each assertion; measured values are printed.
"""
import copy

import numpy as np
import pytest
import torch

from geoformer_b200 import synth
from oracle import geoformer_oracle as O
from tests.util import load_golden, stage_case_inputs

pytestmark = pytest.mark.gpu


def product_model(sd, capture=True):
    """Exactly what bench.py builds (bench.build_model), plus stage capture."""
    from geoformer_b200 import engine, ops
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    g = dict(geo_cfg)
    g["coarse_thr"] = 0.0
    m = GeoFormer(copy.deepcopy(default_cfg), g)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    m = m.eval().to("cuda:0")
    assert m.backbone_precision == "f16" and not m.materialize and m.ransac == "cv2"
    assert engine.FUSED_MATCHING and engine.FUSED_FINE_LAYER and ops.act16() and ops._SIM_IMPL == "f16x3"
    m.capture = capture
    return m


def _rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def _rms_rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def _iou(a, b):
    a, b = set(map(tuple, np.asarray(a).tolist())), set(map(tuple, np.asarray(b).tolist()))
    return len(a & b) / max(1, len(a | b))


def _corner_pts(k0, k1, w, h):
    """Corners of the frame through the homography estimated as hpatches_helper.py:216 does (RANSAC, 3 px)."""
    import cv2
    Hm, _ = cv2.findHomography(np.asarray(k0, np.float32), np.asarray(k1, np.float32), cv2.RANSAC, 3)
    c = np.array([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]], dtype=np.float64).reshape(-1, 1, 2)
    return cv2.perspectiveTransform(c, Hm)[:, 0], c[:, 0]


# name -> (geo-feature rel-rms bar, first-pass IoU, final coarse IoU, fine IoU); IoU = |A & B| / |A | B| of (i, j) pairs.
# (coarse-transformer features, which precede the discrete RANSAC step, are held to 5e-3 in every case: measured 8.5e-4)
# Measured on B200 (profiles/r02_parity_precision_probe.txt): shift_rn 0.972 / 0.988 / 0.956, shift 0.982 / 0.986 / 0.972,
# warp 0.922 / 0.848 / 0.798.  The warp pair is a smooth texture with only 109 reference matches whose neighbouring tokens
# are near-duplicates: one flipped mutual-nearest-neighbour decision is 1 % of the set, and even an EXACT fp32 backbone
# (everything else product) reaches only 0.938 there - its bars are set accordingly; the geometric bar (0.1 px) is not relaxed.
# Its geo features depend on WHICH anchor tokens RANSAC keeps (8 % of the first-pass matches differ): measured 0.5-1.3e-2.
BARS = {
    "full_shift_rn_480x640": (5e-3, 0.93, 0.93, 0.93),
    "full_shift_480x640": (5e-3, 0.93, 0.93, 0.93),
    "full_warp_rn_480x640": (2.5e-2, 0.85, 0.75, 0.70),
}


@pytest.mark.parametrize("name", list(BARS))
def test_product_mode_480x640_vs_reference_run(golden_dir, name):
    g = load_golden(golden_dir, name)
    feat_bar, iou_first, iou_coarse, iou_fine = BARS[name]
    sd, im0, im1 = stage_case_inputs(g)
    h, w = int(g["meta"][0]), int(g["meta"][1])
    model = product_model(sd)
    data = model({"image0": im0.cuda(), "image1": im1.cuda()})
    st = data["_stages"]
    ts = int(g["tok_stride"])
    # 1. backbone (fp16 storage through 20 chained conv layers): max-abs error <= 5e-3 of the feature range (measured 1.4e-3)
    cnn = torch.cat([st["cnn_c0"], st["cnn_c1"]], 0).permute(0, 3, 1, 2)[:, :, ::4, ::5]
    e_cnn = _rel(cnn, g["cnn_c_sub"])
    # 2. coarse / geo transformer outputs (LayerNorm'd residual stream): relative rms error
    e = {k: _rms_rel(st[k][0, ::ts], g[k + "_sub"]) for k in ("coarse0", "coarse1", "geo0", "geo1")}
    # 3. match sets
    first = st["first"]
    i_first = _iou(torch.stack([first["i_ids"], first["j_ids"]], 1).cpu().numpy(), np.stack([g["first_i"], g["first_j"]], 1))
    i_coarse = _iou(torch.stack([data["i_ids"], data["j_ids"]], 1).cpu().numpy(), np.stack([g["i_ids"], g["j_ids"]], 1))
    got_f = torch.cat([data["mkpts0_f"], data["mkpts1_f"]], 1).long().cpu().numpy()
    want_f = np.concatenate([g["mkpts0_f"], g["mkpts1_f"]], 1).astype(np.int64)
    i_fine = _iou(got_f, want_f)
    # 4. downstream homography: frame corners through H(ours) vs H(reference run)
    c_gpu, c0 = _corner_pts(got_f[:, :2], got_f[:, 2:], w, h)
    c_ref, _ = _corner_pts(want_f[:, :2], want_f[:, 2:], w, h)
    d_corner = np.linalg.norm(c_gpu - c_ref, axis=1).mean()
    print(f"{name}: cnn {e_cnn:.2e}  feats {({k: round(v, 5) for k, v in e.items()})}  IoU first/coarse/fine "
          f"{i_first:.3f}/{i_coarse:.3f}/{i_fine:.3f}  counts {len(got_f)}/{len(want_f)}  corner diff {d_corner:.4f} px")
    assert e_cnn <= 5e-3, e_cnn
    assert max(e["coarse0"], e["coarse1"]) <= 5e-3 and max(e["geo0"], e["geo1"]) <= feat_bar, e
    assert i_first >= iou_first and i_coarse >= iou_coarse and i_fine >= iou_fine, (i_first, i_coarse, i_fine)
    assert d_corner <= 0.1, d_corner                                  # north star: corner error agrees within 0.1 px
    if str(g["regime"]) == "shift":                                   # known non-identity ground truth: translation (16, 8)
        gt = c0 + np.array([16.0, 8.0])
        e_gpu, e_ref = np.linalg.norm(c_gpu - gt, axis=1).mean(), np.linalg.norm(c_ref - gt, axis=1).mean()
        print(f"   corner error vs GT translation: ours {e_gpu:.4f} px, reference {e_ref:.4f} px")
        assert abs(e_gpu - e_ref) <= 0.1 and e_gpu <= 0.2


def test_product_mode_batch_of_two_regimes_vs_reference_runs(golden_dir):
    """Batch > 1 in the shipped configuration: [dense pair, shifted pair] in one forward, each sample against its own
    reference run (same seed-0 weights).  Dense regime (the bench workload): >= 97 % of the reference's coarse (i, j) and
    fine matches (measured 0.998 / 0.991); shifted pair >= 93 %.  mconf (the FINE confidence: a product of two 25-way
    softmaxes at temperature 0.1, so a logit error of 0.05 moves it by several percent where two cells compete) of the
    common fine matches: median |diff| <= 3e-2 and 90th percentile <= 0.15 (measured p90 0.094); the maximum is printed."""
    gd, gs = load_golden(golden_dir, "full_dense_480x640"), load_golden(golden_dir, "full_shift_480x640")
    sd = synth.make_state_dict(0)
    d0, d1 = synth.make_pairs(1, 480, 640, "dense", 0)
    s0, s1 = synth.make_pairs(1, 480, 640, "shift", 10)
    model = product_model(sd, capture=False)
    data = model({"image0": torch.cat([d0, s0]).cuda(), "image1": torch.cat([d1, s1]).cuda()})
    b = data["b_ids"].cpu().numpy()
    ij = torch.stack([data["i_ids"], data["j_ids"]], 1).cpu().numpy()
    mb = data["m_bids"].cpu().numpy()
    kf = torch.cat([data["mkpts0_f"], data["mkpts1_f"]], 1).long().cpu().numpy()
    cf = data["mconf"].cpu().numpy()
    for s, (g, bar) in enumerate(((gd, 0.97), (gs, 0.93))):
        want_ij = np.stack([g["i_ids"], g["j_ids"]], 1)
        want_f = np.concatenate([g["mkpts0_f"], g["mkpts1_f"]], 1).astype(np.int64)
        got_ij, got_f = ij[b == s], kf[mb == s]
        r_c = len(set(map(tuple, got_ij.tolist())) & set(map(tuple, want_ij.tolist()))) / len(want_ij)
        ref_conf = {tuple(k): c for k, c in zip(want_f.tolist(), g["mconf"].tolist())}
        common = [(c, ref_conf[tuple(k)]) for k, c in zip(got_f.tolist(), cf[mb == s].tolist()) if tuple(k) in ref_conf]
        r_f = len(common) / len(want_f)
        dc = np.abs(np.array([a - c for a, c in common]))
        dconf, p90 = float(dc.max()), float(np.percentile(dc, 90))
        print(f"sample {s}: coarse recall {r_c:.3f} ({len(got_ij)}/{len(want_ij)}), fine recall {r_f:.3f} "
              f"({len(got_f)}/{len(want_f)}), IoU coarse {_iou(got_ij, want_ij):.3f} fine {_iou(got_f, want_f):.3f}, |dconf| median {np.median(dc):.2e} p90 {p90:.2e} max {dconf:.2e}")
        assert r_c >= bar and r_f >= bar, (s, r_c, r_f)
        assert float(np.median(dc)) <= 3e-2 and p90 <= 0.15, (float(np.median(dc)), p90, dconf)
    # batch invariance of the shipped configuration: the dense sample alone gives the same matches
    one = model({"image0": d0.cuda(), "image1": d1.cuda()})
    assert torch.equal(one["mkpts0_f"], data["mkpts0_f"][data["m_bids"] == 0])
    assert torch.equal(one["mkpts1_f"], data["mkpts1_f"][data["m_bids"] == 0])


# ------------------------------------------------------------------------------------------------ fused fine layer
def _fine_weights(sd, dev):
    from geoformer_b200 import engine
    return engine.PackedWeights(sd, dev, torch.float16).fine


def test_fine_layer_kernel_vs_reference_golden(golden_dir):
    """gf_fine_layer (the largest kernel of a step) on the REFERENCE's own fine-transformer input -> output pairs
    (fixture small_dense: first 16 windows of loftr_fine's input and output; windows are independent batch entries):
    self layer over both images, then the two cross calls (the second sees the updated feat0, transformer.py:99-100).
    Operands are tf32 / fp16 (10-bit mantissa), fp32 accumulation: 1e-2 max-abs on the O(1) LayerNorm'd stream after
    the two chained layers."""
    from geoformer_b200 import ops
    g = load_golden(golden_dir, "small_dense")
    dev = torch.device("cuda:0")
    ops.ensure_init(dev)
    sd = synth.make_state_dict(7, True)
    fw = _fine_weights(sd, dev)
    x = torch.cat([torch.from_numpy(g["fine_in0"]), torch.from_numpy(g["fine_in1"])], 0).to(dev).contiguous()
    m = x.shape[0] // 2
    run = lambda lw, a, b: ops.fine_layer_fused(a.contiguous(), b.contiguous(), lw["wpack"], lw["n1w"], lw["n1b"], lw["n2w"], lw["n2b"])
    x = run(fw[0], x, x)
    y0 = run(fw[1], x[:m], x[m:])
    y1 = run(fw[1], x[m:], y0)
    e0 = (y0.cpu() - torch.from_numpy(g["fine_out0"])).abs().max().item()
    e1 = (y1.cpu() - torch.from_numpy(g["fine_out1"])).abs().max().item()
    print("fine_layer_kernel vs reference golden: max-abs", e0, e1, "range", float(np.abs(g["fine_out0"]).max()))
    assert e0 <= 1e-2 and e1 <= 1e-2, (e0, e1)


@pytest.mark.parametrize("windows", [4, 333, 2501])
@pytest.mark.parametrize("cross", [False, True])
def test_fine_layer_kernel_vs_oracle(windows, cross):
    """One gf_fine_layer call against the CPU oracle's encoder_layer (transformer.py:37-60 restated) on the same
    inputs and weights: ragged window counts (partial last 5-window tile), self and cross.  8e-3 max-abs (tf32 bar)."""
    from geoformer_b200 import ops
    dev = torch.device("cuda:0")
    ops.ensure_init(dev)
    sd = synth.make_state_dict(7, True)
    fw = _fine_weights(sd, dev)
    gen = torch.Generator().manual_seed(windows)
    x = torch.randn(windows, 25, 128, generator=gen)
    src = torch.randn(windows, 25, 128, generator=gen) if cross else x
    lw = fw[1 if cross else 0]
    with torch.no_grad():
        want = O.encoder_layer(sd, f"loftr_fine.layers.{1 if cross else 0}", x, src, 8)
    xd = x.to(dev)
    got = ops.fine_layer_fused(xd, src.to(dev) if cross else xd, lw["wpack"], lw["n1w"], lw["n1b"], lw["n2w"], lw["n2b"])
    err = (got.cpu() - want).abs().max().item()
    print(f"fine_layer_kernel vs oracle: windows {windows} cross {cross} max-abs {err:.3e}")
    assert torch.isfinite(got).all() and err <= 8e-3, err
