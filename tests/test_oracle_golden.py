"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Integer outputs must be bit-exact; floats <= 1e-5
(they are bit-identical on the machine that made the fixtures, the slack only
absorbs a different BLAS thread split on another host)."""
import os

import numpy as np
import pytest
import torch

from geoformer_b200 import synth
from oracle import geoformer_oracle as O
from tests.util import stage_case_inputs


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def _run(g, seed_w):
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    P = synth.make_state_dict(seed=seed_w, randomize_norm=bool(rnd))
    im0, im1 = synth.make_pairs(n, h, w, str(g["regime"]), seed0)
    cap = {}
    cfg = dict(coarse_thr=float(g["coarse_thr"]))
    if "fine_thr" in g:
        cfg["fine_thr"] = float(g["fine_thr"])
    with torch.no_grad():
        out = O.forward(P, im0, im1, cfg, capture=cap)
    return out, cap


def _close(a, b, tol=1e-5):
    a = torch.as_tensor(np.asarray(a)).float()
    b = torch.as_tensor(np.asarray(b)).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel():
        err = (a - b).abs().max().item()
        scale = max(1.0, b.abs().max().item())
        assert err <= tol * scale, err


@pytest.mark.parametrize("name", ["small_dense", "small_shift", "small_rect_thr"])
def test_small_cases_stagewise(golden_dir, name):
    g = _load(golden_dir, name)
    out, cap = _run(g, 7)
    n = int(g["meta"][2])
    _close(torch.cat([cap["cnn_c0"], cap["cnn_c1"]], 0), g["cnn_c"])
    _close(torch.cat([cap["fine0"], cap["fine1"]], 0)[:, ::8, ::4, ::4], g["fine_sub"])
    _close(cap["coarse0"], g["coarse0"]); _close(cap["coarse1"], g["coarse1"])
    _close(cap["conf_first"], g["dect_conf"], 1e-6)
    _close(cap["geo0"], g["geo0"]); _close(cap["geo1"], g["geo1"])
    _close(cap["conf_second"], g["conf"], 3e-5)   # geo gathers feed a strided GEMM in the reference -> different BLAS summation order
    for k in ("b_ids", "i_ids", "j_ids", "m_bids"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    for k in ("mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    _close(out["mconf"], g["mconf"], 1e-3)   # fine conf: 2 more transformer layers amplify the 4e-6 geo difference
    if len(g["b_ids"]):
        _close(cap["fine_in0"][:16], g["fine_in0"]); _close(cap["fine_in1"][:16], g["fine_in1"])
        _close(cap["fine_out0"][:16], g["fine_out0"], 1e-4); _close(cap["fine_out1"][:16], g["fine_out1"], 1e-4)
        _close(out["fine_matrix"][:64], g["fine_matrix"], 1e-3)
    else:
        assert out["mkpts0_f"].shape == (0, 2) and out["fine_matrix"].shape == (0, 25, 25)


def test_mixed_batch_match_lists(golden_dir):
    """One batch holding a dense pair, an UNRELATED pair (noise matches -> RANSAC on garbage, windows mostly out of
    bounds) and a shifted pair: the per-sample branches of geo_module.py:45-94 side by side.  The oracle must reproduce
    the reference's integer outputs exactly (slim fixture: match lists + geo features)."""
    g = _load(golden_dir, "small_mixed")
    out, cap = _run(g, 7)
    assert np.bincount(g["b_ids"], minlength=3).min() > 8            # every sample takes the RANSAC branch
    _close(cap["geo0"], g["geo0"]); _close(cap["geo1"], g["geo1"])
    for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    _close(out["mconf"], g["mconf"], 1e-3)


def test_masked_path_oracle_matches_reference(golden_dir):
    """Optional padding masks (MegaDepth training collation; the product refuses them): the oracle's masked linear
    attention and -1e9 logit fill reproduce the reference run (coarse features, integer match lists)."""
    import torch.nn.functional as Fn
    g = _load(golden_dir, "small_masked")
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    P = synth.make_state_dict(seed=7, randomize_norm=bool(rnd))
    im0, im1 = synth.make_pairs(n, h, w, "dense", seed0)
    hc, wc = h // 8, w // 8
    m0, m1 = torch.zeros(n, hc, wc, dtype=torch.bool), torch.zeros(n, hc, wc, dtype=torch.bool)
    for b in range(n):                                   # same valid regions as make_golden.padded_masks
        m0[b, :hc - 2 - b, :wc - 3] = True
        m1[b, :hc - 1, :wc - 4 - b] = True
    im0 = im0 * Fn.interpolate(m0[:, None].float(), scale_factor=8, mode="nearest")
    im1 = im1 * Fn.interpolate(m1[:, None].float(), scale_factor=8, mode="nearest")
    cap = {}
    with torch.no_grad():
        out = O.forward(P, im0, im1, dict(coarse_thr=0.0), capture=cap, mask0=m0, mask1=m1)
    _close(cap["coarse0"], g["coarse0"]); _close(cap["coarse1"], g["coarse1"])
    for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(out[k].numpy(), g[k]), k
    _close(out["mconf"], g["mconf"], 1e-3)
    # (with coarse_thr = 0 the reference itself keeps matches on padded tokens: fully masked rows / columns have a
    # uniform softmax, every entry ties for the mutual maximum and conf = 1/(L*S) > 0 — the real threshold removes them)


def test_full_size_dense_pair(golden_dir):
    """480x640 single pair, dense regime: the reference's final match list is reproduced exactly."""
    g = _load(golden_dir, "full_dense_480x640")
    out, cap = _run(g, 0)
    assert np.array_equal(out["i_ids"].numpy().astype(np.int32), g["i_ids"])
    assert np.array_equal(out["j_ids"].numpy().astype(np.int32), g["j_ids"])
    assert np.array_equal(out["mkpts0_f"].numpy().astype(np.int16), g["mkpts0_f"])
    assert np.array_equal(out["mkpts1_f"].numpy().astype(np.int16), g["mkpts1_f"])
    _close(out["mconf"], g["mconf"], 1e-3)   # fine conf: 2 more transformer layers amplify the 4e-6 geo difference
    _close(cap["conf_second"].max(), g["conf_max"], 1e-6)
    _close(cap["coarse0"][0, :4, :8], g["coarse0_head"]); _close(cap["geo0"][0, :4, :8], g["geo0_head"])
    # outputs are even integer pixel coordinates
    assert (out["mkpts0_f"] % 2 == 0).all()


@pytest.mark.parametrize("name", ["full_shift_480x640", "full_shift_rn_480x640", "full_warp_rn_480x640"])
def test_full_size_stage_cases(golden_dir, name):
    """480x640 pairs with a NON-identity geometry (translation by (16, 8) px; cv2.warpPerspective pair), flat and peaky
    confidence regimes: the oracle reproduces the reference's first-pass and final match lists exactly and its
    (subsampled) stage features to fp32 round-off.  These fixtures pin the product-mode GPU tests."""
    g = _load(golden_dir, name)
    P, im0, im1 = stage_case_inputs(g)
    cap = {}
    with torch.no_grad():
        out = O.forward(P, im0, im1, dict(coarse_thr=0.0), capture=cap)
    ts = int(g["tok_stride"])
    _close(torch.cat([cap["cnn_c0"], cap["cnn_c1"]], 0)[:, :, ::4, ::5], g["cnn_c_sub"])
    _close(cap["coarse0"][0, ::ts], g["coarse0_sub"]); _close(cap["coarse1"][0, ::ts], g["coarse1_sub"])
    _close(cap["geo0"][0, ::ts], g["geo0_sub"]); _close(cap["geo1"][0, ::ts], g["geo1_sub"])
    assert np.array_equal(out["first_i_ids"].numpy().astype(np.int32), g["first_i"])
    assert np.array_equal(out["first_j_ids"].numpy().astype(np.int32), g["first_j"])
    assert np.array_equal(out["i_ids"].numpy().astype(np.int32), g["i_ids"])
    assert np.array_equal(out["j_ids"].numpy().astype(np.int32), g["j_ids"])
    assert np.array_equal(out["mkpts0_f"].numpy().astype(np.int16), g["mkpts0_f"])
    assert np.array_equal(out["mkpts1_f"].numpy().astype(np.int16), g["mkpts1_f"])
    _close(out["mconf"], g["mconf"], 1e-3)


def test_position_encoding_bug_compat():
    """position_encoding.py:28: exponent is -2k (operator precedence), not the textbook one."""
    pe = O.position_encoding(256, 4, 6)
    x = torch.arange(1, 7).float()
    for k in (0, 1, 5):
        assert torch.allclose(pe[4 * k, 0], torch.sin(x * np.exp(-2.0 * k)), atol=1e-6)
        assert torch.allclose(pe[4 * k + 3, :, 0], torch.cos(torch.arange(1, 5).float() * np.exp(-2.0 * k)), atol=1e-6)


def test_state_dict_schema():
    sd = synth.make_state_dict(0)
    assert len(sd) == 253
    assert sum(v.numel() for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k) == 14187504
