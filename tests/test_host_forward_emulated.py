"""The HOST side of the path on CPU: GeoFormer._forward / geoformer_b200.engine run end to end with every C-ABI operator
replaced by a torch emulation of its contract (tests/emu_ops.py) and are held against the goldens the unmodified
reference produced (tests/golden/make_golden.py).  What this pins without a GPU: stage order, which operator sees which
buffer view / weight pack / count / skip flag, the 2n-sample batching of same-size pairs, the updated-feat0 cross order,
the per-sample geo branches (RANSAC / no homography / no match), the rectangular L != S path, the zero-match corner and the
optional padding masks.  The kernels' own arithmetic is the `-m gpu` tests' job."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from geoformer_b200 import engine, ops, synth
from oracle import geoformer_oracle as O
from tests import emu_ops as emu
from tests.util import load_golden


def _model(monkeypatch, sd, coarse_thr, mode):
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    emu.install(monkeypatch)
    g = dict(geo_cfg)
    g["coarse_thr"] = coarse_thr
    m = GeoFormer(copy.deepcopy(default_cfg), g)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    m = m.eval()
    if mode == "accurate":          # the configuration of the golden-match GPU tests
        m.backbone_precision = "fp32"
        ops.set_precision(linear="ref", similarity="ref", attention="ref", activations="f32")
    else:                           # the shipped configuration (bench.py): fp16 backbone / intermediates, fused kernels
        m.backbone_precision = "f16"
        assert ops.act16() and engine.FUSED_MATCHING and engine.FUSED_FINE_LAYER
    m.capture = True
    return m


def _forward(m, im0, im1, **extra):
    data = {"image0": im0, "image1": im1, **extra}
    with torch.no_grad():
        return m._forward(data, im0, im1)        # forward() itself refuses CPU tensors (no CPU product path)


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).float(), torch.as_tensor(np.asarray(b)).float()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


def _final_set(d):
    k = torch.cat([torch.as_tensor(np.asarray(d["mkpts0_f"])).float(), torch.as_tensor(np.asarray(d["mkpts1_f"])).float()], 1)
    return {(int(b), *r) for b, r in zip(np.asarray(d["m_bids"]).tolist(), k.long().tolist())}


@pytest.mark.parametrize("name", ["small_dense", "small_shift", "small_mixed"])
def test_accurate_configuration_reproduces_reference_golden(monkeypatch, golden_dir, name):
    g = load_golden(golden_dir, name)
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    sd = synth.make_state_dict(7, bool(rnd))
    m = _model(monkeypatch, sd, float(g["coarse_thr"]), "accurate")
    im0, im1 = synth.make_pairs(n, h, w, str(g["regime"]), seed0)
    d = _forward(m, im0, im1)
    st = d["_stages"]
    assert _rel(st["geo0"], g["geo0"]) <= 1e-4 and _rel(st["geo1"], g["geo1"]) <= 1e-4
    if "coarse0" in g:
        assert _rel(torch.cat([st["cnn_c0"], st["cnn_c1"]], 0).permute(0, 3, 1, 2), g["cnn_c"]) <= 2e-5
        assert _rel(st["coarse0"], g["coarse0"]) <= 5e-5 and _rel(st["coarse1"], g["coarse1"]) <= 5e-5
        assert _rel(st["fine_out0"][:16], g["fine_out0"]) <= 1e-3
    for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(d[k].numpy(), g[k]), k
    assert np.abs(d["mconf"].numpy() - g["mconf"]).max() <= 5e-3
    assert d["mkpts0_f"].dtype == torch.float32 and d["m_bids"].dtype == torch.int64 and d["b_ids"].dtype == torch.int64
    # same-size pairs: both images ride one 2n-sample batch; default inputs never touch the mask operators
    assert emu.CALLS.get("mask_rows_", 0) == 0 and emu.CALLS.get("mask_fill_sim_", 0) == 0
    assert emu.CALLS.get("fine_layer_fused", 0) == 0 and emu.CALLS["linattn_window"] == 3       # per-op fine layers (fp32 mode)


def test_shipped_configuration_routes_and_stays_close(monkeypatch, golden_dir):
    """The configuration bench.py times takes the fused routes (fused matcher twice, one fused kernel per fine layer
    call, fp16 Q|K|V and fine windows) and — with exact arithmetic behind the contracts — lands on the reference's matches
    up to the fp16 storage of backbone activations and intermediates."""
    g = load_golden(golden_dir, "small_dense")
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    m = _model(monkeypatch, synth.make_state_dict(7, bool(rnd)), 0.0, "shipped")
    im0, im1 = synth.make_pairs(n, h, w, "dense", seed0)
    d = _forward(m, im0, im1)
    st = d["_stages"]
    assert emu.CALLS["coarse_match_fused"] == 2 and emu.CALLS.get("similarity", 0) == 0
    assert emu.CALLS["fine_layer_fused"] == 3 and emu.CALLS.get("linattn_window", 0) == 0
    assert emu.CALLS["stem_conv"] == 1 and emu.CALLS.get("conv_ref", 0) == 0
    assert st["fine0"].dtype == torch.float16 and st["fine_in"].dtype == torch.float32
    assert _rel(st["coarse0"], g["coarse0"]) <= 1e-2 and _rel(st["geo0"], g["geo0"]) <= 1e-2
    got, want = _final_set(d), _final_set(g)
    assert len(got & want) >= 0.9 * len(want), (len(got), len(want), len(got & want))
    assert "conf_matrix" not in d                      # the L x S matrix only exists with materialize=True


def test_zero_match_corner_and_materialized_keys(monkeypatch, golden_dir):
    g = load_golden(golden_dir, "small_rect_thr")
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    m = _model(monkeypatch, synth.make_state_dict(7, bool(rnd)), 0.2, "accurate")
    m.materialize = True
    im0, im1 = synth.make_pairs(n, h, w, "dense", seed0)
    d = _forward(m, im0, im1)
    assert d["b_ids"].numel() == 0 and d["mkpts0_f"].shape == (0, 2) and d["mconf"].numel() == 0
    assert d["fine_matrix"].shape == (0, 25, 25) and d["conf_matrix"].shape == (n, (h // 8) * (w // 8), (h // 8) * (w // 8))
    assert _rel(d["_stages"]["geo0"], g["geo0"]) <= 2e-5            # no match: geo layers skipped, PE'd CNN features
    assert emu.CALLS.get("geo_self_attention", 0) == 0 and emu.CALLS.get("geo_cross_attention", 0) == 0
    assert emu.CALLS.get("fine_match", 0) == 0


def test_rectangular_pair_vs_oracle(monkeypatch):
    """image0 96x128 against image1 64x96 (L = 192, S = 96): two backbone passes, per-image layer calls, rectangular
    matching, windows bounds-checked against each image's own size."""
    sd = synth.make_state_dict(7, True)
    m = _model(monkeypatch, sd, 0.0, "accurate")
    a, b = synth.make_image(96, 128, 1), synth.make_image(64, 96, 2)
    d = _forward(m, a, b)
    with torch.no_grad():
        want = O.forward(sd, a, b, dict(coarse_thr=0.0))
    for k in ("b_ids", "i_ids", "j_ids", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(d[k].numpy(), want[k].numpy()), k
    assert tuple(d["hw0_c"].tolist()) == (12, 16) and tuple(d["hw1_c"].tolist()) == (8, 12)
    assert emu.CALLS["stem_conv" if m.backbone_precision == "f16" else "conv_ref"] > 0


def _padded_case(n, h, w, seed0):
    """Inputs of tests/golden/make_golden.py::masked_case."""
    im0, im1 = synth.make_pairs(n, h, w, "dense", seed0)
    hc, wc = h // 8, w // 8
    m0, m1 = torch.zeros(n, hc, wc, dtype=torch.bool), torch.zeros(n, hc, wc, dtype=torch.bool)
    for b in range(n):
        m0[b, :hc - 2 - b, :wc - 3] = True
        m1[b, :hc - 1, :wc - 4 - b] = True
    im0 = im0 * F.interpolate(m0[:, None].float(), scale_factor=8, mode="nearest")
    im1 = im1 * F.interpolate(m1[:, None].float(), scale_factor=8, mode="nearest")
    return im0, im1, m0, m1


def test_padding_masks_reproduce_reference_golden(monkeypatch, golden_dir):
    """data['mask0'] / data['mask1'] (full_model.py:79-83): masked linear attention in all 8 coarse layers (self layers
    with the image's own mask, cross layers with query / source masks of the two images), -1e9 fill before both dual
    softmaxes; geo module and fine level unchanged.  Integer outputs identical to the reference run."""
    g = load_golden(golden_dir, "small_masked")
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    m = _model(monkeypatch, synth.make_state_dict(7, bool(rnd)), 0.0, "accurate")
    im0, im1, m0, m1 = _padded_case(n, h, w, seed0)
    d = _forward(m, im0, im1, mask0=m0, mask1=m1)
    st = d["_stages"]
    assert _rel(st["coarse0"], g["coarse0"]) <= 5e-5 and _rel(st["coarse1"], g["coarse1"]) <= 5e-5
    for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(d[k].numpy(), g[k]), k
    assert np.abs(d["mconf"].numpy() - g["mconf"]).max() <= 5e-3
    # 4 self layers: one pass over the 2n-sample Q|K|V buffer each; 4 cross layers: q and kv of both directions
    assert emu.CALLS["mask_rows_"] == 4 * 1 + 4 * 4 and emu.CALLS["mask_fill_sim_"] == 2
    assert emu.CALLS.get("coarse_match_fused", 0) == 0


def test_padding_masks_in_the_shipped_configuration(monkeypatch, golden_dir):
    """Same inputs through the fp16-storage route: the masks clear fp16 Q|K|V rows, the matcher takes the materialising
    kernels (the fused one never holds the matrix to fill); >= 80 % of the reference's final matches (measured 87 %: fp16
    storage of the backbone activations and of Q|K|V on 12 x 16-token images)."""
    g = load_golden(golden_dir, "small_masked")
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    m = _model(monkeypatch, synth.make_state_dict(7, bool(rnd)), 0.0, "shipped")
    im0, im1, m0, m1 = _padded_case(n, h, w, seed0)
    d = _forward(m, im0, im1, mask0=m0, mask1=m1)
    assert emu.CALLS["mask_fill_sim_"] == 2 and emu.CALLS.get("coarse_match_fused", 0) == 0
    got, want = _final_set(d), _final_set(g)
    assert len(got & want) >= 0.8 * len(want), (len(got), len(want), len(got & want))


def test_padding_masks_different_image_sizes_vs_oracle(monkeypatch):
    """Masks on a rectangular pair (per-image layer calls instead of the 2n-sample batch)."""
    sd = synth.make_state_dict(7, True)
    m = _model(monkeypatch, sd, 0.0, "accurate")
    a, b = synth.make_image(96, 128, 1), synth.make_image(64, 96, 2)
    m0 = torch.ones(1, 12, 16, dtype=torch.bool); m0[:, 10:] = False
    m1 = torch.ones(1, 8, 12, dtype=torch.bool); m1[:, :, 9:] = False
    d = _forward(m, a, b, mask0=m0, mask1=m1)
    with torch.no_grad():
        want = O.forward(sd, a, b, dict(coarse_thr=0.0), mask0=m0, mask1=m1)
    for k in ("b_ids", "i_ids", "j_ids", "mkpts0_f", "mkpts1_f"):
        assert np.array_equal(d[k].numpy(), want[k].numpy()), k
    assert emu.CALLS["mask_rows_"] == 4 * 2 + 4 * 4


def test_mask_argument_checks():
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    m = GeoFormer(copy.deepcopy(default_cfg), dict(geo_cfg)).eval()
    x = torch.zeros(1, 1, 32, 32)
    with pytest.raises(ValueError):                  # full_model.py:82-83 reads both masks
        m({"image0": x, "image1": x, "mask0": torch.ones(1, 4, 4, dtype=torch.bool)})
    with pytest.raises(ValueError):
        emu.token_mask(torch.ones(1, 3, 4, dtype=torch.bool), 1, 16)


@pytest.mark.parametrize("hw0,hw1,n,regime,thr", [
    ((64, 96), (64, 96), 3, "mixed", 0.02),          # a real threshold: some samples lose most of their matches
    ((56, 120), (56, 120), 2, "shift", 0.0),         # 7 x 15 coarse cells (odd sizes)
    ((64, 64), (96, 64), 2, "rect", 0.0),            # batch of rectangular pairs, L = 64 vs S = 96
    ((64, 96), (64, 96), 2, "unrelated", 0.0),       # noise matches: RANSAC on garbage, windows mostly out of bounds
])
def test_shape_and_regime_sweep_vs_oracle(monkeypatch, hw0, hw1, n, regime, thr):
    """Integer outputs of the host path (accurate configuration, operators emulated) equal the oracle's over image
    shapes, batch sizes, thresholds and pair regimes beyond the golden fixtures."""
    sd = synth.make_state_dict(7, True)
    m = _model(monkeypatch, sd, thr, "accurate")
    if regime in ("rect", "unrelated"):
        a = torch.cat([synth.make_image(hw0[0], hw0[1], 10 + i) for i in range(n)], 0)
        b = torch.cat([synth.make_image(hw1[0], hw1[1], 20 + i) for i in range(n)], 0)
    else:
        a, b = synth.make_pairs(n, hw0[0], hw0[1], regime, 5)
    d = _forward(m, a, b)
    with torch.no_grad():
        want = O.forward(sd, a, b, dict(coarse_thr=thr))
    assert want["b_ids"].numel() > 20
    for k in ("b_ids", "i_ids", "j_ids", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f", "m_bids"):
        assert np.array_equal(d[k].numpy(), want[k].numpy()), k


def test_in_place_weight_edits_invalidate_the_packed_weights(monkeypatch):
    """The kernel-ready weight pack follows in-place parameter edits (version counters), load_state_dict and is built
    once otherwise."""
    sd = synth.make_state_dict(7, True)
    m = _model(monkeypatch, sd, 0.0, "accurate")
    a, b = synth.make_pairs(1, 64, 96, "shift", 3)
    d0 = _forward(m, a, b)
    pw0 = m._packed
    d1 = _forward(m, a, b)
    assert m._packed is pw0 and torch.equal(d0["mkpts0_f"], d1["mkpts0_f"])            # no repack between identical calls
    with torch.no_grad():
        m.loftr_coarse.layers[3].mlp[0].weight.mul_(-1.0)                                # in-place edit of one weight
    d2 = _forward(m, a, b)
    assert m._packed is not pw0
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["loftr_coarse.layers.3.mlp.0.weight"] *= -1.0
    with torch.no_grad():
        want = O.forward(sd2, a, b, dict(coarse_thr=0.0))
    assert np.array_equal(d2["mkpts0_f"].numpy(), want["mkpts0_f"].numpy()) and np.array_equal(d2["j_ids"].numpy(), want["j_ids"].numpy())
    assert not torch.equal(d2["_stages"]["coarse0"], d0["_stages"]["coarse0"])
    pw2 = m._packed
    m.load_state_dict({k: v.clone() for k, v in sd.items()})                              # back to the original weights
    d3 = _forward(m, a, b)
    assert m._packed is not pw2 and torch.equal(d3["mkpts0_f"], d0["mkpts0_f"]) and torch.equal(d3["j_ids"], d0["j_ids"])


def test_device_ransac_branch_host_glue(monkeypatch):
    """GeoFormer.ransac == 'gpu': the host glue around gf_ransac_homography (device-resident match lists in, 3n counters
    back, anchor lists / homographies used straight from the device buffers) gives the cv2-mode result when the
    estimator behind the contract is OpenCV itself."""
    sd = synth.make_state_dict(7, True)
    a, b = synth.make_pairs(3, 64, 96, "mixed", 4)
    m = _model(monkeypatch, sd, 0.0, "accurate")
    d_cv = _forward(m, a, b)
    m.ransac = "gpu"
    d_gpu = _forward(m, a, b)
    assert emu.CALLS["ransac_homography"] == 1
    for k in ("b_ids", "i_ids", "j_ids", "mkpts0_f", "mkpts1_f", "m_bids"):
        assert torch.equal(d_cv[k], d_gpu[k]), k
    gi_cv, gi_gpu = d_cv["_stages"]["geo_info"], d_gpu["_stages"]["geo_info"]
    assert np.array_equal(gi_cv["has_h"], gi_gpu["has_h"]) and np.array_equal(gi_cv["anchor_cnt"], gi_gpu["anchor_cnt"])
    assert np.allclose(gi_cv["hmat"], gi_gpu["hmat"])


def test_per_instance_precision_scope(monkeypatch):
    """GeoFormer.precision: the instance's options apply during ITS forward only (module-level options restored
    afterwards, also on error), overrides of different instances never mix, and the default (None) takes no lock."""
    import contextlib
    import threading
    import types
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    seen = []

    def fake_forward(self, data, a, b):
        seen.append((self.tag, ops._LINEAR_IMPL, ops._SIM_IMPL, ops._ATTN_IMPL, ops.act16()))
        if data.get("boom"):
            raise RuntimeError("boom")
        return data
    monkeypatch.setattr(GeoFormer, "_forward", fake_forward)
    img = types.SimpleNamespace(is_cuda=True, device="cpu")
    acc, fast = (GeoFormer(copy.deepcopy(default_cfg), dict(geo_cfg)).eval() for _ in range(2))
    acc.tag, fast.tag = "acc", "fast"
    acc.precision = dict(linear="ref", similarity="ref", attention="ref", activations="f32")
    acc({"image0": img, "image1": img}); fast({"image0": img, "image1": img}); acc({"image0": img, "image1": img})
    assert seen == [("acc", "ref", "ref", "ref", False), ("fast", "tf32", "f16x3", "tf32", True), ("acc", "ref", "ref", "ref", False)]
    with pytest.raises(RuntimeError):
        acc({"image0": img, "image1": img, "boom": True})
    assert (ops._LINEAR_IMPL, ops._SIM_IMPL, ops._ATTN_IMPL, ops.act16()) == ("tf32", "f16x3", "tf32", True)     # restored
    # concurrent calls: an override never leaks into another thread's overridden forward
    seen.clear()
    other = GeoFormer(copy.deepcopy(default_cfg), dict(geo_cfg)).eval()
    other.tag, other.precision = "other", dict(linear="tf32", similarity="ref", attention="tf32", activations="f32")
    ts = [threading.Thread(target=lambda m=m: [m({"image0": img, "image1": img}) for _ in range(50)]) for m in (acc, other)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert len(seen) == 100
    assert all(r[1:] == (("ref", "ref", "ref", False) if r[0] == "acc" else ("tf32", "ref", "tf32", False)) for r in seen)
    assert ops._PRECISION_LOCK.acquire(blocking=False); ops._PRECISION_LOCK.release()
