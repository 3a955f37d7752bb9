"""Optional GPU RANSAC (csrc/ransac.cu, GeoFormer.ransac = "gpu") against the call it replaces,
cv2.findHomography(kp0, kp1, cv2.RANSAC, 8.0) (reference model/geo_module.py:45-52).  Not bit-identical by design
(different sampling / refit), so the checks are geometric: corner transfer error between the two homographies,
inlier-set overlap, and exact host re-derivation of the anchor lists from the returned inlier flags."""
import numpy as np
import pytest
import torch

from geoformer_b200 import synth

pytestmark = pytest.mark.gpu

HW_I, HW_C, SCALE = (480, 640), (60, 80), 8


def _make_sample(rng, hmat, n_pts, outlier_frac):
    """Coarse-grid matches: p0 on the 8-px grid, p1 = H p0 quantised to the grid (as first-pass matches are)."""
    toks = rng.choice(HW_C[0] * HW_C[1], size=n_pts, replace=False)
    p0 = np.stack([(toks % HW_C[1]) * SCALE, (toks // HW_C[1]) * SCALE], 1).astype(np.float64)
    q = np.concatenate([p0, np.ones((n_pts, 1))], 1) @ hmat.T
    p1 = np.round(q[:, :2] / q[:, 2:] / SCALE) * SCALE
    out = rng.random(n_pts) < outlier_frac
    p1[out] = np.stack([rng.integers(0, HW_C[1], out.sum()), rng.integers(0, HW_C[0], out.sum())], 1) * SCALE
    keep = (p1[:, 0] >= 0) & (p1[:, 0] < HW_I[1]) & (p1[:, 1] >= 0) & (p1[:, 1] < HW_I[0])
    return p0[keep].astype(np.float32), p1[keep].astype(np.float32)


def _corner_dist(ha, hb):
    c = np.array([[0, 0, 1], [639, 0, 1], [639, 479, 1], [0, 479, 1]], dtype=np.float64)
    a, b = c @ ha.T, c @ hb.T
    return np.linalg.norm(a[:, :2] / a[:, 2:] - b[:, :2] / b[:, 2:], axis=1).mean()


def test_gpu_ransac_agrees_with_cv2():
    import cv2
    from geoformer_b200 import ops
    rng = np.random.default_rng(0)
    hs = [np.eye(3),
          np.array([[0.95, 0.05, 12.0], [-0.04, 1.02, 8.0], [2e-5, -1e-5, 1.0]]),
          np.array([[1.1, -0.08, -20.0], [0.06, 0.93, 15.0], [-3e-5, 4e-5, 1.0]]),
          np.array([[0.8, 0.0, 60.0], [0.0, 0.8, 40.0], [0.0, 0.0, 1.0]])]
    samples = [_make_sample(rng, h, 1500, f) for h, f in zip(hs, [0.0, 0.3, 0.5, 0.2])]
    samples.append(_make_sample(rng, hs[1], 6, 0.0))                 # <= 8 matches: no homography (geo_module.py:47)
    k0 = np.concatenate([s[0] for s in samples]); k1 = np.concatenate([s[1] for s in samples])
    counts = np.array([len(s[0]) for s in samples], dtype=np.int32)
    b_ids = np.repeat(np.arange(len(samples)), counts).astype(np.int64)
    dev = "cuda:0"
    ops.ensure_init(torch.device(dev))
    hm, has_h, inl, aidx, acnt = ops.ransac_homography(torch.from_numpy(k0).to(dev), torch.from_numpy(k1).to(dev),
                                                       torch.from_numpy(b_ids).to(dev), torch.from_numpy(counts).to(dev),
                                                       len(samples), HW_C, HW_C, SCALE, 8.0, 1024, seed=0)
    hm, has_h, inl, aidx, acnt = (t.cpu().numpy() for t in (hm, has_h, inl, aidx, acnt))
    assert has_h.tolist() == [1, 1, 1, 1, 0]
    offs = np.concatenate([[0], np.cumsum(counts)])
    for b in range(len(samples)):
        a, c = k0[offs[b]:offs[b + 1]], k1[offs[b]:offs[b + 1]]
        flags = inl[offs[b]:offs[b + 1]].astype(bool)
        if has_h[b]:
            want_h, want_mask = cv2.findHomography(a.astype(np.int64), c.astype(np.int64), cv2.RANSAC, 8.0)
            want_mask = want_mask[:, 0] == 1
            got_h = hm[0, b].reshape(3, 3).astype(np.float64)
            # both estimators see +-4 px grid quantisation and up to 50 % outliers: each must recover the generating
            # homography, the GPU one at least as well as OpenCV (+0.5 px slack), and they must agree with each other
            e_gpu, e_cv = _corner_dist(got_h, hs[b]), _corner_dist(want_h, hs[b])
            assert e_gpu < 1.5 and e_gpu < e_cv + 0.5, (b, e_gpu, e_cv)
            assert _corner_dist(got_h, want_h) < 3.0, (b, _corner_dist(got_h, want_h))
            iou = (flags & want_mask).sum() / max(1, (flags | want_mask).sum())
            assert iou > 0.9, (b, iou)
            prod = got_h @ hm[1, b].reshape(3, 3).astype(np.float64)
            assert np.abs(prod / prod[2, 2] - np.eye(3)).max() < 1e-3
            a, c = a[flags], c[flags]
        # anchors: ascending unique tokens of the (inlier) matches, as the reference's boolean maps
        for side, pts in enumerate((a, c)):
            want = np.unique((pts[:, 1].astype(np.int64) // SCALE) * HW_C[1] + pts[:, 0].astype(np.int64) // SCALE)
            assert acnt[side, b] == len(want)
            assert np.array_equal(aidx[side, b, :len(want)], want)


def test_gpu_ransac_is_deterministic_and_empty_safe():
    from geoformer_b200 import ops
    dev = "cuda:0"
    ops.ensure_init(torch.device(dev))
    rng = np.random.default_rng(1)
    a, c = _make_sample(rng, np.eye(3), 400, 0.2)
    args = lambda: (torch.from_numpy(a).to(dev), torch.from_numpy(c).to(dev), torch.zeros(len(a), dtype=torch.int64, device=dev),
                    torch.tensor([len(a), 0], dtype=torch.int32, device=dev), 2, HW_C, HW_C, SCALE, 8.0, 256)
    r1 = ops.ransac_homography(*args(), seed=3)
    r2 = ops.ransac_homography(*args(), seed=3)
    for x, y in zip(r1, r2):
        assert torch.equal(x, y)
    assert r1[1].tolist() == [1, 0] and r1[4][:, 1].tolist() == [0, 0]      # empty sample: no H, no anchors


def test_forward_with_gpu_ransac_matches_cv2_mode():
    """Full forward with ransac='gpu' vs the default host cv2 mode on the same dense pairs: same set of samples get a
    homography, and the final fine matches overlap almost entirely (anchors / windows differ only at RANSAC's margin)."""
    from tests.test_gpu_forward import build_model, _match_set
    model = build_model(synth.make_state_dict(0), 0.0, backbone="f16", linear="tf32", sim="f16x3")
    model.materialize = False
    im0, im1 = synth.make_pairs(2, 240, 320, "dense", 0)
    d_cv = model({"image0": im0.cuda(), "image1": im1.cuda()})
    model.ransac = "gpu"
    d_gpu = model({"image0": im0.cuda(), "image1": im1.cuda()})
    s_cv, s_gpu = _match_set(d_cv), _match_set(d_gpu)
    assert len(s_cv) > 100
    assert len(s_cv & s_gpu) / len(s_cv | s_gpu) > 0.9, (len(s_cv), len(s_gpu), len(s_cv & s_gpu))
