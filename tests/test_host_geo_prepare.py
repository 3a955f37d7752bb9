"""Host side of the geo module (geoformer_b200.engine.geo_prepare_host): the batch-vectorised bookkeeping around
cv2.findHomography must equal the per-sample restatement of reference model/geo_module.py:39-94 bit for bit
(homography, fp64 inverse cast to fp32, anchor lists in ascending token order, <= 8 matches -> no homography)."""
import numpy as np
import torch


def test_geo_prepare_host_equals_per_sample_restatement():
    import cv2
    from geoformer_b200 import engine
    rng = np.random.default_rng(0)
    hw, scale = (60, 80), 8
    ks0, ks1, counts = [], [], [1500, 5, 0, 4000, 800, 9]
    for b, m in enumerate(counts):
        toks = rng.choice(hw[0] * hw[1], size=m, replace=False)
        p0 = np.stack([(toks % hw[1]) * scale, (toks // hw[1]) * scale], 1).astype(np.float32)
        p1 = p0 + (scale if b % 2 else 0)
        out = rng.random(m) < 0.3
        p1[out] = np.stack([rng.integers(0, hw[1], out.sum()), rng.integers(0, hw[0], out.sum())], 1) * scale
        ks0.append(p0); ks1.append(np.clip(p1, 0, [(hw[1] - 1) * scale, (hw[0] - 1) * scale]).astype(np.float32))
    hm, has_h, aidx, acnt = engine.geo_prepare_host(np.concatenate(ks0), np.concatenate(ks1), np.array(counts), hw, hw,
                                                    scale, 8.0)
    assert has_h.tolist() == [1, 0, 0, 1, 1, 1]
    for b in range(len(counts)):
        a, c = ks0[b].astype(np.int64), ks1[b].astype(np.int64)
        M = mask = None
        if len(a) > 8:                                               # geo_module.py:47
            M, mask = cv2.findHomography(a, c, cv2.RANSAC, 8.0)
        if M is not None:
            assert np.array_equal(hm[0, b], M.astype(np.float32).reshape(9))
            assert np.array_equal(hm[1, b], torch.inverse(torch.from_numpy(M)[None])[0].float().numpy().reshape(9))
            a, c = a[mask[:, 0] == 1], c[mask[:, 0] == 1]
        for side, pts in enumerate((a, c)):
            want = np.unique((pts[:, 1] // scale) * hw[1] + pts[:, 0] // scale)
            assert acnt[side, b] == len(want)
            assert np.array_equal(aidx[side, b, :len(want)], want)
