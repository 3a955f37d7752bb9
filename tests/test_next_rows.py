"""'Next' rows of SURVEY.md §8f around the hot path: image ingest (rank 4) and downstream evaluation (rank 3).
Golden vectors come from the reference's own helper functions (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from geoformer_b200 import evaluate as E
from geoformer_b200.ingest import resize_dims
from oracle import eval_oracle as EO


@pytest.fixture(scope="module")
def g(golden_dir):
    z = np.load(os.path.join(golden_dir, "eval_ingest.npz"))
    return {k: z[k] for k in z.files}


def test_error_auc_and_reproj_match_reference(g):
    errs = g["errs"]
    assert np.array_equal(E.error_auc(errs[~np.isnan(errs)], g["thr"]), g["auc"])
    assert np.array_equal(E.error_auc(np.array([]), g["thr"]), g["auc_empty"])
    assert np.array_equal(E.reproj_dists(g["p1"], g["p2"], g["H"]), g["dists"])
    fire = [E.fire_auc(g["fire_s"]), E.fire_auc(g["fire_p"]), E.fire_auc(g["fire_a"])]
    assert np.allclose(fire + [sum(fire) / 3.0], g["fire"], rtol=0, atol=1e-15)


def test_resize_dims_match_reference(g):
    for wo, ho, imsize, df, wt, ht, sx, sy in g["dims"]:
        got = resize_dims(int(wo), int(ho), imsize=int(imsize), dfactor=int(df), value_to_scale=min)
        assert got == (int(wt), int(ht), (sx, sy))


def test_oracle_resize_is_cv2_bit_exact(g):
    import cv2
    ht, wt = [int(v) for v in g["im_hw"]]
    want = (g["im_resized"][0, 0] * 255).round().astype(np.uint8)
    assert np.array_equal(EO.resize_gray_u8(g["im"], wt, ht), want)
    rng = np.random.RandomState(1)
    for (ho, wo, h2, w2) in [(768, 1024, 480, 640), (97, 211, 64, 96), (480, 640, 480, 640)]:
        im = rng.randint(0, 256, (ho, wo)).astype(np.uint8)
        assert np.array_equal(EO.resize_gray_u8(im, w2, h2), cv2.resize(im, (w2, h2)))


def test_batched_corner_errors_and_homographies():
    import cv2
    rng = np.random.RandomState(5)
    matches, H_gt, sizes = [], [], []
    for i in range(12):
        Hm = np.eye(3) + rng.randn(3, 3) * np.array([[1e-2, 1e-2, 3], [1e-2, 1e-2, 3], [1e-5, 1e-5, 0]])
        p = rng.rand(60 if i != 3 else 2, 2) * 400
        q = np.concatenate([p, np.ones((len(p), 1))], 1) @ Hm.T
        q = q[:, :2] / q[:, 2:] + rng.randn(len(p), 2) * 0.3
        matches.append(np.concatenate([p, q], 1).astype(np.float32))
        H_gt.append(Hm); sizes.append([640 + i, 480 - i])
    res = E.estimate_homographies(matches, 3.0)
    for m, (Hp, inl) in zip(matches, res):
        if len(m) < 4:
            assert Hp is None
            continue
        Hd, _ = cv2.findHomography(m[:, :2], m[:, 2:4], cv2.RANSAC, 3.0)
        assert np.array_equal(Hp, Hd)                    # same call, thread pool only
    got = E.corner_errors([r[0] for r in res], H_gt, np.array(sizes))
    want = np.array([EO.corner_error_one(r[0], Hg, s[0], s[1]) for r, Hg, s in zip(res, H_gt, sizes)])
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12, equal_nan=True) and np.isnan(got[3])
    summ = E.homography_summary(got)
    assert summ["failed"] == 1 and summ["accuracy"].shape == (4,) and np.all(np.diff(summ["accuracy"]) >= 0)


@pytest.mark.gpu
def test_gpu_ingest_bit_exact_vs_cv2(g):
    import cv2
    from geoformer_b200.ingest import gray_to_tensor
    dev = torch.device("cuda:0")
    t, sc = gray_to_tensor(g["im"], dev, imsize=200, dfactor=8, value_to_scale=min)
    assert torch.equal(t.cpu(), torch.from_numpy(g["im_resized"]))        # == reference loader output, bit for bit
    rng = np.random.RandomState(2)
    for (ho, wo, imsize) in [(768, 1024, 480), (2912, 2912, 768), (600, 800, 480), (480, 640, 480), (333, 517, 480)]:
        im = rng.randint(0, 256, (ho, wo)).astype(np.uint8)
        t, sc = gray_to_tensor(im, dev, imsize=imsize, dfactor=8, value_to_scale=min)
        wt, ht, sc2 = resize_dims(wo, ho, imsize=imsize, dfactor=8, value_to_scale=min)
        want = torch.from_numpy(cv2.resize(im, (wt, ht))).float().div(255)[None, None]
        assert sc == sc2 and torch.equal(t.cpu(), want), (ho, wo)
