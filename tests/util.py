"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

from geoformer_b200 import synth


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def stage_case_inputs(g):
    """(state dict, image0, image1) of a `full_stage_case` fixture (tests/golden/make_golden.py): synth images, or the
    uint8 pair stored inside the fixture."""
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    P = synth.make_state_dict(seed=int(g["wseed"]), randomize_norm=bool(rnd))
    if "image0_u8" in g:
        im0 = torch.from_numpy(g["image0_u8"]).float().div(255)[None, None]
        im1 = torch.from_numpy(g["image1_u8"]).float().div(255)[None, None]
    else:
        im0, im1 = synth.make_pairs(n, h, w, str(g["regime"]), seed0)
    return P, im0, im1


# ------------------------------------------------------------------------------------------------
# synthetic HPatches-shaped tree + a deterministic stand-in for the matcher wrapper (tests of geoformer_b200.hpatches
# and tests/golden/make_golden.py --only-hpatches, which runs the reference's own eval loop on the same inputs)
# ------------------------------------------------------------------------------------------------
HP_SEQS = (("i_ajuntament", (96, 128)), ("i_castle", (120, 160)), ("v_bird", (96, 128)), ("v_boat", (128, 96)),
           ("v_circus", (96, 128)), ("i_dome", (96, 128)))


def make_hpatches_tree(root, seqs=HP_SEQS, seed=0):
    """<root>/<seq>/{1..6}.ppm + H_1_{2..6}: small random-texture images (contents only matter to real matchers) and
    mild ground-truth homographies."""
    import cv2
    rng = np.random.RandomState(seed)
    for name, (h, w) in seqs:
        d = os.path.join(root, name)
        os.makedirs(d, exist_ok=True)
        base = cv2.GaussianBlur(rng.randint(0, 256, (h, w, 3)).astype(np.uint8), (0, 0), 1.5)
        cv2.imwrite(os.path.join(d, "1.ppm"), base)
        for k in range(2, 7):
            src = np.float32([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]])
            dst = src + rng.uniform(-0.06, 0.06, (4, 2)).astype(np.float32) * np.float32([w, h])
            Hm = cv2.getPerspectiveTransform(src, dst).astype(np.float64)
            np.savetxt(os.path.join(d, "H_1_{}".format(k)), Hm)
            cv2.imwrite(os.path.join(d, "{}.ppm".format(k)), cv2.warpPerspective(base, Hm, (w, h)))
    return root


def stub_matcher(scaled: bool, fail=("v_boat", 3), few=("i_castle", 4), empty=("v_circus", 5)):
    """matcher(im1_path, im2_path) with the wrapper's return convention (geoformer.py:88-99): noisy ground-truth
    correspondences with 20 % outliers, deterministic per pair; one pair raises, one returns 3 matches (no homography),
    one returns none.  scaled=True mimics no_match_upscale: coordinates in a resized frame + the upscale 4-vector."""
    def matcher(im1_path, im2_path):
        from PIL import Image
        seq = os.path.basename(os.path.dirname(im1_path))
        idx = int(os.path.basename(im2_path)[0])
        if (seq, idx) == fail:
            raise RuntimeError("stub matcher: simulated failure on {} {}".format(seq, idx))
        rng = np.random.RandomState(sum(ord(c) for c in seq) * 7 + idx)
        Hm = np.loadtxt(os.path.join(os.path.dirname(im1_path), "H_1_{}".format(idx)))
        w, h = Image.open(im1_path).size
        n = 0 if (seq, idx) == empty else (3 if (seq, idx) == few else int(rng.randint(30, 200)))
        p1 = rng.rand(n, 2) * [w - 1, h - 1]
        q = np.concatenate([p1, np.ones((n, 1))], 1) @ Hm.T
        p2 = q[:, :2] / q[:, 2:] + rng.randn(n, 2) * rng.uniform(0.3, 3.0)
        out = rng.rand(n) < 0.2
        p2[out] = rng.rand(int(out.sum()), 2) * [w - 1, h - 1]
        scores = rng.rand(n).astype(np.float32)
        if not scaled:
            m = np.concatenate([p1, p2], 1).astype(np.float32)
            return m, m[:, :2].copy(), m[:, 2:].copy(), scores
        up = np.array([1.6, 1.6, 1.5, 1.55])
        m = (np.concatenate([p1, p2], 1) / up).astype(np.float32)
        return m, m[:, :2].copy(), m[:, 2:].copy(), scores, up
    return matcher


# ------------------------------------------------------------------------------------------------
# synthetic FIRE / ISC-HE shaped inputs (tests of geoformer_b200.fire_isc and make_golden.py --only-fire-isc)
# ------------------------------------------------------------------------------------------------
def _named_homography(key: str, size: float) -> np.ndarray:
    rng = np.random.RandomState(sum(ord(c) * (i + 1) for i, c in enumerate(key)) % (2 ** 31))
    Hm = np.eye(3)
    Hm[:2, :2] += rng.uniform(-0.03, 0.03, (2, 2))
    Hm[:2, 2] = rng.uniform(-0.05, 0.05, 2) * size
    Hm[2, :2] = rng.uniform(-2e-5, 2e-5, 2) * (512.0 / size)
    return Hm


def _apply_h(Hm, p):
    q = np.concatenate([p, np.ones((len(p), 1))], 1) @ Hm.T
    return q[:, :2] / q[:, 2:]


def make_fire_tree(root, size=2912):
    """<root>/gt/control_points_<cat><id>_1_2.txt for 71 S + 48 P + 14 A pairs (the counts fire_helper.compute_auc
    asserts), 10 control points each: columns (x_refer, y_refer, x_query, y_query).  No images: the loop never opens them."""
    gt = os.path.join(root, "gt")
    os.makedirs(gt, exist_ok=True)
    files = []
    for cat, n in (("S", 71), ("P", 48), ("A", 14)):
        for i in range(1, n + 1):
            key = "{}{:02d}".format(cat, i)
            rng = np.random.RandomState(1000 + len(files))
            raw = rng.rand(10, 2) * (size - 1)                       # query image
            dst = _apply_h(_named_homography(key, size), raw)       # reference image
            name = "control_points_{}_1_2.txt".format(key)
            np.savetxt(os.path.join(gt, name), np.concatenate([dst, raw], 1))
            files.append(name)
    rng = np.random.RandomState(7)
    return [files[i] for i in rng.permutation(len(files))], os.path.join(root, "images"), gt


def make_isc_tree(root, n=24):
    """(query.jpg, refer.jpg, gt.txt) triples with real (small) image files — the loop reads their sizes — and control
    points normalised to [0, 1] (my_helper.py:142-152)."""
    import cv2
    os.makedirs(root, exist_ok=True)
    triples = []
    for i in range(n):
        rng = np.random.RandomState(500 + i)
        (w1, h1), (w2, h2) = (64 + 8 * (i % 3), 48 + 8 * (i % 2)), (80, 56 + 8 * (i % 4))
        q, r, g = (os.path.join(root, "{}_{}".format(i, nm)) for nm in ("q.jpg", "r.jpg", "gt.txt"))
        cv2.imwrite(q, rng.randint(0, 256, (h1, w1, 3)).astype(np.uint8))
        cv2.imwrite(r, rng.randint(0, 256, (h2, w2, 3)).astype(np.uint8))
        raw = rng.rand(8, 2) * [w1 - 1, h1 - 1]
        dst = _apply_h(_named_homography("isc{}".format(i), 64.0), raw)
        np.savetxt(g, np.concatenate([raw / [w1, h1], dst / [w2, h2]], 1))
        triples.append((q, r, g))
    return triples


def stub_matcher_named(kind: str, scaled: bool, size: float = 2912.0, fail_key="P03", few_key="S05"):
    """Stand-in matcher for the FIRE ('fire': paths <dir>/<key>_<k>.jpg, query first) and ISC ('isc': paths <i>_q.jpg /
    <i>_r.jpg) trees: noisy correspondences of the pair's generating homography, wrapper return convention."""
    def matcher(im1_path, im2_path):
        base = os.path.basename(im1_path)
        key = base.split("_")[0] if kind == "fire" else "isc" + base.split("_")[0]
        if key in (fail_key, "isc3"):
            raise RuntimeError("stub matcher: simulated failure on " + key)
        sz = size if kind == "fire" else 64.0
        rng = np.random.RandomState(sum(ord(c) for c in key) * 13 + 1)
        n = 3 if key in (few_key, "isc5") else int(rng.randint(40, 150))
        p1 = rng.rand(n, 2) * (sz - 1) * (1.0 if kind == "fire" else 0.9)
        p2 = _apply_h(_named_homography(key, sz), p1) + rng.randn(n, 2) * rng.uniform(0.2, 2.5) * (sz / 512.0 if kind == "fire" else 0.3)
        out = rng.rand(n) < 0.15
        p2[out] = rng.rand(int(out.sum()), 2) * (sz - 1)
        scores = rng.rand(n).astype(np.float32)
        if not scaled:
            m = np.concatenate([p1, p2], 1).astype(np.float32)
            return m, m[:, :2].copy(), m[:, 2:].copy(), scores
        up = np.array([3.79, 3.79, 3.5, 3.6]) if kind == "fire" else np.array([1.25, 1.2, 1.1, 1.3])
        m = (np.concatenate([p1, p2], 1) / up).astype(np.float32)
        return m, m[:, :2].copy(), m[:, 2:].copy(), scores, up
    return matcher
