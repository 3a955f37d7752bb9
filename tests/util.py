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
