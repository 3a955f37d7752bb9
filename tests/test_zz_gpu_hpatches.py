"""GPU run of geoformer_b200.hpatches: BatchedMatcher (decode -> shape buckets -> GPU ingest on the batch's stream ->
MatchPipeline -> per-pair wrapper tuples) against the wrapper's per-pair sequence on the same model, and the whole
benchmark loop end to end on a synthetic HPatches-shaped tree.  The metric arithmetic itself is pinned on CPU against the
reference's own loop (tests/test_host_hpatches.py).
(This file sorts after the established suite on purpose: it was written in a session without GPU access.)"""
import copy

import numpy as np
import pytest
import torch

from geoformer_b200 import hpatches as HP
from geoformer_b200 import synth
from tests.util import make_hpatches_tree

pytestmark = pytest.mark.gpu


def _model():
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    g = dict(geo_cfg)
    g["coarse_thr"] = 0.0
    m = GeoFormer(copy.deepcopy(default_cfg), g)
    m.load_state_dict(synth.make_state_dict(7, True), strict=True)
    return m.eval().to("cuda:0")


def test_batched_matcher_vs_per_pair_wrapper_sequence(tmp_path):
    import cv2
    from geoformer_b200.ingest import resize_dims
    root = make_hpatches_tree(str(tmp_path), seqs=(("i_a", (120, 160)), ("v_b", (128, 96)), ("v_c", (96, 128))), seed=5)
    model = _model()
    pairs = [(p.im1, p.im2) for p in HP.list_pairs(root)]
    pairs.append((pairs[0][0], str(tmp_path / "missing.ppm")))
    bm = HP.BatchedMatcher(model, "cuda:0", imsize=96, no_match_upscale=True, batch=4, depth=2)
    got = dict(bm.match_many(pairs))
    assert sorted(got) == list(range(16)) and isinstance(got[15], FileNotFoundError)
    same = total = 0
    for k in (0, 3, 7, 12):
        ims, scs = [], []
        for p in pairs[k]:
            im = cv2.imread(p, cv2.IMREAD_GRAYSCALE)                       # data_io.py:48-62 on the host
            wt, ht, sc = resize_dims(im.shape[1], im.shape[0], imsize=96, dfactor=8, value_to_scale=min)
            ims.append(torch.from_numpy(cv2.resize(im, (wt, ht))).float().div(255)[None, None].cuda()); scs.append(sc)
        d = model({"image0": ims[0], "image1": ims[1]})
        want = np.concatenate([d["mkpts0_f"].cpu().numpy(), d["mkpts1_f"].cpu().numpy()], 1)
        matches, k1, k2, scores, upscale = got[k]
        assert np.allclose(upscale, np.array(scs[0] + scs[1])) and len(scores) == len(matches)
        a, b = {tuple(r) for r in want.tolist()}, {tuple(r) for r in matches.tolist()}
        same += len(a & b); total += len(a | b)
    print(f"batched vs per-pair: {same} of {total} matches identical")
    assert total > 100 and same >= 0.9 * total, (same, total)
    # a second call reuses the pipeline; the single-pair form has the wrapper's signature
    m, a, b, s, up = bm(*pairs[1])
    assert m.shape[1] == 4 and a.shape == b.shape and up.shape == (4,)


def test_hpatches_benchmark_loop_end_to_end(tmp_path, capsys):
    root = make_hpatches_tree(str(tmp_path))
    bm = HP.BatchedMatcher(_model(), "cuda:0", imsize=96, no_match_upscale=True, batch=8, depth=2)
    logged = []
    res = HP.eval_hpatches(bm, root, "GeoFormer_b200", task="both", scale_H=True, ransac_thres=3, lprint_=logged.append)
    assert res["pairs"] == 30 and res["match_failed"] == 0 and len(res["dists_sa"]) == 30
    assert res["n_matches"].min() >= 0 and res["n_matches"].max() > 20           # random-init weights still match at thr 0
    assert any(l.startswith(">>Finished, pairs=30 match_failed=0") for l in logged)
    assert "Hest AUC" in capsys.readouterr().out and 0.0 <= res["auc"] <= 1.0
