"""Property tests (hypothesis) of the host-side logic around the hot path: pair sharding, the wrappers' resize
arithmetic, the AUC metric and the vectorised anchor bookkeeping.  No GPU needed."""
import numpy as np
from hypothesis import given, settings, strategies as st

from geoformer_b200 import engine, evaluate
from geoformer_b200.dist import shard_pairs
from geoformer_b200.ingest import resize_dims


@given(st.integers(0, 500), st.integers(1, 16))
def test_shard_pairs_is_a_partition(n, world):
    shards = [shard_pairs(n, r, world) for r in range(world)]
    assert sorted(p for s in shards for p in s) == list(range(n))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1          # balanced within one pair


@given(st.integers(16, 4000), st.integers(16, 4000), st.sampled_from([None, 480, 640, 840]), st.sampled_from([1, 8, 16]))
def test_resize_dims_contract(wo, ho, imsize, dfactor):
    wt, ht, (sx, sy) = resize_dims(wo, ho, imsize=imsize, dfactor=dfactor, value_to_scale=min)
    if wt == 0 or ht == 0:
        return                                            # degenerate sliver: the reference divides by zero here too
    assert wt % dfactor == 0 and ht % dfactor == 0 and wt <= wo and ht <= ho
    assert abs(sx * wt - wo) < 1e-6 and abs(sy * ht - ho) < 1e-6
    if imsize is None or min(wo, ho) <= imsize:
        assert (wt, ht) == (wo // dfactor * dfactor, ho // dfactor * dfactor)          # only floored
    else:
        assert min(wt, ht) <= imsize and min(wt, ht) > imsize - dfactor - 1              # min side lands on imsize (then floored)


@given(st.lists(st.floats(0, 50, allow_nan=False), min_size=1, max_size=60))
def test_error_auc_bounds_and_monotonicity(errs):
    thr = [1, 3, 5, 10]
    a = evaluate.error_auc(np.array(errs), thr)
    assert ((a >= -1e-12) & (a <= 1 + 1e-12)).all()
    better = evaluate.error_auc(np.array(errs) * 0.5, thr)                # halving every error can only help
    assert (better >= a - 1e-12).all()
    assert np.allclose(evaluate.error_auc(np.zeros(len(errs)), thr), 1.0)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.lists(st.integers(0, 40), min_size=1, max_size=4))
def test_anchor_lists_are_sorted_unique_tokens_of_the_matches(seed, counts):
    """Samples with <= 8 matches take no RANSAC (geo_module.py:47): all their match tokens become anchors, in
    ascending order without duplicates, separately per image; larger samples give a subset."""
    rng = np.random.default_rng(seed)
    hw, scale = (12, 16), 8
    k0, k1 = [], []
    for m in counts:
        t0, t1 = rng.integers(0, hw[0] * hw[1], m), rng.integers(0, hw[0] * hw[1], m)
        k0.append(np.stack([(t0 % hw[1]) * scale, (t0 // hw[1]) * scale], 1))
        k1.append(np.stack([(t1 % hw[1]) * scale, (t1 // hw[1]) * scale], 1))
    k0 = np.concatenate(k0).astype(np.float32).reshape(-1, 2)
    k1 = np.concatenate(k1).astype(np.float32).reshape(-1, 2)
    hm, has_h, aidx, acnt = engine.geo_prepare_host(k0, k1, np.array(counts), hw, hw, scale, 8.0)
    offs = np.concatenate([[0], np.cumsum(counts)])
    for b, m in enumerate(counts):
        for side, kk in enumerate((k0, k1)):
            pts = kk[offs[b]:offs[b + 1]].astype(np.int64)
            toks = np.unique((pts[:, 1] // scale) * hw[1] + pts[:, 0] // scale)
            got = aidx[side, b, :acnt[side, b]]
            assert np.array_equal(got, np.unique(got)) and set(got.tolist()) <= set(toks.tolist())
            if not has_h[b]:
                assert np.array_equal(got, toks)
        if m <= 8:
            assert has_h[b] == 0 and not hm[:, b].any()
