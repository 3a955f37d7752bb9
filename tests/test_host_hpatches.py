"""geoformer_b200.hpatches (the batched / multi-rank restatement of the reference's HPatches benchmark loop,
eval_tool/immatch/utils/hpatches_helper.py:94-317) against what the UNMODIFIED reference loop logged, printed and
handed to its summary functions on the same synthetic tree with the same stand-in matcher
(tests/golden/hpatches_eval.json, written by tests/golden/make_golden.py --only-hpatches).  CPU only."""
import copy
import json
import os
import re
import socket

import numpy as np
import pytest
import torch

from geoformer_b200 import hpatches as HP
from tests.util import make_hpatches_tree, stub_matcher


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    return make_hpatches_tree(str(tmp_path_factory.mktemp("hpatches")))


@pytest.fixture(scope="module")
def golden(golden_dir_module):
    return json.load(open(os.path.join(golden_dir_module, "hpatches_eval.json")))


@pytest.fixture(scope="module")
def golden_dir_module():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _strip_time(line):
    return re.sub(r" match_time=.*$", "", line)


def _check(res, logged, stdout, g):
    want_logged = [_strip_time(l) for l in g["logged"]]
    assert [_strip_time(l) for l in logged] == want_logged
    # the helper prints the matcher's exception text and the homography table to stdout
    assert stdout.strip().splitlines()[-7:] == g["stdout"].strip().splitlines()[-7:]
    if "i_err" in g:
        assert {int(k): v for k, v in g["i_err"].items()} == pytest.approx(res["i_err"], rel=0, abs=0)
        assert {int(k): v for k, v in g["v_err"].items()} == pytest.approx(res["v_err"], rel=0, abs=0)
        assert g["n_matches"] == [int(x) for x in res["n_matches"]]
        assert g["n_feats"] == [int(x) for x in res["records"][:, 3:5].reshape(-1)]
    np.testing.assert_array_equal(np.asarray(g["dists_sa"]), res["dists_sa"])          # incl. the NaNs of failed estimates
    np.testing.assert_array_equal(np.asarray(g["dists_si"]), res["dists_si"])
    np.testing.assert_array_equal(np.asarray(g["dists_sv"]), res["dists_sv"])
    assert res["auc"] == g["auc"]


@pytest.mark.parametrize("tag", ["scaled_both", "plain_both", "scaled_homography"])
def test_serial_matcher_reproduces_reference_loop(tree, golden, capsys, tag):
    g = golden[tag]
    logged = []
    res = HP.eval_hpatches(stub_matcher(g["scaled"]), tree, "stub", task=g["task"], scale_H=g["scaled"],
                           ransac_thres=g["ransac_thres"], lprint_=logged.append)
    _check(res, logged, capsys.readouterr().out, g)
    assert res["pairs"] == 30 and res["match_failed"] == 1 and res["h_failed"] == 3


class _ManyAtOnce:
    """match_many stand-in: the stub's per-pair results, delivered out of order (as shape buckets are)."""
    device = torch.device("cpu")

    def __init__(self, scaled):
        self.fn = stub_matcher(scaled)

    def match_many(self, pairs):
        order = list(range(len(pairs)))[::-1]
        order = order[1::2] + order[0::2]
        for k in order:
            try:
                yield k, self.fn(*pairs[k])
            except Exception as e:          # noqa: BLE001
                yield k, e


def test_batched_matcher_interface_reproduces_reference_loop(tree, golden, capsys):
    g = golden["scaled_both"]
    logged = []
    res = HP.eval_hpatches(_ManyAtOnce(True), tree, "stub", task="both", scale_H=True, ransac_thres=3, lprint_=logged.append)
    _check(res, logged, capsys.readouterr().out, g)


def test_list_pairs_order_and_debug(tree):
    pairs = HP.list_pairs(tree)
    assert [p.seq for p in pairs[::5]] == ["v_circus", "v_boat", "v_bird", "i_dome", "i_castle", "i_ajuntament"]     # reversed sort
    assert [p.im_idx for p in pairs[:5]] == [2, 3, 4, 5, 6] and pairs[7].index == 7
    assert pairs[0].H_gt.shape == (3, 3) and pairs[0].im1.endswith("v_circus/1.ppm")


def _rank_worker(rank, world, port, root, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    logged = []
    res = HP.eval_hpatches(_ManyAtOnce(True), root, "stub", task="both", scale_H=True, ransac_thres=3, lprint_=logged.append,
                           rank=rank, world=world)
    q.put((rank, logged, res["i_err"], res["dists_sa"].tolist(), res["auc"], res["summary_homography"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_print_the_single_process_tables(tree, golden):
    """Pairs shard p -> rank p mod 2, records are all-gathered (gloo): both ranks end with the reference's numbers."""
    import torch.multiprocessing as mp
    g = golden["scaled_both"]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, tree, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, logged, i_err, dists, auc, table in out:
        assert [_strip_time(l) for l in logged] == [_strip_time(l) for l in g["logged"]]
        assert i_err == {int(k): v for k, v in g["i_err"].items()}
        np.testing.assert_array_equal(np.asarray(dists), np.asarray(g["dists_sa"]))
        assert auc == g["auc"] and table.strip().splitlines() == g["stdout"].strip().splitlines()[-6:]


def test_batched_matcher_groups_by_shape_and_matches_the_per_pair_wrapper(monkeypatch, tmp_path):
    """BatchedMatcher on CPU (operators emulated, serial runner): decode -> buckets of equal resized shape -> batched
    ingest + forward -> per-pair wrapper tuples, against the reference wrapper's per-pair sequence (cv2.imread,
    resize_im, cv2.resize, batch-1 forward; geoformer.py:43-99) on the same model."""
    import cv2
    from geoformer_b200 import synth
    from geoformer_b200.ingest import resize_dims
    from geoformer_b200.model.full_model import GeoFormer
    from geoformer_b200.model.geo_config import default_cfg as geo_cfg
    from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    from tests import emu_ops as emu
    emu.install(monkeypatch)
    root = make_hpatches_tree(str(tmp_path), seqs=(("i_a", (120, 160)), ("v_b", (128, 96)), ("v_c", (96, 128))), seed=5)
    g = dict(geo_cfg); g["coarse_thr"] = 0.0
    model = GeoFormer(copy.deepcopy(default_cfg), g)
    model.load_state_dict(synth.make_state_dict(7, True), strict=True)
    model = model.eval()
    monkeypatch.setattr(GeoFormer, "forward", lambda self, data: self._forward(data, data["image0"], data["image1"]))
    shapes = []

    def runner(batches, prepare, post):
        guarded = HP._Guarded(model)
        for desc in batches:
            shapes.append((desc["_shape"], len(desc["_items"])))
            with torch.no_grad():
                yield post(guarded(prepare(desc)))

    pairs = [(p.im1, p.im2) for p in HP.list_pairs(root)]
    pairs.append((pairs[0][0], os.path.join(root, "missing.ppm")))                  # undecodable -> per-pair failure
    bm = HP.BatchedMatcher(model, "cpu", imsize=96, no_match_upscale=True, batch=4, runner=runner)
    got = dict(bm.match_many(pairs))
    assert sorted(got) == list(range(16)) and isinstance(got[15], FileNotFoundError)
    # 120x160 -> 96x128 joins the 96x128 sequence: 10 pairs in batches of 4 + 4 + 2; the 128x96 sequence rides alone
    assert sorted(shapes) == sorted([(((96, 128), (96, 128)), 4), (((96, 128), (96, 128)), 4), (((96, 128), (96, 128)), 2),
                                     (((128, 96), (128, 96)), 4), (((128, 96), (128, 96)), 1)])
    same = total = 0
    for k in (0, 3, 7, 12):                                                          # per-pair wrapper sequence
        ims, scs = [], []
        for p in pairs[k]:
            im = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
            wt, ht, sc = resize_dims(im.shape[1], im.shape[0], imsize=96, dfactor=8, value_to_scale=min)
            ims.append(torch.from_numpy(cv2.resize(im, (wt, ht))).float().div(255)[None, None]); scs.append(sc)
        with torch.no_grad():
            d = model({"image0": ims[0], "image1": ims[1]})
        want = np.concatenate([d["mkpts0_f"].numpy(), d["mkpts1_f"].numpy()], 1)
        matches, k1, k2, scores, upscale = got[k]
        assert np.allclose(upscale, np.array(scs[0] + scs[1])) and matches.shape[1] == 4 and len(scores) == len(matches)
        assert np.array_equal(matches[:, :2], k1) and np.array_equal(matches[:, 2:], k2)
        a, b = {tuple(r) for r in want.tolist()}, {tuple(r) for r in matches.tolist()}
        same += len(a & b); total += len(a | b)
    assert total > 100 and same >= 0.97 * total, (same, total)       # batch-1 vs batched: identical up to fp32 round-off flips
    # the other wrapper convention (no_match_upscale False): coordinates scaled back to the original image
    bm2 = HP.BatchedMatcher(model, "cpu", imsize=96, no_match_upscale=False, batch=4, runner=runner)
    m2, a2, b2, s2 = bm2(*pairs[0])
    m1 = got[0]
    assert np.allclose(m2, m1[0] * m1[4][None]) and np.allclose(a2, m1[1] * m1[4][:2]) and np.allclose(b2, m1[2] * m1[4][2:])


def test_failed_batches_become_per_pair_failures(tmp_path):
    """A batch whose ingest or forward raises is reported as one failure per pair (the helper's try / except counts
    match_failed per pair, hpatches_helper.py:193-196); the other batches are unaffected."""
    root = make_hpatches_tree(str(tmp_path), seqs=(("i_a", (96, 128)), ("v_b", (128, 96))), seed=2)

    class Model:
        def __call__(self, data):
            if data["image0"].shape[-1] == 96:                       # the 128x96 bucket blows up in the forward
                raise RuntimeError("forward failed")
            n = data["image0"].shape[0]
            data.update(mkpts0_f=torch.zeros(n, 2), mkpts1_f=torch.ones(n, 2), mconf=torch.full((n,), 0.5), m_bids=torch.arange(n))
            return data

    def runner(batches, prepare, post):
        g = HP._Guarded(Model())
        for k, desc in enumerate(batches):
            yield post(g(prepare(desc) if k != 0 else {"_items": desc["_items"], "_error": MemoryError("ingest failed")}))

    import cv2
    monkey_resize = lambda src, dst: dst.copy_(torch.from_numpy(cv2.resize(src.numpy(), (dst.shape[1], dst.shape[0]))).float().div(255))
    from geoformer_b200 import ops
    orig = ops.resize_gray_u8
    ops.resize_gray_u8 = monkey_resize
    try:
        bm = HP.BatchedMatcher(None, "cpu", imsize=96, batch=2, runner=runner)
        got = dict(bm.match_many([(p.im1, p.im2) for p in HP.list_pairs(root)]))
    finally:
        ops.resize_gray_u8 = orig
    kinds = [type(got[k]).__name__ if isinstance(got[k], Exception) else "ok" for k in range(10)]
    # list order: 5 pairs of v_b (128 x 96) then 5 of i_a; batches of 2: the first v_b batch fails in the ingest, the other
    # v_b batches (2 + the flushed 1) in the forward, every i_a batch succeeds
    assert kinds[:5].count("MemoryError") == 2 and kinds[:5].count("RuntimeError") == 3 and kinds[5:] == ["ok"] * 5
    ok = [got[k] for k in range(10) if not isinstance(got[k], Exception)]
    assert all(m[0].shape == (1, 4) and m[4].shape == (4,) for m in ok)
