"""geoformer_b200.fire_isc (FIRE and ISC-HE control-point benchmarks as batched / multi-rank loops) against what the
UNMODIFIED reference loops (fire_helper.eval_fire, my_helper.eval_homography_my) logged, printed and returned on the same
synthetic inputs with the same stand-in matchers (tests/golden/fire_isc_eval.json, make_golden.py --only-fire-isc)."""
import json
import os
import re
import socket

import numpy as np
import pytest
import torch

from geoformer_b200 import fire_isc as FI
from tests.util import make_fire_tree, make_isc_tree, stub_matcher_named

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fire_isc_eval.json")


@pytest.fixture(scope="module")
def golden():
    return json.load(open(GOLDEN))


@pytest.fixture(scope="module")
def trees(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("fire_isc"))
    return make_fire_tree(os.path.join(root, "fire")), make_isc_tree(os.path.join(root, "isc"))


def _strip(line):
    return re.sub(r" match_time=.*$", "", line)


class _ManyAtOnce:
    """match_many stand-in delivering the stub's results out of order (as shape buckets do)."""
    device = torch.device("cpu")

    def __init__(self, fn):
        self.fn = fn

    def match_many(self, pairs):
        order = list(range(len(pairs)))
        order = order[2::3] + order[0::3][::-1] + order[1::3]
        for k in order:
            try:
                yield k, self.fn(*pairs[k])
            except Exception as e:          # noqa: BLE001
                yield k, e


def _run(tag, trees, matcher_wrap, **kw):
    (files, im_dir, gt_dir), triples = trees
    scaled = tag.endswith("scaled")
    logged = []
    if tag.startswith("fire"):
        res = FI.eval_fire(matcher_wrap(stub_matcher_named("fire", scaled)), files, im_dir, gt_dir, "stub", scale_H=scaled,
                           ransac_thres=15, lprint_=logged.append, **kw)
        return res, logged, res["mAUC"]
    res = FI.eval_homography_isc(matcher_wrap(stub_matcher_named("isc", scaled)), triples, "stub", scale_H=scaled,
                                 ransac_thres=3, lprint_=logged.append, **kw)
    return res, logged, res["auc"]


def _check(tag, g, res, logged, value, stdout=None):
    assert [_strip(l) for l in logged] == [_strip(l) for l in g["logged"]]
    assert value == g["value"]
    if tag.startswith("fire"):
        for k in ("dists_ss", "dists_sp", "dists_sa"):
            np.testing.assert_array_equal(np.asarray(g[k]), res[k])
    else:
        np.testing.assert_array_equal(np.asarray(g["dists_all"]), res["dists_all"])
    if stdout is not None:              # everything the helper prints: failure text, dashes, AUC table, failed / inaccurate line
        assert stdout.strip().splitlines() == g["stdout"].strip().splitlines()


@pytest.mark.parametrize("tag", ["fire_scaled", "fire_plain", "isc_scaled", "isc_plain"])
def test_serial_matcher_reproduces_reference_loops(tag, trees, golden, capsys):
    res, logged, value = _run(tag, trees, lambda fn: fn)
    _check(tag, golden[tag], res, logged, value, capsys.readouterr().out)


@pytest.mark.parametrize("tag", ["fire_scaled", "isc_plain"])
def test_batched_interface_reproduces_reference_loops(tag, trees, golden, capsys):
    res, logged, value = _run(tag, trees, _ManyAtOnce)
    _check(tag, golden[tag], res, logged, value, capsys.readouterr().out)


def test_fire_pair_naming_and_counts():
    q, r, cat = FI.fire_pair_paths("control_points_P37_1_2.txt", "/d")
    assert (q, r, cat) == ("/d/P37_2.jpg", "/d/P37_1.jpg", "P")            # matcher(query, refer), fire_helper.py:113-116
    with pytest.raises(AssertionError):                                     # the helper's hard pair counts
        FI.compute_fire_auc([1.0] * 70, [1.0] * 48, [1.0] * 14)
    assert FI.compute_fire_auc([0.5, 30.0], [2.5], [], strict_counts=False)["s"] == pytest.approx((25 * 50) / 2500)


def _rank_worker(rank, world, port, root, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    trees = (make_fire_tree(os.path.join(root, "fire")), None)
    res, logged, value = _run("fire_scaled", trees, _ManyAtOnce, rank=rank, world=world)
    q.put((rank, logged, value, res["dists_ss"].tolist(), res["tail"]))
    dist.barrier()
    dist.destroy_process_group()


def test_fire_on_two_ranks(tmp_path, golden):
    import torch.multiprocessing as mp
    g = golden["fire_scaled"]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, logged, value, dss, tail in out:
        assert [_strip(l) for l in logged] == [_strip(l) for l in g["logged"]] and value == g["value"]
        np.testing.assert_array_equal(np.asarray(dss), np.asarray(g["dists_ss"]))
        assert tail in g["stdout"]
