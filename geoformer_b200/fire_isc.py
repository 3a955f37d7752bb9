"""The reference's two control-point homography benchmarks as throughput pipelines:

* FIRE retina registration — ``eval_FIRE.py:112-120`` -> ``eval_tool/immatch/utils/fire_helper.py::eval_fire`` (70-239),
* ISC-HE copy detection  — ``eval_ISC.py:126-135`` -> ``eval_tool/immatch/utils/my_helper.py::eval_homography_my`` (58-223).

Both loops are: matcher on a pair -> ``cv2.findHomography(RANSAC)`` -> (optionally) undo the wrapper's resize on the
estimate -> mean distance of the annotated control points -> AUC.  Here the matcher runs over many pairs at once
(``geoformer_b200.hpatches.BatchedMatcher``), the per-pair scoring runs in a thread pool and pairs shard over ranks;
the per-pair arithmetic and every logged / printed line follow the helpers (cited per line), including their quirks
(the "Eval hpatches" banner, 1e6 as the error of a failed estimate, FIRE's hard pair counts 71 / 48 / 14)."""
from __future__ import annotations

import os
from typing import Callable, Sequence, Tuple

import numpy as np

from . import evaluate
from .hpatches import run_pairs

# per-pair record: index, category (0 S, 1 P, 2 A, 3 other), #matches, avg control-point distance, inlier ratio,
# homography failed, match failed, inaccurate, match seconds
REC = 9
BIG = 1e6                                            # fire_helper.py:147, my_helper.py:136


def scale_homography(sw: float, sh: float) -> np.ndarray:
    return np.array([[sw, 0, 0], [0, sh, 0], [0, 0, 1]], dtype=float)


def _estimate(match_res, scale_H: bool, ransac_thres: float):
    """fire_helper.py:132-145 == my_helper.py:116-133: RANSAC homography, mapped back to the original image frames when
    the wrapper returned resized coordinates + its upscale vector.  (None, []) on any failure (the bare except)."""
    import cv2
    try:
        matches = match_res[0]
        H_pred, inliers = cv2.findHomography(matches[:, :2], matches[:, 2:4], cv2.RANSAC, ransac_thres)
        if scale_H:
            scale = match_res[4]
            H_pred = np.linalg.inv(scale_homography(1 / scale[2], 1 / scale[3])) @ H_pred @ scale_homography(1 / scale[0], 1 / scale[1])
        return H_pred, inliers
    except Exception:                                # noqa: BLE001
        return None, []


def _control_point_error(H_pred, raw: np.ndarray, dst: np.ndarray) -> Tuple[float, float, float]:
    """mean / max / median distance of the control points (fire_helper.py:174-181, my_helper.py:163-169)."""
    import cv2
    dst_pred = cv2.perspectiveTransform(raw.reshape(-1, 1, 2), H_pred).squeeze()
    dis = (dst - dst_pred) ** 2
    dis = np.sqrt(dis[:, 0] + dis[:, 1])
    return float(dis.mean()), float(dis.max()), float(np.median(dis))


def _score(index: int, category: int, match_res, scale_H, ransac_thres, load_points: Callable, mae_lim, mee_lim, secs) -> np.ndarray:
    rec = np.zeros(REC)
    rec[0], rec[1], rec[8] = index, category, secs
    failed = isinstance(match_res, Exception) or match_res is None
    rec[6] = 1.0 if failed else 0.0
    rec[2] = 0 if failed else len(match_res[0])
    H_pred, inliers = (None, []) if failed else _estimate(match_res, scale_H, ransac_thres)
    if H_pred is None:
        rec[3], rec[4], rec[5] = BIG, 0.0, 1.0
    else:
        raw, dst = load_points()
        rec[3], mae, mee = _control_point_error(H_pred, raw, dst)
        rec[4] = np.mean(inliers)
        rec[7] = 1.0 if (mae > mae_lim or mee > mee_lim) else 0.0
    return rec


def _finish(recs, n_local, wall, timed, h_failed_solver, ransac_thres, lprint_):
    """The lines both helpers log after the loop (fire_helper.py:222-229, my_helper.py:205-215)."""
    ok = recs[:, 6] == 0
    mt = float(np.mean(recs[ok, 8])) if (timed and ok.any()) else wall / max(1, n_local)
    lprint_(f">>Finished, pairs={int(ok.sum())} match_failed={int(recs[:, 6].sum())} matches={np.mean(recs[:, 2]):.1f} match_time={mt:.2f}s")
    return mt


def _tail(recs) -> str:
    n = len(recs)
    failed, inaccurate = int(recs[:, 5].sum()), int(recs[:, 7].sum())
    return (f"Failed:{'%.2f' % (100 * failed / n)}%, Inaccurate:{'%.2f' % (100 * inaccurate / n)}%, "
            f"Acceptable:{'%.2f' % (100 * (n - inaccurate - failed) / n)}%")


# ------------------------------------------------------------------------------------------------ FIRE
def fire_pair_paths(pair_file: str, im_dir: str) -> Tuple[str, str, str]:
    """control_points_<cat+id>_<refer>_<query>.txt -> (query image, reference image, category letter)
    (fire_helper.py:109-119): the matcher is called as matcher(query, refer)."""
    parts = pair_file.replace(".txt", "").split("_")
    refer, query = parts[2] + "_" + parts[3], parts[2] + "_" + parts[4]
    return os.path.join(im_dir, query + ".jpg"), os.path.join(im_dir, refer + ".jpg"), parts[2][0]


def compute_fire_auc(s_error, p_error, a_error, strict_counts: bool = True) -> dict:
    """fire_helper.py:11-42 (success-rate AUC for thresholds 1..25 px per category, mean of the three)."""
    if strict_counts:
        assert len(s_error) == 71 and len(p_error) == 48 and len(a_error) == 14      # fire_helper.py:12-14
    auc = {k: evaluate.fire_auc(np.asarray(e, dtype=float), 25) for k, e in (("s", s_error), ("p", p_error), ("a", a_error))}
    auc["mAUC"] = (auc["s"] + auc["p"] + auc["a"]) / 3.0
    return auc


def eval_fire(matcher, match_pairs: Sequence[str], im_dir: str, gt_dir: str, method: str = "", task: str = "homography",
              scale_H: bool = False, ransac_thres: float = 2, lprint_: Callable = print, debug: bool = False,
              rank: int = 0, world: int = 1, strict_counts: bool = True, score_threads: int = 8) -> dict:
    """``fire_helper.eval_fire`` with the matcher called on many pairs at once; returns the helper's mAUC under 'mAUC'."""
    np.set_printoptions(precision=4)
    assert task == "homography"
    lprint_(f"\n>>>>Eval hpatches: task={task} method={method} scale_H={scale_H} rthres={ransac_thres}")
    files = [f for i, f in enumerate(match_pairs) if not (debug and i > 10)]
    info = [fire_pair_paths(f, im_dir) for f in files]
    mine = list(range(rank, len(files), world))
    cat_code = {"S": 0, "P": 1, "A": 2}

    def points(f):
        def load():
            g = np.loadtxt(os.path.join(gt_dir, f))                          # fire_helper.py:153-159
            return np.ascontiguousarray(g[:, 2:4]), np.ascontiguousarray(g[:, 0:2])
        return load

    recs, wall, timed = run_pairs(
        matcher, [(info[i][0], info[i][1]) for i in mine],
        lambda k, res, secs: _score(mine[k], cat_code.get(info[mine[k]][2], 3), res, scale_H, ransac_thres,
                                    points(files[mine[k]]), 50, 20, secs),          # fire_helper.py:182-183
        REC, world, score_threads)
    recs = recs[np.argsort(recs[:, 0], kind="stable")]
    assert len(recs) == len(files) and np.array_equal(recs[:, 0], np.arange(len(files)))
    _finish(recs, len(mine), wall, timed, "cv", ransac_thres, lprint_)
    print("-" * 88); print("-" * 88)
    lprint_("==== Homography Estimation ====")
    lprint_(f"Hest solver=cv est_failed={int(recs[:, 5].sum())} ransac_thres={ransac_thres} inlier_rate={np.mean(recs[:, 4]):.2f}")
    d = {c: recs[recs[:, 1] == v, 3] for c, v in cat_code.items()}
    auc = compute_fire_auc(d["S"], d["P"], d["A"], strict_counts)
    summary = f'Hest AUC: m={auc["mAUC"]}\ns={auc["s"]}\np={auc["p"]}\na={auc["a"]}\n'
    print(summary)
    print("-" * 40); print(_tail(recs)); print("-" * 40)
    return dict(records=recs, summary=summary, tail=_tail(recs), dists_ss=d["S"], dists_sp=d["P"], dists_sa=d["A"],
                wall_s=wall, **auc)


# ------------------------------------------------------------------------------------------------ ISC-HE
def eval_homography_isc(matcher, match_pairs: Sequence[Tuple[str, str, str]], method: str = "", task: str = "homography",
                        scale_H: bool = False, ransac_thres: float = 2, lprint_: Callable = print, debug: bool = False,
                        rank: int = 0, world: int = 1, score_threads: int = 8) -> dict:
    """``my_helper.eval_homography_my``: match_pairs = (query image, reference image, control-point file) triples whose
    points are normalised to [0, 1] and scaled by the two image sizes (my_helper.py:142-152); AUC of the mean
    control-point distance at 3 / 5 / 10 px; returns the helper's value under 'auc'."""
    from PIL import Image
    np.set_printoptions(precision=4)
    assert task == "homography"
    lprint_(f"\n>>>>Eval hpatches: task={task} method={method} scale_H={scale_H} rthres={ransac_thres}")
    triples = [t for i, t in enumerate(match_pairs) if not (debug and i > 10)]
    mine = list(range(rank, len(triples), world))

    def points(t):
        def load():
            with Image.open(t[0]) as im:
                w1, h1 = im.size
            with Image.open(t[1]) as im:
                w2, h2 = im.size
            g = np.loadtxt(t[2])
            return g[:, 0:2] * [w1, h1], g[:, 2:4] * [w2, h2]
        return load

    recs, wall, timed = run_pairs(
        matcher, [(triples[i][0], triples[i][1]) for i in mine],
        lambda k, res, secs: _score(mine[k], 3, res, scale_H, ransac_thres, points(triples[mine[k]]), 10, 5, secs),   # my_helper.py:170
        REC, world, score_threads)
    recs = recs[np.argsort(recs[:, 0], kind="stable")]
    assert len(recs) == len(triples) and np.array_equal(recs[:, 0], np.arange(len(triples)))
    _finish(recs, len(mine), wall, timed, "cv", ransac_thres, lprint_)
    lprint_("==== Homography Estimation ====")
    lprint_(f"Hest solver=cv est_failed={int(recs[:, 5].sum())} ransac_thres={ransac_thres} inlier_rate={np.mean(recs[:, 4]):.2f}")
    auc_sa = evaluate.error_auc(recs[:, 3], [3, 5, 10])                         # my_helper.py:42-50,215
    summary = f"Hest AUC: a={auc_sa}\n\n"
    print(summary)
    print("-" * 40); print(_tail(recs)); print("-" * 40)
    return dict(records=recs, summary=summary, tail=_tail(recs), dists_all=recs[:, 3], auc_table=auc_sa, auc=float(auc_sa[-1]),
                wall_s=wall)


# ------------------------------------------------------------------------------------------------ command line
def fire_pairs(data_root: str) -> Tuple[list, str, str]:
    """Pair files and directories as eval_FIRE.py:27-31 lists them (P37_1_2 has no usable ground truth)."""
    gt_dir, im_dir = os.path.join(data_root, "Ground Truth"), os.path.join(data_root, "Images")
    files = [x for x in os.listdir(gt_dir) if x.endswith(".txt") and not x.endswith("P37_1_2.txt")]
    return files, im_dir, gt_dir


def isc_pairs(data_root: str) -> list:
    """(query, refer, ground truth) triples as eval_ISC.py:31-42 builds them."""
    q_dir, r_dir, g_dir = (os.path.join(data_root, d) for d in ("query", "refer", "gd"))
    out = []
    for i in os.listdir(q_dir):
        name = i.split("_")[0]
        out.append((os.path.join(q_dir, i), os.path.join(r_dir, name + "_1.jpg"), os.path.join(g_dir, name + "_2-" + name + "_1.txt")))
    return out


def main(argv=None):
    """``python -m geoformer_b200.fire_isc fire|isc --data_root <.../FIRE | .../ISC-HE> --ckpt saved_ckpt/geoformer.ckpt``
    (also under torchrun).  Defaults are the reference's yml entries: fire -> imsize 768, ransac 15; isc-he -> 480, 3;
    both with match threshold 0.2 and no_match_upscale (eval_configs/geoformer.yml)."""
    import argparse
    import copy
    import torch
    import torch.distributed as dist
    from . import synth
    from .hpatches import BatchedMatcher
    from .model.full_model import GeoFormer
    from .model.geo_config import default_cfg as geo_cfg
    from .model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    ap = argparse.ArgumentParser(description="Benchmark FIRE / ISC-HE (batched, multi-GPU)")
    ap.add_argument("benchmark", choices=["fire", "isc"])
    ap.add_argument("--data_root", required=True)
    ap.add_argument("--ckpt", default=None, help="reference checkpoint; default: synthetic weights (threshold forced to 0)")
    ap.add_argument("--match_threshold", type=float, default=0.2)
    ap.add_argument("--ransac_thres", type=float, default=None)
    ap.add_argument("--imsize", type=int, default=None)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--depth", type=int, default=3)
    a = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    thr = a.match_threshold if a.ckpt else 0.0
    conf, g = copy.deepcopy(default_cfg), dict(geo_cfg)
    conf["match_coarse"]["thr"] = thr
    g["coarse_thr"] = thr
    model = GeoFormer(conf, g)
    sd = torch.load(a.ckpt, map_location="cpu") if a.ckpt else synth.make_state_dict(0)
    model.load_state_dict(sd.get("state_dict", sd), strict=False)
    model = model.eval().to(device)
    imsize = a.imsize or (768 if a.benchmark == "fire" else 480)
    rthr = a.ransac_thres if a.ransac_thres is not None else (15 if a.benchmark == "fire" else 3)
    matcher = BatchedMatcher(model, device, imsize=imsize, no_match_upscale=True, batch=a.batch, depth=a.depth)
    say = print if rank == 0 else (lambda *_: None)
    if a.benchmark == "fire":
        files, im_dir, gt_dir = fire_pairs(a.data_root)
        res = eval_fire(matcher, files, im_dir, gt_dir, "GeoFormer_b200", scale_H=True, ransac_thres=rthr, lprint_=say,
                        rank=rank, world=world)
    else:
        res = eval_homography_isc(matcher, isc_pairs(a.data_root), "GeoFormer_b200", scale_H=True, ransac_thres=rthr,
                                  lprint_=say, rank=rank, world=world)
    say(f"{len(res['records'])} pairs in {res['wall_s']:.1f} s on {world} GPU(s) incl. decode, ingest and scoring")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
