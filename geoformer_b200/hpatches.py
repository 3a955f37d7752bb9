"""HPatches benchmark as a throughput pipeline — the caller of the hot path in the reference's headline evaluation
(``eval_Hpatches.py:106`` -> ``eval_tool/immatch/utils/hpatches_helper.py:94-317`` -> wrapper
``eval_tool/immatch/modules/geoformer.py:77-99``), SURVEY.md §3.1 / §8f.

The reference walks the 580 pairs serially: read + resize two images on the host, one batch-1 forward, a blocking
device->host copy, ``cv2.findHomography``, corner error.  Here the same per-pair arithmetic is organised for a GPU that
matches > 400 pairs per second:

* ``BatchedMatcher`` — ``match_pairs`` for many pairs: images are decoded by a host thread pool, pairs whose RESIZED
  shapes agree are grouped into batches (HPatches at imsize 480 is mostly 640x480), each batch is uploaded as uint8 and
  resized on the GPU bit-exactly to ``cv2.resize`` (``geoformer_b200.ingest``), run through ``MatchPipeline`` (several
  batches in flight) and split back into the wrapper's per-pair tuples
  ``(matches, kpts1, kpts2, scores[, upscale])`` (geoformer.py:88-99).
* ``eval_hpatches`` — the helper's loop and its two summaries (mean matching accuracy at 1..15 px, homography
  accuracy / AUC of the corner error) with the per-pair ``cv2.findHomography`` calls in a thread pool; pairs shard over
  ranks (``p -> rank p mod world``) and the per-pair records are all-gathered, so every rank prints the same tables.

Metric arithmetic follows the helper line by line (cited below) including its quirks: sequences are visited in reverse
order, every sequence whose name does not start with 'i' counts as viewpoint, and the matching table is normalised by
the fixed counts of the full release (52 illumination / 56 viewpoint sequences x 5 pairs, hpatches_helper.py:59-60).
"""
from __future__ import annotations

import glob
import os
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import evaluate
from .ingest import resize_dims

THRES_RANGE = np.arange(1, 16)                       # hpatches_helper.py:141
N_I, N_V = 52, 56                                    # hpatches_helper.py:59-60 (sequence counts of the full release)
# per-pair record: index, type (0 = 'i', 1 = other), #matches, #feats0, #feats1, 15 x MMA fraction, corner distance,
# inlier ratio, homography failed, match failed, match seconds
REC = 2 + 3 + len(THRES_RANGE) + 5


@dataclass
class Pair:
    index: int
    seq: str
    im_idx: int
    im1: str
    im2: str
    H_gt: np.ndarray


def list_pairs(data_root: str, debug: bool = False) -> List[Pair]:
    """Pairs in the helper's order (hpatches_helper.py:136,161-172): sequence directories sorted then REVERSED, images
    2..6 against image 1, ground truth from H_1_k; ``debug`` keeps the first 11 sequences."""
    pairs: List[Pair] = []
    seq_dirs = sorted(glob.glob("{}/*".format(data_root)))
    for seq_idx, seq_dir in enumerate(seq_dirs[::-1]):
        if debug and seq_idx > 10:
            break
        sname = seq_dir.split("/")[-1]
        for im_idx in range(2, 7):
            pairs.append(Pair(len(pairs), sname, im_idx, os.path.join(seq_dir, "1.ppm"),
                              os.path.join(seq_dir, "{}.ppm".format(im_idx)),
                              np.loadtxt(os.path.join(seq_dir, "H_1_{}".format(im_idx)))))
    return pairs


# ------------------------------------------------------------------------------------------------ matching many pairs
class _Guarded:
    """model(data) that reports a failed batch instead of raising (the helper counts failures per pair,
    hpatches_helper.py:173-196) - what MatchPipeline needs of a model: __call__ and _weights."""

    def __init__(self, model):
        self.model = model

    def _weights(self, device):
        return self.model._weights(device)

    def __call__(self, data):
        if "_error" in data:            # the ingest of this batch already failed
            return data
        try:
            return self.model(data)
        except Exception as e:          # noqa: BLE001 - surfaced per pair by BatchedMatcher
            data["_error"] = e
            return data


class BatchedMatcher:
    """``GeoFormer.match_pairs`` (geoformer.py:77-99) over many pairs at once.

    model: a ``geoformer_b200.model.full_model.GeoFormer`` on ``device``; imsize / no_match_upscale as the wrapper's
    yml entries (eval_configs/geoformer.yml: hpatch -> 480 / True); batch: pairs per forward; depth: batches in flight.
    runner(batches, prepare, post) -> iterator of post() results: defaults to a MatchPipeline on ``device``; tests
    inject a serial runner so that grouping / splitting run without a GPU."""

    def __init__(self, model, device, imsize: int = 480, no_match_upscale: bool = True, batch: int = 16, depth: int = 3,
                 dfactor: int = 8, decode_threads: int = 8, runner: Optional[Callable] = None):
        self.model, self.device = model, torch.device(device)
        self.imsize, self.no_match_upscale, self.batch, self.depth, self.dfactor = imsize, no_match_upscale, batch, depth, dfactor
        self.decode_threads = decode_threads
        self.runner = runner or self._pipeline_runner
        self._pipe = None

    # -- host side: decode + grouping ----------------------------------------------------------------
    def _decode(self, item):
        import cv2
        k, (p1, p2) = item
        try:
            ims = []
            for p in (p1, p2):
                im = cv2.imread(p, cv2.IMREAD_GRAYSCALE)                 # data_io.py:51
                if im is None:
                    raise FileNotFoundError(p)
                ho, wo = im.shape
                wt, ht, scale = resize_dims(wo, ho, imsize=self.imsize, dfactor=self.dfactor, value_to_scale=min)
                ims.append((im, (ht, wt), scale))
            return k, ims, None
        except Exception as e:          # noqa: BLE001
            return k, None, e

    def _batches(self, pairs: Sequence[Tuple[str, str]], failed: list) -> Iterator[dict]:
        """Batch descriptors of pairs with equal resized shapes, emitted as soon as `batch` of them are decoded.  Decoding
        runs ahead of the consumer in a producer thread (bounded queue), so the GPU is not idle while images are read."""
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=self.depth + 2)
        END = object()

        def produce():
            buckets: Dict[tuple, list] = {}
            try:
                with ThreadPoolExecutor(max_workers=self.decode_threads) as ex:
                    window = 2 * self.batch                                # decoded images in flight beyond the queue
                    items = list(enumerate(pairs))
                    for lo in range(0, len(items), window):
                        for k, ims, err in ex.map(self._decode, items[lo:lo + window]):
                            if err is not None:
                                failed.append((k, err))
                                continue
                            key = (ims[0][1], ims[1][1])
                            buckets.setdefault(key, []).append((k, ims))
                            if len(buckets[key]) == self.batch:
                                q.put({"_shape": key, "_items": buckets.pop(key)})
                for key in sorted(buckets):
                    q.put({"_shape": key, "_items": buckets[key]})
                q.put(END)
            except BaseException as e:      # noqa: BLE001 - re-raised in the consumer
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        while True:
            item = q.get()
            if item is END:
                break
            if isinstance(item, BaseException):
                raise item
            yield item
        t.join()

    # -- device side (runs on the batch's stream inside the pipeline) --------------------------------------
    def _ingest(self, desc: dict) -> dict:
        """uint8 upload + cv2-exact GPU resize of every image of the batch into [n,1,H,W] tensors (data_io.py:48-62)."""
        from . import ops
        (hw0, hw1), items = desc["_shape"], desc["_items"]
        n = len(items)
        try:
            im0 = torch.empty((n, 1) + tuple(hw0), device=self.device, dtype=torch.float32)
            im1 = torch.empty((n, 1) + tuple(hw1), device=self.device, dtype=torch.float32)
            for b, (_, ims) in enumerate(items):
                for dst, (raw, _, _) in ((im0[b, 0], ims[0]), (im1[b, 0], ims[1])):
                    host = torch.from_numpy(np.ascontiguousarray(raw))
                    src = (host.pin_memory() if self.device.type == "cuda" else host).to(self.device, non_blocking=True)
                    ops.resize_gray_u8(src, dst)
        except Exception as e:          # noqa: BLE001 - a failed batch is a per-pair failure, not the end of the run
            return {"_items": items, "_error": e}
        return {"image0": im0, "image1": im1, "_items": items}

    def _post(self, data: dict) -> list:
        """Per-pair wrapper tuples from one batch (geoformer.py:50-54,73,85-99)."""
        items = data["_items"]
        if "_error" in data:
            return [(k, data["_error"]) for k, _ in items]
        k0 = data["mkpts0_f"].cpu().numpy()
        k1 = data["mkpts1_f"].cpu().numpy()
        sc = data["mconf"].cpu().numpy()
        mb = data["m_bids"].cpu().numpy()
        out = []
        for b, (k, ims) in enumerate(items):
            sel = mb == b
            kpts1, kpts2, scores = k0[sel], k1[sel], sc[sel]
            sc1, sc2 = ims[0][2], ims[1][2]
            upscale = np.array([sc1 + sc2])                              # [[sx1, sy1, sx2, sy2]]
            matches = np.concatenate([kpts1, kpts2], axis=1)
            if self.no_match_upscale:
                out.append((k, (matches, kpts1, kpts2, scores, upscale.squeeze(0))))
            else:
                out.append((k, (upscale * matches, sc1 * kpts1, sc2 * kpts2, scores)))
        return out

    def _pipeline_runner(self, batches, prepare, post):
        from .pipeline import MatchPipeline
        if self._pipe is None:
            self._pipe = MatchPipeline(_Guarded(self.model), depth=self.depth, device=self.device, prepare=prepare)
        return self._pipe.run_iter(batches, post)

    def match_many(self, pairs: Sequence[Tuple[str, str]]) -> Iterator[Tuple[int, object]]:
        """Yields (position in `pairs`, wrapper tuple | Exception) for every pair, in completion order."""
        failed: list = []
        for res in self.runner(self._batches(pairs, failed), self._ingest, self._post):
            yield from res
        yield from failed

    def __call__(self, im1_path: str, im2_path: str):
        """Single-pair form with the wrapper's signature (raises like the wrapper on failure)."""
        (_, res), = list(self.match_many([(im1_path, im2_path)]))
        if isinstance(res, Exception):
            raise res
        return res


# ------------------------------------------------------------------------------------------------ scoring one pair
def scale_homography(sw: float, sh: float) -> np.ndarray:
    return np.array([[sw, 0, 0], [0, sh, 0], [0, 0, 1]], dtype=float)       # hpatches_helper.py:89-92


def _image_size(path: str) -> Tuple[int, int]:
    from PIL import Image
    with Image.open(path) as im:                                             # hpatches_helper.py:227-228
        return im.size


def score_pair(pair: Pair, match_res, task: str, scale_H: bool, ransac_thres: float, seconds: float = 0.0) -> np.ndarray:
    """One iteration of the helper's inner loop (hpatches_helper.py:166-240) as a record of REC floats."""
    import cv2
    rec = np.zeros(REC)
    rec[0], rec[1] = pair.index, 0.0 if pair.seq[0] == "i" else 1.0
    H_gt = pair.H_gt
    scale = np.ones(4)
    failed = isinstance(match_res, Exception) or match_res is None
    if failed:
        matches, p1s, p2s = [], [], []                                       # :193-196
    else:
        matches, p1s, p2s = match_res[0:3]
        if scale_H:                                                          # :185-192
            scale = match_res[4]
            H_gt = np.linalg.inv(scale_homography(scale[2], scale[3])) @ H_gt @ scale_homography(scale[0], scale[1])
    rec[2], rec[3], rec[4] = len(matches), len(p1s), len(p2s)
    if "matching" in task:                                                   # :199-211
        dist = np.array([float("inf")]) if len(matches) == 0 else evaluate.reproj_dists(matches[:, :2], matches[:, 2:], H_gt)
        rec[5:5 + len(THRES_RANGE)] = [np.mean(dist <= thr) for thr in THRES_RANGE]
    o = 5 + len(THRES_RANGE)
    if "homography" in task:                                                 # :213-240
        try:
            H_pred, inliers = cv2.findHomography(matches[:, :2], matches[:, 2:4], cv2.RANSAC, ransac_thres)
        except Exception:                                                    # noqa: BLE001 - the helper's bare except
            H_pred = None
        if H_pred is None:
            rec[o], rec[o + 1], rec[o + 2] = np.nan, 0.0, 1.0
        else:
            w, h = _image_size(pair.im1)
            w, h = w / scale[0], h / scale[1]
            corners = np.array([[0, 0, 1], [0, h - 1, 1], [w - 1, 0, 1], [w - 1, h - 1, 1]])
            real = np.dot(corners, np.transpose(H_gt))
            real = real[:, :2] / real[:, 2:]
            warped = np.dot(corners, np.transpose(H_pred))
            warped = warped[:, :2] / warped[:, 2:]
            rec[o] = np.mean(np.linalg.norm(real - warped, axis=1))
            rec[o + 1] = np.mean(inliers)
    rec[o + 3] = 1.0 if failed else 0.0
    rec[o + 4] = seconds
    return rec


# ------------------------------------------------------------------------------------------------ summaries
def summary_matching(recs: np.ndarray, thres: Sequence[float]) -> str:
    """hpatches_helper.py:55-87 on the gathered records (rows sorted by pair index)."""
    np.set_printoptions(precision=4)
    seq_type = np.where(recs[:, 1] == 0, "i", "v")
    n_matches = recs[:, 2]
    n_feats = recs[:, 3:5].reshape(-1)                                        # len(p1s), len(p2s) per pair, interleaved
    # the helper appends sname[0] to seq_type but splits sums by == 'i' / == 'v' and the accuracies by 'i' / not 'i';
    # names outside {i, v} do not occur in the dataset
    s = "#Features: mean={:.0f} min={:d} max={:d}\n".format(np.mean(n_feats), int(np.min(n_feats)), int(np.max(n_feats)))
    s += "#(Old)Matches: a={:.0f}, i={:.0f}, v={:.0f}\n".format(np.sum(n_matches) / ((N_I + N_V) * 5),
                                                               np.sum(n_matches[seq_type == "i"]) / (N_I * 5),
                                                               np.sum(n_matches[seq_type == "v"]) / (N_V * 5))
    s += "#Matches: a={:.0f}, i={:.0f}, v={:.0f}\n".format(np.mean(n_matches), np.mean(n_matches[seq_type == "i"]),
                                                          np.mean(n_matches[seq_type == "v"]))
    i_err, v_err = matching_sums(recs)
    thres = np.array(thres)
    ierr = np.array([i_err[th] / (N_I * 5) for th in thres])
    verr = np.array([v_err[th] / (N_V * 5) for th in thres])
    aerr = np.array([(i_err[th] + v_err[th]) / ((N_I + N_V) * 5) for th in thres])
    s += "MMA@{} px:\na={}\ni={}\nv={}\n".format(thres, aerr, ierr, verr)
    return s


def matching_sums(recs: np.ndarray) -> Tuple[Dict[int, float], Dict[int, float]]:
    """i_err / v_err of the helper (:142-143, 207-211): per-threshold sums of the per-pair accuracies.  Accumulated in
    pair order, as the helper's running `+=` does."""
    i_err = {int(t): 0 for t in THRES_RANGE}
    v_err = {int(t): 0 for t in THRES_RANGE}
    for r in recs:
        tgt = i_err if r[1] == 0 else v_err
        for c, t in enumerate(THRES_RANGE):
            tgt[int(t)] += r[5 + c]
    return i_err, v_err


def summary_homography(recs: np.ndarray, thres: Sequence[float]) -> Tuple[str, float, dict]:
    """hpatches_helper.py:38-53: accuracy and AUC of the corner distance for all / illumination / viewpoint pairs."""
    o = 5 + len(THRES_RANGE)
    d_a = recs[:, o]
    d_i, d_v = d_a[recs[:, 1] == 0], d_a[recs[:, 1] == 1]
    correct = lambda d: np.mean([[float(x <= t) for t in thres] for x in d], axis=0)
    c_a, c_i, c_v = correct(d_a), correct(d_i), correct(d_v)
    a_a, a_i, a_v = (evaluate.error_auc(d, thres) for d in (d_a, d_i, d_v))
    s = f"Hest Correct: a={c_a}\ni={c_i}\nv={c_v}\n"
    s += f"Hest AUC: a={a_a}\ni={a_i}\nv={a_v}\n"
    return s, float(a_a[-1]), dict(correct_a=c_a, correct_i=c_i, correct_v=c_v, auc_a=a_a, auc_i=a_i, auc_v=a_v)


# ------------------------------------------------------------------------------------------------ the benchmark loop
def run_pairs(matcher, paths: Sequence[Tuple[str, str]], score: Callable, rec_len: int, world: int = 1,
              score_threads: int = 8) -> Tuple[np.ndarray, float, bool]:
    """The part every benchmark loop shares: run `matcher` over this rank's pairs — ``match_many`` when it has one
    (``BatchedMatcher``), else the reference's serial ``matcher(im1_path, im2_path)`` calls with their try / except
    (hpatches_helper.py:173-196) — score every result with ``score(position, result | Exception, seconds)`` in a thread
    pool (cv2.findHomography releases the GIL) and, with world > 1, all-gather the records of all ranks.
    Returns (records [pairs, rec_len], wall seconds of this rank, whether per-pair times exist)."""
    t_start = time.time()
    if hasattr(matcher, "match_many"):
        stream: Iterable = matcher.match_many(list(paths))
        timed = False
    else:
        def serial():
            for k, (p1, p2) in enumerate(paths):
                t0 = time.time()
                try:
                    yield k, (matcher(p1, p2), time.time() - t0)
                except Exception as e:      # noqa: BLE001
                    print(str(e))
                    yield k, (e, time.time() - t0)
        stream, timed = serial(), True
    futures = []
    with ThreadPoolExecutor(max_workers=score_threads) as ex:
        for k, res in stream:
            secs = 0.0
            if timed:
                res, secs = res
            if isinstance(res, Exception) and not timed:
                print(str(res))
            futures.append(ex.submit(score, k, res, secs))
        recs = np.stack([f.result() for f in futures]) if futures else np.zeros((0, rec_len))
    wall = time.time() - t_start
    if world > 1:
        from .dist import all_gather_rows
        dev = getattr(matcher, "device", torch.device("cpu"))
        recs = all_gather_rows(torch.from_numpy(recs).to(dev)).cpu().numpy()
    return recs, wall, timed


def eval_hpatches(matcher, data_root: str, method: str = "", task: str = "both", scale_H: bool = False,
                  ransac_thres: float = 2, thres: Sequence[float] = (1, 3, 5, 10), lprint_: Callable = print,
                  debug: bool = False, rank: int = 0, world: int = 1, score_threads: int = 8) -> dict:
    """``helper.eval_hpatches`` (hpatches_helper.py:94-317) with the matcher called on many pairs at once.

    matcher: an object with ``match_many(pairs) -> iterator of (position, wrapper tuple | Exception)``
    (``BatchedMatcher``), or a plain ``matcher(im1_path, im2_path)`` callable as the reference takes (called serially).
    h_solver is OpenCV ('cv', the eval script's default, eval_Hpatches.py:93-96).  With world > 1 every rank scores
    ``pairs[rank::world]`` and the records are all-gathered (torch.distributed must be initialised)."""
    np.set_printoptions(precision=4)                                         # hpatches_helper.py:129 (global, as there)
    if task == "both":
        task = "matching+homography"
    thres = list(thres)
    pairs = list_pairs(data_root, debug)
    lprint_(f"\n>>>>Eval hpatches: task={task} method={method} scale_H={scale_H} rthres={ransac_thres} thres={thres} ")
    mine = [pairs[i] for i in range(rank, len(pairs), world)]
    recs, wall, timed = run_pairs(matcher, [(p.im1, p.im2) for p in mine],
                                  lambda k, res, secs: score_pair(mine[k], res, task, scale_H, ransac_thres, secs),
                                  REC, world, score_threads)
    recs = recs[np.argsort(recs[:, 0], kind="stable")]
    assert len(recs) == len(pairs) and np.array_equal(recs[:, 0], np.arange(len(pairs))), "every pair exactly once"
    o = 5 + len(THRES_RANGE)
    n_matches = recs[:, 2]
    match_failed = int(recs[:, o + 3].sum())
    # the helper times successful matcher calls only (:175-181); a batched matcher has no per-pair time, so the
    # line reports wall-clock / pairs of this rank instead
    n_timed = int((recs[:, o + 3] == 0).sum())
    mt = float(np.mean(recs[recs[:, o + 3] == 0, o + 4])) if (timed and n_timed) else (wall / max(1, len(mine)))
    lprint_(f">>Finished, pairs={n_timed} match_failed={match_failed} matches={np.mean(n_matches):.1f} match_time={mt:.2f}s")
    out = dict(records=recs, task=task, pairs=len(pairs), match_failed=match_failed, n_matches=n_matches,
               wall_s=wall, pairs_per_s=len(mine) / max(wall, 1e-9))
    if "matching" in task:
        i_err, v_err = matching_sums(recs)
        out.update(i_err=i_err, v_err=v_err, summary_matching=summary_matching(recs, thres))
        lprint_("==== Image Matching ====")
        lprint_(out["summary_matching"])
    if "homography" in task:
        h_failed = int(recs[:, o + 2].sum())
        lprint_("==== Homography Estimation ====")
        lprint_(f"Hest solver=cv est_failed={h_failed} ransac_thres={ransac_thres} inlier_rate={np.mean(recs[:, o + 1]):.2f}")
        s, auc, tab = summary_homography(recs, thres)
        print(s)                                                             # the helper prints this table (:52)
        out.update(h_failed=h_failed, inlier_rate=float(np.mean(recs[:, o + 1])), summary_homography=s, auc=auc,
                   dists_sa=recs[:, o], dists_si=recs[recs[:, 1] == 0, o], dists_sv=recs[recs[:, 1] == 1, o], **tab)
    return out


def main(argv=None):
    """``python -m geoformer_b200.hpatches --data_root .../hpatches-sequences-release --ckpt saved_ckpt/geoformer.ckpt``
    (single GPU) or under ``torchrun --nproc-per-node N`` (pairs sharded over N GPUs).  Defaults are the reference's
    (eval_Hpatches.py:84-101 and eval_configs/geoformer.yml 'hpatch': imsize 480, threshold 0.2, no_match_upscale)."""
    import argparse
    import copy
    ap = argparse.ArgumentParser(description="Benchmark HPatches (batched, multi-GPU)")
    ap.add_argument("--data_root", required=True)
    ap.add_argument("--ckpt", default=None, help="reference checkpoint; default: synthetic weights (threshold forced to 0)")
    ap.add_argument("--task", default="homography", choices=["matching", "homography", "both"])
    ap.add_argument("--ransac_thres", type=float, default=3)
    ap.add_argument("--match_threshold", type=float, default=0.2)
    ap.add_argument("--imsize", type=int, default=480)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--debug", action="store_true")
    a = ap.parse_args(argv)
    import torch.distributed as dist
    from . import synth
    from .model.full_model import GeoFormer
    from .model.geo_config import default_cfg as geo_cfg
    from .model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
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
    if a.ckpt:
        sd = torch.load(a.ckpt, map_location="cpu")
        sd = sd.get("state_dict", sd)
    else:
        sd = synth.make_state_dict(0)
    model.load_state_dict(sd, strict=False)
    model = model.eval().to(device)
    matcher = BatchedMatcher(model, device, imsize=a.imsize, no_match_upscale=True, batch=a.batch, depth=a.depth)
    say = print if rank == 0 else (lambda *_: None)
    res = eval_hpatches(matcher, a.data_root, "GeoFormer_b200", task=a.task, scale_H=True, ransac_thres=a.ransac_thres,
                        lprint_=say, debug=a.debug, rank=rank, world=world)
    say(f"{res['pairs']} pairs, {res['pairs_per_s'] * world:.1f} pairs/s over {world} GPU(s) incl. decode, ingest and scoring")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
