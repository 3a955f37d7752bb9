"""Downstream evaluation step: what the HPatches / ISC / FIRE helpers do with the matcher's
output — RANSAC homography from the matches, corner error against the ground truth, accuracy / AUC tables
(eval_tool/immatch/utils/hpatches_helper.py:13-25, 213-240; fire_helper.py:11-42).  The per-pair cv2 calls run in
a thread pool (OpenCV releases the GIL) and the metric arithmetic is vectorised over all pairs, so the serial
per-pair loop of the reference stops being the bottleneck when 8 GPUs produce > 2000 match lists per second."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

_POOL: Optional[ThreadPoolExecutor] = None


def _pool() -> ThreadPoolExecutor:
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=max(2, min(32, os.cpu_count() or 2)))
    return _POOL


def error_auc(errors: np.ndarray, thresholds: Sequence[float]) -> np.ndarray:
    """Area under the recall-vs-error curve up to each threshold, normalised (hpatches_helper.py:13-25)."""
    errors = np.asarray(errors, dtype=float)
    if errors.size == 0:
        return np.zeros(len(thresholds))
    n = errors.size
    e = np.concatenate([[0.0], np.sort(errors)])
    recall = np.arange(n + 1) / n
    out = []
    for thr in thresholds:
        k = int(np.searchsorted(e, thr))
        out.append((np.trapezoid if hasattr(np, "trapezoid") else np.trapz)(np.append(recall[:k], recall[k - 1]), x=np.append(e[:k], thr)) / thr)
    return np.array(out, dtype=float)


def reproj_dists(p1s: np.ndarray, p2s: np.ndarray, homography: np.ndarray) -> np.ndarray:
    """Distance between p2s and H * p1s (hpatches_helper.py:27-36)."""
    p1h = np.concatenate([p1s, np.ones((p1s.shape[0], 1))], axis=1)
    p2p = p1h @ np.asarray(homography).T
    p2p = p2p[:, :2] / p2p[:, 2:]
    return np.sqrt(np.sum((p2s - p2p) ** 2, axis=1))


def estimate_homographies(matches: Sequence[np.ndarray], ransac_thres: float = 3.0) -> List[Tuple[Optional[np.ndarray], np.ndarray]]:
    """cv2.findHomography(matches[:, :2], matches[:, 2:4], RANSAC, thr) for many pairs in parallel
    (hpatches_helper.py:213-220); (None, []) where OpenCV fails or throws (counted as h_failed by the caller)."""
    import cv2

    def one(m):
        try:
            if len(m) < 4:
                return None, np.zeros(0, dtype=np.uint8)
            H, inl = cv2.findHomography(m[:, :2], m[:, 2:4], cv2.RANSAC, ransac_thres)
            return H, (inl[:, 0] if inl is not None else np.zeros(0, dtype=np.uint8))
        except cv2.error:
            return None, np.zeros(0, dtype=np.uint8)

    return list(_pool().map(one, matches))


def corner_errors(H_pred: Sequence[Optional[np.ndarray]], H_gt: Sequence[np.ndarray], sizes_wh: np.ndarray) -> np.ndarray:
    """Mean distance of the 4 image corners warped by the predicted vs the ground-truth homography, for all pairs at
    once (hpatches_helper.py:228-239); NaN where no homography was estimated."""
    n = len(H_gt)
    sizes_wh = np.asarray(sizes_wh, dtype=float).reshape(n, 2)
    w, h = sizes_wh[:, 0], sizes_wh[:, 1]
    one, zero = np.ones(n), np.zeros(n)
    corners = np.stack([np.stack([zero, zero, one], 1), np.stack([zero, h - 1, one], 1),
                        np.stack([w - 1, zero, one], 1), np.stack([w - 1, h - 1, one], 1)], 1)        # [n, 4, 3]
    ok = np.array([H is not None for H in H_pred])
    Hp = np.stack([H if H is not None else np.eye(3) for H in H_pred]).astype(float)
    Hg = np.stack([np.asarray(H, dtype=float) for H in H_gt])
    real = corners @ np.transpose(Hg, (0, 2, 1))
    real = real[..., :2] / real[..., 2:]
    pred = corners @ np.transpose(Hp, (0, 2, 1))
    pred = pred[..., :2] / pred[..., 2:]
    err = np.mean(np.linalg.norm(real - pred, axis=2), axis=1)
    err[~ok] = np.nan
    return err


def homography_summary(corner_dists: np.ndarray, thresholds: Sequence[float] = (1, 3, 5, 10)) -> dict:
    """Accuracy (fraction of pairs with corner error <= thr) and AUC rows of the README table
    (hpatches_helper.py:262-317)."""
    d = np.asarray(corner_dists, dtype=float)
    acc = np.array([np.mean(d <= t) for t in thresholds]) if d.size else np.zeros(len(thresholds))
    return {"accuracy": acc, "auc": error_auc(d, thresholds), "failed": int(np.isnan(d).sum())}


def fire_auc(errors: np.ndarray, limit: int = 25) -> float:
    """AUC of the success-rate curve for thresholds 1..limit px (fire_helper.py:24-40, one category)."""
    e = np.asarray(errors, dtype=float)
    return float(sum(np.sum(e < i) * 100 / len(e) for i in range(1, limit + 1)) / (limit * 100))
