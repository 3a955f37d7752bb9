"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the two 'next' rows around the hot path: per-pair restatements of the
reference's downstream evaluation arithmetic and of its image ingest.  Pinned by tests/golden/eval_ingest.npz, which
was produced by the reference's own functions (tests/golden/make_golden.py::eval_and_ingest_case)."""
import numpy as np


def corner_error_one(H_pred, H_gt, w, h):
    """eval_tool/immatch/utils/hpatches_helper.py:221-239 for one pair (w, h already divided by the match scale)."""
    if H_pred is None:
        return np.nan
    corners = np.array([[0, 0, 1], [0, h - 1, 1], [w - 1, 0, 1], [w - 1, h - 1, 1]])
    real = np.dot(corners, np.transpose(H_gt))
    real = real[:, :2] / real[:, 2:]
    pred = np.dot(corners, np.transpose(H_pred))
    pred = pred[:, :2] / pred[:, 2:]
    return np.mean(np.linalg.norm(real - pred, axis=1))


def resize_gray_u8(src, wt, ht):
    """numpy restatement of OpenCV's 8-bit INTER_LINEAR resize (cv::resize, HResizeLinear/VResizeLinear fixed point,
    INTER_RESIZE_COEF_BITS = 11) as called by eval_tool/immatch/utils/data_io.py:58; verified bit-exact against
    cv2.resize 4.13 in the build container."""
    ho, wo = src.shape

    def taps(n_dst, n_src, scale):
        d = np.arange(n_dst)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        f = (f - s).astype(np.float32)
        lo = s < 0
        f[lo] = 0; s[lo] = 0
        hi = s >= n_src - 1
        f[hi] = 0; s[hi] = n_src - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
        a1 = np.rint(f * np.float32(2048)).astype(np.int32)
        return s, np.minimum(s + 1, n_src - 1), a0, a1

    x0, x1, a0, a1 = taps(wt, wo, wo / wt)
    y0, y1, b0, b1 = taps(ht, ho, ho / ht)
    S = src.astype(np.int32)
    rows = S[:, x0] * a0[None, :] + S[:, x1] * a1[None, :]
    out = (((b0[:, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
