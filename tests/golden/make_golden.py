"""Generate golden vectors by running the UNMODIFIED reference (read-only at
/root/reference) in the build container.  The reference cannot travel to the GPU
box, so its outputs are committed here as small .npz fixtures.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference is imported with arithmetic-free stub modules for packages that
are absent from this image (yacs, kornia, skimage, imgaug, pydegensac) — recipe
from SURVEY.md Appendix A.  Weights and images come from geoformer_b200.synth,
which is reference-independent, so the GPU box can rebuild the same inputs.
"""
import copy
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from geoformer_b200 import synth  # noqa: E402


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            return copy.deepcopy(self)

    stub("yacs"); stub("yacs.config", CfgNode=CfgNode)
    stub("kornia"); stub("kornia.geometry"); stub("kornia.geometry.subpix", dsnt=None)
    stub("kornia.utils"); stub("kornia.utils.grid", create_meshgrid=None)
    stub("skimage"); stub("skimage.feature", peak_local_max=None)
    stub("imgaug"); stub("imgaug.augmenters")
    sys.modules["imgaug"].augmenters = sys.modules["imgaug.augmenters"]
    stub("pydegensac")
    sys.path.insert(0, "/root/reference")
    from model.full_model import GeoFormer
    from model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
    from model.geo_config import default_cfg as geo_cfg
    return GeoFormer, default_cfg, geo_cfg


def build_reference(sd, coarse_thr, fine_thr=0.1):
    GeoFormer, default_cfg, geo_cfg = import_reference()
    g = dict(geo_cfg)
    g["coarse_thr"] = coarse_thr
    g["fine_thr"] = fine_thr
    model = GeoFormer(copy.deepcopy(default_cfg), g).eval()
    ref_sd = model.state_dict()
    assert set(ref_sd) == set(sd), (set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    missing = model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return model


def run_reference(model, im0, im1, masks=None):
    cap = {}
    hooks = []

    def grab(name):
        def fn(mod, inp, out):
            cap.setdefault(name, []).append(out)
        return fn

    for nm in ("backbone", "loftr_coarse", "geo_module", "fine_preprocess", "loftr_fine"):
        hooks.append(getattr(model, nm).register_forward_hook(grab(nm)))
    with torch.no_grad():
        inp = {"image0": im0.clone(), "image1": im1.clone()}
        if masks is not None:
            inp.update(mask0=masks[0].clone(), mask1=masks[1].clone())
        data = model(inp)
    for h in hooks:
        h.remove()
    return data, cap


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def small_case(name, h, w, regime, n, seed0, randomize_norm, coarse_thr=0.0, fine_thr=0.1, slim=False):
    sd = synth.make_state_dict(seed=7, randomize_norm=randomize_norm)
    model = build_reference(sd, coarse_thr, fine_thr)
    im0, im1 = synth.make_pairs(n, h, w, regime, seed0)
    data, cap = run_reference(model, im0, im1)
    fc, ff = cap["backbone"][0]
    g0, g1 = cap["geo_module"][0]
    c0, c1 = cap["loftr_coarse"][0]
    out = dict(
        meta=np.array([h, w, n, seed0, int(randomize_norm)]), regime=np.array(regime),
        coarse_thr=np.array(coarse_thr), fine_thr=np.array(fine_thr),
        cnn_c=fc, fine_sub=ff[:, ::8, ::4, ::4],            # subsampled fine map keeps the fixture small
        coarse0=c0, coarse1=c1, geo0=g0, geo1=g1,
        dect_conf=data["dect_conf_matrix"], conf=data["conf_matrix"],
        b_ids=data["b_ids"], i_ids=data["i_ids"], j_ids=data["j_ids"],
        mkpts0_c=data["mkpts0_c"], mkpts1_c=data["mkpts1_c"],
        mkpts0_f=data["mkpts0_f"], mkpts1_f=data["mkpts1_f"], mconf=data["mconf"], m_bids=data["m_bids"],
        fine_matrix=data["fine_matrix"][:64],
    )
    if "fine_preprocess" in cap:
        a, b = cap["fine_preprocess"][0]
        out.update(fine_in0=a[:16], fine_in1=b[:16])
    if "loftr_fine" in cap:
        a, b = cap["loftr_fine"][0]
        out.update(fine_out0=a[:16], fine_out1=b[:16])
    if slim:            # match lists + geo features only (keeps the fixture small)
        for k in ("cnn_c", "fine_sub", "coarse0", "coarse1", "dect_conf", "conf", "fine_matrix", "fine_in0", "fine_in1",
                  "fine_out0", "fine_out1"):
            out.pop(k, None)
        first = cap["coarse_matching"][0] if "coarse_matching" in cap else None
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(out))
    print(name, "M_c", len(data["b_ids"]), "M_f", len(data["mkpts0_f"]),
          "per sample", np.bincount(np.asarray(data["b_ids"]), minlength=n).tolist())


def padded_masks(n, hc, wc):
    """Valid regions of zero-padded images at coarse resolution (MegaDepth collation, datasets/megadepth.py:121-128)."""
    m0, m1 = torch.zeros(n, hc, wc, dtype=torch.bool), torch.zeros(n, hc, wc, dtype=torch.bool)
    for b in range(n):
        m0[b, :hc - 2 - b, :wc - 3] = True
        m1[b, :hc - 1, :wc - 4 - b] = True
    return m0, m1


def masked_case(name, h, w, n, seed0):
    """Optional padding-mask path (oracle only; the product refuses masks): images are zero outside the valid region."""
    sd = synth.make_state_dict(seed=7, randomize_norm=True)
    model = build_reference(sd, 0.0)
    im0, im1 = synth.make_pairs(n, h, w, "dense", seed0)
    m0, m1 = padded_masks(n, h // 8, w // 8)
    im0 = im0 * F.interpolate(m0[:, None].float(), scale_factor=8, mode="nearest")
    im1 = im1 * F.interpolate(m1[:, None].float(), scale_factor=8, mode="nearest")
    data, cap = run_reference(model, im0, im1, (m0, m1))
    c0, c1 = cap["loftr_coarse"][0]
    out = dict(meta=np.array([h, w, n, seed0, 1]), coarse0=c0, coarse1=c1,
               b_ids=data["b_ids"], i_ids=data["i_ids"], j_ids=data["j_ids"],
               mkpts0_f=data["mkpts0_f"], mkpts1_f=data["mkpts1_f"], mconf=data["mconf"], m_bids=data["m_bids"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(out))
    print(name, "M_c", len(data["b_ids"]), "M_f", len(data["mkpts0_f"]))


def full_case(name, h, w, regime, seed0, coarse_thr=0.0):
    """Full-size case: only match lists and scalar summaries are stored."""
    sd = synth.make_state_dict(seed=0)
    model = build_reference(sd, coarse_thr)
    im0, im1 = synth.make_pairs(1, h, w, regime, seed0)
    data, cap = run_reference(model, im0, im1)
    out = dict(
        meta=np.array([h, w, 1, seed0, 0]), regime=np.array(regime), coarse_thr=np.array(coarse_thr),
        i_ids=data["i_ids"].to(torch.int32), j_ids=data["j_ids"].to(torch.int32),
        mkpts0_f=data["mkpts0_f"].to(torch.int16), mkpts1_f=data["mkpts1_f"].to(torch.int16),
        mconf=data["mconf"],
        dect_conf_max=data["dect_conf_matrix"].max(), conf_max=data["conf_matrix"].max(),
        conf_rowsum_head=data["conf_matrix"][0, :16].sum(-1),
        geo0_head=cap["geo_module"][0][0][0, :4, :8], coarse0_head=cap["loftr_coarse"][0][0][0, :4, :8],
    )
    assert torch.equal(data["mkpts0_f"].to(torch.int16).float(), data["mkpts0_f"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(out))
    print(name, "M_c", len(data["b_ids"]), "M_f", len(data["mkpts0_f"]))


def textured_image_u8(h, w, seed):
    """Smooth multi-octave texture as uint8 (stored inside the fixture: bicubic interpolation is not bit-reproducible
    across CPUs, the stored bytes are)."""
    g = torch.Generator().manual_seed(seed)
    acc = torch.zeros(1, 1, h, w)
    for k, amp in ((4, 1.0), (8, 0.8), (16, 0.6), (32, 0.5), (64, 0.35), (128, 0.25)):
        z = torch.rand(1, 1, max(2, h // (512 // k) + 2), max(2, w // (512 // k) + 2), generator=g)
        acc += amp * F.interpolate(z, size=(h, w), mode="bicubic", align_corners=True)
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return (acc[0, 0].numpy() * 255).round().astype(np.uint8)


def full_stage_case(name, h, w, regime, seed0, wseed, randomize_norm, tok_stride=24):
    """Full-size (480x640) case with stage features for the PRODUCT-mode parity tests: match lists in full, features
    subsampled (every `tok_stride`-th token; every 4th x 5th coarse cell of the CNN map) to keep the fixture ~1 MB.
    regime 'shift' uses synth.make_pairs (rebuilt on the GPU box); 'warp' stores its uint8 images in the fixture:
    image1 = cv2.warpPerspective(image0, H_gt)."""
    import cv2
    sd = synth.make_state_dict(seed=wseed, randomize_norm=randomize_norm)
    model = build_reference(sd, 0.0)
    extra = {}
    if regime == "warp":
        im0u = textured_image_u8(h, w, seed0)
        src = np.float32([[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]])
        dst = src + np.float32([[12, 7], [-9, 10], [-14, -8], [10, -11]])
        h_gt = cv2.getPerspectiveTransform(src, dst)
        im1u = cv2.warpPerspective(im0u, h_gt, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
        im0 = torch.from_numpy(im0u).float().div(255)[None, None]
        im1 = torch.from_numpy(im1u).float().div(255)[None, None]
        extra.update(image0_u8=im0u, image1_u8=im1u, h_gt=h_gt)
    else:
        im0, im1 = synth.make_pairs(1, h, w, regime, seed0)
    first = {}
    orig = model.coarse_matching.forward
    calls = []

    def spy(f0, f1, data, **kw):                       # first-pass match list (overwritten by the second call)
        r = orig(f0, f1, data, **kw)
        calls.append({k: data[k].clone() for k in ("b_ids", "i_ids", "j_ids", "mconf")})
        return r
    model.coarse_matching.forward = spy
    data, cap = run_reference(model, im0, im1)
    fc, _ = cap["backbone"][0]
    c0, c1 = cap["loftr_coarse"][0]
    g0, g1 = cap["geo_module"][0]
    out = dict(
        meta=np.array([h, w, 1, seed0, int(randomize_norm)]), wseed=np.array(wseed), regime=np.array(regime),
        tok_stride=np.array(tok_stride),
        cnn_c_sub=fc[:, :, ::4, ::5], coarse0_sub=c0[0, ::tok_stride], coarse1_sub=c1[0, ::tok_stride],
        geo0_sub=g0[0, ::tok_stride], geo1_sub=g1[0, ::tok_stride],
        first_i=calls[0]["i_ids"].to(torch.int32), first_j=calls[0]["j_ids"].to(torch.int32), first_conf=calls[0]["mconf"],
        i_ids=data["i_ids"].to(torch.int32), j_ids=data["j_ids"].to(torch.int32), mconf_c=calls[1]["mconf"],
        mkpts0_f=data["mkpts0_f"].to(torch.int16), mkpts1_f=data["mkpts1_f"].to(torch.int16), mconf=data["mconf"],
        dect_conf_max=data["dect_conf_matrix"].max(), conf_max=data["conf_matrix"].max(), **extra)
    assert torch.equal(data["mkpts0_f"].to(torch.int16).float(), data["mkpts0_f"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(out))
    print(name, "first", len(calls[0]["i_ids"]), "M_c", len(data["b_ids"]), "M_f", len(data["mkpts0_f"]),
          "dect max", float(data["dect_conf_matrix"].max()), os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KB")


def eval_and_ingest_case():
    """Golden vectors for the 'next' rows: downstream evaluation helpers (hpatches_helper.cal_error_auc /
    cal_reproj_dists, fire_helper.compute_auc) and the image ingest (data_io.resize_im + cv2.resize + to_tensor)."""
    import_reference()
    import cv2
    import torchvision.transforms as transforms
    from eval_tool.immatch.utils.hpatches_helper import cal_error_auc, cal_reproj_dists
    from eval_tool.immatch.utils.fire_helper import compute_auc
    from eval_tool.immatch.utils.data_io import resize_im
    rng = np.random.RandomState(3)
    errs = np.abs(rng.randn(57)) * 4
    errs[5] = np.nan
    thr = [1, 3, 5, 10]
    auc = cal_error_auc(errs[~np.isnan(errs)], thr)
    auc_empty = cal_error_auc(np.array([]), thr)
    p1 = rng.rand(40, 2) * 400
    Hm = np.array([[1.01, 0.02, 3.0], [-0.01, 0.99, -2.0], [1e-5, 2e-5, 1.0]])
    p2 = rng.rand(40, 2) * 400
    dists = cal_reproj_dists(p1, p2, Hm)
    s, p, a = np.abs(rng.randn(71)) * 8, np.abs(rng.randn(48)) * 12, np.abs(rng.randn(14)) * 10
    fire = compute_auc(list(s), list(p), list(a))
    dims = []
    for (wo, ho, imsize, df) in [(1024, 768, 480, 8), (800, 600, 480, 8), (2912, 2912, 768, 8), (517, 333, 480, 8), (640, 480, 480, 8)]:
        wt, ht, sc = resize_im(wo, ho, imsize=imsize, dfactor=df, value_to_scale=min)
        dims.append([wo, ho, imsize, df, wt, ht, sc[0], sc[1]])
    im = rng.randint(0, 256, (333, 517)).astype(np.uint8)
    wt, ht, sc = resize_im(517, 333, imsize=200, dfactor=8, value_to_scale=min)
    t = transforms.functional.to_tensor(cv2.resize(im, (wt, ht))).unsqueeze(0)
    np.savez_compressed(os.path.join(HERE, "eval_ingest.npz"), errs=errs, thr=np.array(thr), auc=auc, auc_empty=auc_empty,
                        p1=p1, p2=p2, H=Hm, dists=dists, fire_s=s, fire_p=p, fire_a=a,
                        fire=np.array([fire['s'], fire['p'], fire['a'], fire['mAUC']]), dims=np.array(dims),
                        im=im, im_resized=t.numpy(), im_hw=np.array([ht, wt]))
    print("eval_ingest", auc, fire, t.shape)


def hpatches_case():
    """The reference's own benchmark loop (eval_tool/immatch/utils/hpatches_helper.py::eval_hpatches, unmodified) on a
    synthetic HPatches-shaped tree with a deterministic stand-in matcher (tests/util.py): everything it logs, prints and
    hands to its summary functions, for both wrapper conventions (no_match_upscale on / off), stored as JSON."""
    import contextlib
    import io
    import json
    import tempfile
    import_reference()
    import eval_tool.immatch.utils.hpatches_helper as helper
    from tests.util import make_hpatches_tree, stub_matcher
    out = {}
    with tempfile.TemporaryDirectory() as root:
        make_hpatches_tree(root)
        for tag, scaled, task, rthr in (("scaled_both", True, "both", 3), ("plain_both", False, "both", 3),
                                        ("scaled_homography", True, "homography", 2)):
            logged, grabbed = [], {}
            orig_m, orig_h = helper.eval_summary_matching, helper.eval_summary_homography

            def spy_m(results, thres=[1, 3, 5, 10], save_npy=None):
                i_err, v_err, (seq_type, n_feats, n_matches) = results
                grabbed.update(i_err={int(k): float(v) for k, v in i_err.items()}, v_err={int(k): float(v) for k, v in v_err.items()},
                               seq_type=[str(x) for x in seq_type], n_feats=[int(x) for x in n_feats], n_matches=[int(x) for x in n_matches])
                return orig_m(results, thres, save_npy)

            def spy_h(dists_sa, dists_si, dists_sv, thres):
                grabbed.update(dists_sa=[float(x) for x in dists_sa], dists_si=[float(x) for x in dists_si], dists_sv=[float(x) for x in dists_sv])
                r = orig_h(dists_sa, dists_si, dists_sv, thres)
                grabbed["auc"] = float(r)
                return r
            helper.eval_summary_matching, helper.eval_summary_homography = spy_m, spy_h
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                helper.eval_hpatches(stub_matcher(scaled), root, "stub", task=task, scale_H=scaled, h_solver="cv",
                                     ransac_thres=rthr, lprint_=logged.append, print_out=False)
            helper.eval_summary_matching, helper.eval_summary_homography = orig_m, orig_h
            out[tag] = dict(scaled=scaled, task=task, ransac_thres=rthr, logged=logged, stdout=buf.getvalue(), **grabbed)
            print(tag, "logged", len(logged), "lines; auc", grabbed.get("auc"))
    with open(os.path.join(HERE, "hpatches_eval.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def fire_isc_case():
    """The reference's FIRE and ISC-HE benchmark loops (fire_helper.eval_fire, my_helper.eval_homography_my, unmodified)
    on synthetic inputs with deterministic stand-in matchers (tests/util.py): logged lines, stdout, returned value and
    the per-pair distances handed to the summaries."""
    import contextlib
    import io
    import json
    import tempfile
    import_reference()
    import eval_tool.immatch.utils.fire_helper as fire
    import eval_tool.immatch.utils.my_helper as isc
    from tests.util import make_fire_tree, make_isc_tree, stub_matcher_named
    out = {}
    with tempfile.TemporaryDirectory() as root:
        files, im_dir, gt_dir = make_fire_tree(os.path.join(root, "fire"))
        triples = make_isc_tree(os.path.join(root, "isc"))
        for tag, scaled, rthr in (("fire_scaled", True, 15), ("fire_plain", False, 15), ("isc_scaled", True, 3), ("isc_plain", False, 3)):
            logged, grabbed = [], {}
            buf = io.StringIO()
            if tag.startswith("fire"):
                orig = fire.eval_summary_homography

                def spy(ss, sp, sa):
                    grabbed.update(dists_ss=[float(x) for x in ss], dists_sp=[float(x) for x in sp], dists_sa=[float(x) for x in sa])
                    return orig(ss, sp, sa)
                fire.eval_summary_homography = spy
                with contextlib.redirect_stdout(buf):
                    r = fire.eval_fire(stub_matcher_named("fire", scaled), files, im_dir, gt_dir, "stub", task="homography",
                                       scale_H=scaled, h_solver="cv", ransac_thres=rthr, lprint_=logged.append)
                fire.eval_summary_homography = orig
            else:
                orig = isc.eval_summary_homography

                def spy(sa, thres):
                    grabbed.update(dists_all=[float(x) for x in sa])
                    return orig(sa, thres)
                isc.eval_summary_homography = spy
                with contextlib.redirect_stdout(buf):
                    r = isc.eval_homography_my(stub_matcher_named("isc", scaled), triples, "stub", task="homography",
                                               scale_H=scaled, h_solver="cv", ransac_thres=rthr, lprint_=logged.append)
                isc.eval_summary_homography = orig
            out[tag] = dict(scaled=scaled, ransac_thres=rthr, logged=logged, stdout=buf.getvalue(), value=float(r), **grabbed)
            print(tag, "value", float(r))
    with open(os.path.join(HERE, "fire_isc_eval.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "--only-hpatches" in sys.argv:
        hpatches_case()
        sys.exit(0)
    if "--only-fire-isc" in sys.argv:
        fire_isc_case()
        sys.exit(0)
    if "--only-masked" in sys.argv:
        masked_case("small_masked", 96, 128, 2, 40)
        sys.exit(0)
    if "--only-stage" in sys.argv:
        full_stage_case("full_shift_480x640", 480, 640, "shift", 10, 0, False)            # flat confidences (default norms)
        full_stage_case("full_shift_rn_480x640", 480, 640, "shift", 10, 7, True)          # peaky confidences
        full_stage_case("full_warp_rn_480x640", 480, 640, "warp", 3, 7, True)             # cv2.warpPerspective pair
        sys.exit(0)
    if "--only-mixed" in sys.argv:
        small_case("small_mixed", 96, 128, "mixed", 3, 30, True, slim=True)     # dense + unrelated (noise matches) + shift in one batch
        sys.exit(0)
    if "--only-eval" not in sys.argv:
        masked_case("small_masked", 96, 128, 2, 40)
        small_case("small_mixed", 96, 128, "mixed", 3, 30, True, slim=True)
        small_case("small_dense", 96, 128, "dense", 2, 0, True)
        small_case("small_shift", 96, 128, "shift", 1, 10, True)
        small_case("small_rect_thr", 64, 96, "dense", 1, 20, False, coarse_thr=0.2)   # zero-match corner
    if "--only-eval" not in sys.argv:
        full_case("full_dense_480x640", 480, 640, "dense", 0)
        full_stage_case("full_shift_480x640", 480, 640, "shift", 10, 0, False)
        full_stage_case("full_shift_rn_480x640", 480, 640, "shift", 10, 7, True)
        full_stage_case("full_warp_rn_480x640", 480, 640, "warp", 3, 7, True)
    eval_and_ingest_case()
