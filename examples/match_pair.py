"""Match two image files with the B200 implementation — the same call sequence the reference's ``inference.py``
wrapper performs (inference.py:12-99: load gray -> min-side resize to a multiple of 8 -> GeoFormer.forward ->
mkpts0_f / mkpts1_f / mconf -> scale back to the original resolution), with the image ingest on the GPU.

    python examples/match_pair.py a.png b.png [--ckpt saved_ckpt/geoformer.ckpt] [--imsize 640] [--thr 0.2]

Without --ckpt the deterministic synthetic weights of geoformer_b200.synth are used (random features: the match
threshold is then forced to 0)."""
import argparse
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoformer_b200 import synth  # noqa: E402
from geoformer_b200.ingest import load_gray_scale_tensor_gpu  # noqa: E402
from geoformer_b200.model.full_model import GeoFormer  # noqa: E402
from geoformer_b200.model.geo_config import default_cfg as geo_cfg  # noqa: E402
from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("image0"); ap.add_argument("image1")
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--imsize", type=int, default=640)
    ap.add_argument("--thr", type=float, default=0.2)
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args()

    thr = args.thr if args.ckpt else 0.0
    conf, g = copy.deepcopy(default_cfg), dict(geo_cfg)
    conf["match_coarse"]["thr"] = thr
    g["coarse_thr"] = thr
    model = GeoFormer(conf, g)
    if args.ckpt:
        sd = torch.load(args.ckpt, map_location="cpu")
        sd = sd.get("state_dict", sd)
    else:
        sd = synth.make_state_dict(0)
    model.load_state_dict(sd, strict=False)                 # strips the 'matcher.' prefix like the reference
    model = model.eval().to(args.device)

    gray0, sc0 = load_gray_scale_tensor_gpu(args.image0, args.device, imsize=args.imsize, dfactor=8, value_to_scale=min)
    gray1, sc1 = load_gray_scale_tensor_gpu(args.image1, args.device, imsize=args.imsize, dfactor=8, value_to_scale=min)
    out = model({"image0": gray0, "image1": gray1})
    k0 = out["mkpts0_f"].cpu().numpy() * np.asarray(sc0)    # back to the original image resolution
    k1 = out["mkpts1_f"].cpu().numpy() * np.asarray(sc1)
    conf_ = out["mconf"].cpu().numpy()
    print(f"{len(k0)} matches ({gray0.shape[-1]}x{gray0.shape[-2]} vs {gray1.shape[-1]}x{gray1.shape[-2]} after resize)")
    for a, b, c in list(zip(k0, k1, conf_))[:10]:
        print(f"  ({a[0]:7.1f}, {a[1]:7.1f}) -> ({b[0]:7.1f}, {b[1]:7.1f})   conf {c:.3f}")


if __name__ == "__main__":
    main()
