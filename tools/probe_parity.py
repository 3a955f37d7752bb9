"""Which precision choice costs how much match-set parity at 480x640?  (diagnostic; prints a table)"""
import copy, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoformer_b200 import ops, engine
from geoformer_b200.model.full_model import GeoFormer
from geoformer_b200.model.geo_config import default_cfg as geo_cfg
from geoformer_b200.model.loftr_src.loftr.utils.cvpr_ds_config import default_cfg
from tests.util import load_golden, stage_case_inputs

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def iou(a, b):
    a, b = set(map(tuple, np.asarray(a).tolist())), set(map(tuple, np.asarray(b).tolist()))
    return len(a & b) / max(1, len(a | b))

def rms(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()

modes = {
    "product": dict(bb="f16", linear="tf32", sim="f16x3", attn="tf32", act="f16"),
    "bb=fp32": dict(bb="fp32", linear="tf32", sim="f16x3", attn="tf32", act="f16"),
    "bb=fp32,act=f32": dict(bb="fp32", linear="tf32", sim="f16x3", attn="tf32", act="f32"),
    "bb=fp16,lin=ref": dict(bb="f16", linear="ref", sim="ref", attn="ref", act="f32"),
    "all ref": dict(bb="fp32", linear="ref", sim="ref", attn="ref", act="f32"),
}
for name in sys.argv[1:] or ["full_shift_rn_480x640", "full_warp_rn_480x640", "full_shift_480x640"]:
    g = load_golden(GOLD, name)
    sd, im0, im1 = stage_case_inputs(g)
    ts = int(g["tok_stride"])
    for mname, m in modes.items():
        gc = dict(geo_cfg); gc["coarse_thr"] = 0.0
        model = GeoFormer(copy.deepcopy(default_cfg), gc)
        model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        model.backbone_precision = m["bb"]
        model = model.eval().to("cuda:0")
        ops.set_precision(linear=m["linear"], similarity=m["sim"], attention=m["attn"], activations=m["act"])
        model.capture = True
        d = model({"image0": im0.cuda(), "image1": im1.cuda()})
        st = d["_stages"]
        cnn = torch.cat([st["cnn_c0"], st["cnn_c1"]], 0).permute(0, 3, 1, 2)[:, :, ::4, ::5]
        first = st["first"]
        got_f = torch.cat([d["mkpts0_f"], d["mkpts1_f"]], 1).long().cpu().numpy()
        want_f = np.concatenate([g["mkpts0_f"], g["mkpts1_f"]], 1).astype(np.int64)
        print(f"{name:24s} {mname:18s} cnn {rms(cnn, g['cnn_c_sub']):.2e} coarse {rms(st['coarse0'][0, ::ts], g['coarse0_sub']):.2e} "
              f"geo {rms(st['geo0'][0, ::ts], g['geo0_sub']):.2e}  IoU first {iou(torch.stack([first['i_ids'], first['j_ids']], 1).cpu().numpy(), np.stack([g['first_i'], g['first_j']], 1)):.3f} "
              f"coarse {iou(torch.stack([d['i_ids'], d['j_ids']], 1).cpu().numpy(), np.stack([g['i_ids'], g['j_ids']], 1)):.3f} fine {iou(got_f, want_f):.3f}  n {len(got_f)}/{len(want_f)}", flush=True)
