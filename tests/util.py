"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

from geoformer_b200 import synth


def load_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def stage_case_inputs(g):
    """(state dict, image0, image1) of a `full_stage_case` fixture (tests/golden/make_golden.py): synth images, or the
    uint8 pair stored inside the fixture."""
    h, w, n, seed0, rnd = [int(v) for v in g["meta"]]
    P = synth.make_state_dict(seed=int(g["wseed"]), randomize_norm=bool(rnd))
    if "image0_u8" in g:
        im0 = torch.from_numpy(g["image0_u8"]).float().div(255)[None, None]
        im1 = torch.from_numpy(g["image1_u8"]).float().div(255)[None, None]
    else:
        im0, im1 = synth.make_pairs(n, h, w, str(g["regime"]), seed0)
    return P, im0, im1
