"""Time the cuDNN backbone under different channel paddings / dtypes (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import synth, engine, ops

dev = torch.device("cuda:0")
ops.ensure_init(dev)
sd = synth.make_state_dict(0)
img = torch.rand(32, 1, 480, 640, device=dev)
torch.backends.cudnn.benchmark = True
ref = None
for dt_name, dt in (("f16", torch.float16), ("fp16", torch.float16)):
    for cpad in (1, 8, 16, 32, 64):
        os.environ["GF_BACKBONE_CPAD"] = str(cpad)
        pw = engine.PackedWeights(sd, dev, dt)
        for _ in range(3):
            c, f = engine.backbone_forward(pw, img)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            c, f = engine.backbone_forward(pw, img)
        e1.record(); torch.cuda.synchronize()
        if ref is None:
            ref = (c.clone(), f.clone())
        dc = (c - ref[0]).abs().max().item() / ref[0].abs().max().item()
        df = (f - ref[1]).abs().max().item() / ref[1].abs().max().item()
        print(f"{dt_name} cpad={cpad:3d}: {e0.elapsed_time(e1)/5:8.2f} ms / 32 images   rel diff vs first: coarse {dc:.2e} fine {df:.2e}", flush=True)
