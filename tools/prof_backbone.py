"""Time the two backbones of the library on 32 images of 480x640 (run on the GPU box): the fp16 tcgen05 implicit-GEMM path
(product) and the fp32 FFMA reference kernels (accurate mode), and report their relative difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import synth, engine, ops

dev = torch.device("cuda:0")
ops.ensure_init(dev)
sd = synth.make_state_dict(0)
img = torch.rand(32, 1, 480, 640, device=dev)
out = {}
for name, dt, reps in (("fp32 FFMA reference (conv_ref.cu)", torch.float32, 1), ("fp16 tcgen05 (conv_tc.cu)", torch.float16, 5)):
    pw = engine.PackedWeights(sd, dev, dt)
    for _ in range(2):
        c, f = engine.backbone_forward(pw, img)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        c, f = engine.backbone_forward(pw, img)
    e1.record(); torch.cuda.synchronize()
    out[name] = (c.float(), f.float())
    print(f"{name:36s} {e0.elapsed_time(e1) / reps:9.2f} ms / 32 images", flush=True)
(rc, rf), (tc, tf) = out.values()
print(f"fp16 vs fp32: coarse max-rel {((tc - rc).abs().max() / rc.abs().max()).item():.2e}, fine {((tf - rf).abs().max() / rf.abs().max()).item():.2e}")
