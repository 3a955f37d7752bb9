"""Diagnostic: per-batch wall time and host-side RANSAC time through MatchPipeline (looks for sporadic stalls)."""
import os, sys, time, copy, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from geoformer_b200 import synth, engine, ops
from geoformer_b200.pipeline import MatchPipeline
import bench
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ransac = sys.argv[2] if len(sys.argv) > 2 else "cv2"
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
model = bench.build_model(dev, "f16", ransac)
host = [synth.make_pairs(16, 480, 640, "dense", 100 * p) for p in range(2)]
devb = [(a.to(dev), b.to(dev)) for a, b in host]
rt = []
orig = engine.geo_prepare_host
def timed(*a, **k):
    t0 = time.perf_counter(); r = orig(*a, **k); rt.append((time.perf_counter() - t0) * 1e3); return r
engine.geo_prepare_host = timed
jt = []
pipe = MatchPipeline(model, depth=depth, device=dev)
oj = pipe._job
def tj(slot, data, post):
    t0 = time.perf_counter(); r = oj(slot, data, post); jt.append((time.perf_counter() - t0) * 1e3); return r
pipe._job = tj
post = lambda d: int(d["mkpts0_f"].shape[0])
pipe.run(({"image0": devb[i % 2][0], "image1": devb[i % 2][1]} for i in range(6)), post)
torch.cuda.synchronize(); rt.clear(); jt.clear()
t0 = time.perf_counter()
pipe.run(({"image0": devb[i % 2][0], "image1": devb[i % 2][1]} for i in range(nb)), post)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) * 1e3
print(f"depth={depth} ransac={ransac}: {tot / nb:.1f} ms/batch  ({16 * nb / tot * 1e3:.0f} pairs/s)")
print("job ms:   ", " ".join(f"{x:.0f}" for x in jt))
print("ransac ms:", " ".join(f"{x:.0f}" for x in rt))
st = torch.cuda.memory_stats()
print("alloc retries", st.get("num_alloc_retries"), "cudaMalloc calls", st.get("num_device_alloc"), "frees", st.get("num_device_free"),
      "reserved GB", torch.cuda.memory_reserved() / 2**30)
