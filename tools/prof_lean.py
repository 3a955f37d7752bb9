import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops
dev = torch.device("cuda:0"); ops.ensure_init(dev)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(153600, 256, device=dev, generator=g)
w = torch.randn(768, 256, device=dev, generator=g) / 16
for _ in range(3):
    y = ops.linear(x, w)
torch.cuda.synchronize()
