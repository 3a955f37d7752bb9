"""Three backbone conv layers, two launches each, for an `ncu --set full -k regex:conv3x3_tc` capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geoformer_b200 import ops
from geoformer_b200.engine import pack_conv3x3
dev = torch.device("cuda:0"); ops.ensure_init(dev)
for (cin, cout, cp_in, cp_out, h, w) in [(128, 128, 128, 128, 240, 320), (196, 196, 200, 200, 240, 320), (256, 256, 256, 256, 120, 160)]:
    x = torch.randn(32, h, w, cp_in, device=dev).half()
    wt, bias = pack_conv3x3(torch.randn(cout, cin, 3, 3) * 0.03, torch.zeros(cout), cp_in, cp_out, dev)
    for _ in range(2):
        ops.conv3x3(x, wt, bias, None, 1)
    torch.cuda.synchronize()
