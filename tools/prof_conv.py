"""Time the tcgen05 conv at backbone sizes (32 images); the torch/cuDNN column is a tool-side yardstick only (the package
itself calls no vendor library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from geoformer_b200 import ops
from geoformer_b200.engine import pack_conv3x3
dev = torch.device("cuda:0"); ops.ensure_init(dev)
torch.backends.cudnn.benchmark = True
def timeit(fn, k=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for (cin, cout, cp_in, cp_out, h, w) in [(128,128,128,128,240,320), (196,196,200,200,240,320), (196,128,200,128,240,320),
                                         (196,196,200,200,120,160), (256,256,256,256,120,160), (256,196,256,200,120,160), (256,256,256,256,60,80)]:
    b = 32
    x = torch.randn(b, h, w, cp_in, device=dev).half()
    wgt = torch.randn(cout, cin, 3, 3) * 0.03
    wt, bias = pack_conv3x3(wgt, torch.zeros(cout), cp_in, cp_out, dev)
    wc = torch.zeros(cp_out, cp_in, 3, 3); wc[:cout, :cin] = wgt
    wc = wc.to(dev).half().contiguous(memory_format=torch.channels_last)
    bc = torch.zeros(cp_out, device=dev).half()
    xc = x.permute(0, 3, 1, 2)
    t_my = timeit(lambda: ops.conv3x3(x, wt, bias, None, 1))
    t_cd = timeit(lambda: F.relu_(F.conv2d(xc, wc, bc, 1, 1)))
    fl = 2.0 * b * h * w * cin * cout * 9
    print(f"{cin:3d}->{cout:3d} @{h}x{w}: tcgen05 {t_my:7.3f} ms ({fl/t_my/1e9:6.0f} TFLOP/s alg)   cuDNN {t_cd:7.3f} ms ({fl/t_cd/1e9:6.0f} TFLOP/s)", flush=True)
