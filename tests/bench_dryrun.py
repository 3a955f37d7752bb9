"""Dry run of bench.py's OWN arm without a GPU (test infrastructure; started as a subprocess by
tests/test_host_pipeline_bench.py, one process per simulated rank).

What it is for: bench.py's control flow around the kernels — warm-up / settle / timed loops through MatchPipeline's
worker threads, the per-batch match-list exchange issued from the main thread, barriers, the max-over-ranks reduction,
the rank-0-only passes, the diagnostics blocks and the assembly of the JSON line — only ever ran on GPU boxes, and a
mismatch of the collective sequence between ranks hangs a multi-GPU run (it did once).  Here the same file
runs with the device plumbing faked: operators -> tests/emu_ops.py (torch CPU), CUDA streams / events -> no-ops and
perf_counter, NCCL -> gloo, direct C-ABI launches -> no-ops.  Numbers it prints are meaningless; finishing on every rank
with one JSON line on rank 0 is the result.

    RANK=0 WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29511 python tests/bench_dryrun.py --gpus 2 --steps 2 --warmup 1 ...
"""
import contextlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pytest  # noqa: E402
import torch  # noqa: E402


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t)

    def synchronize(self):
        pass


def _fake_nvml():
    """pynvml stand-in so that rank 0 takes its clock-sampling branches (incl. the rank-0-only densely sampled pass)."""
    import types
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    m.nvmlInit = lambda: None
    m.nvmlDeviceGetHandleByIndex = lambda i: object()
    m.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    m.nvmlDeviceGetClockInfo = lambda h, c: 1600
    m.nvmlDeviceGetCurrentClocksEventReasons = lambda h: 0x4
    sys.modules["pynvml"] = m


def install():
    _fake_nvml()
    mp = pytest.MonkeyPatch()
    from tests import emu_ops
    emu_ops.install(mp)                                          # ops.* -> torch CPU emulations, ensure_init, current_stream
    mp.setattr(torch.cuda, "Stream", _Stream)
    mp.setattr(torch.cuda, "Event", _Event)
    mp.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    mp.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    mp.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    mp.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    mp.setattr(torch.cuda, "set_device", lambda d: None)
    mp.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    mp.setattr(torch.Tensor, "record_stream", lambda self, s: None)
    from geoformer_b200 import _lib
    from geoformer_b200.model.full_model import GeoFormer
    mp.setattr(_lib, "call", lambda name, *a: None)              # bench.py's direct launches (kernel timing blocks)
    mp.setattr(GeoFormer, "forward", lambda self, data: self._forward(data, data["image0"], data["image1"]))
    return mp


def main():
    install()
    src = open(os.path.join(ROOT, "bench.py")).read()
    for old, new in (('device = torch.device("cuda", local)', 'device = torch.device("cpu")'),
                     ('dist.init_process_group("nccl", device_id=device, ', 'dist.init_process_group("gloo", ')):
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    os.environ.setdefault("GF_BENCH_PREWARM", "1")
    if os.environ.get("GF_DRYRUN_INJECT_RANK0_COLLECTIVE"):      # the r02 bug, to show that this harness catches it
        old = "run_resident(max(4, args.steps // 2), collective=False)"
        assert src.count(old) == 1
        src = src.replace(old, "run_resident(max(4, args.steps // 2), collective=True)")
    if os.environ.get("GF_DRYRUN_FAIL_EXTRA_ON_RANK1"):          # rank 1's measurement of the informational configs raises
        from geoformer_b200 import synth
        orig = synth.make_pairs

        def make_pairs(n, h, w, regime, seed):
            if 8000 <= seed < 9000:                              # bench.other_configs seeds: 7000 + 1000 * rank + 100 * p
                raise MemoryError("simulated failure on rank 1")
            return orig(n, h, w, regime, seed)
        synth.make_pairs = make_pairs
    g = {"__name__": "__main__", "__file__": os.path.join(ROOT, "bench.py")}
    exec(compile(src, os.path.join(ROOT, "bench.py"), "exec"), g)


if __name__ == "__main__":
    main()
