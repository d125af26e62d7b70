# Developer aid: runs bench.py's b200 arm end to end on the CPU at a tiny size -- emulated library, torch CPU tensors standing
# for device memory -- to catch Python-level mistakes in bench.py before GPU time is spent.  The numbers it prints are meaningless.
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.emul.build_emul_library import build
from heffte_b200 import _lib
_lib.LIB_PATH = build()

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None


class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, *a): import time; self.t = time.perf_counter()
    def elapsed_time(self, other): return max((other.t - self.t) * 1e3, 1e-3)


torch.cuda.Event = FakeEvent
_Gen = torch.Generator
torch.Generator = lambda device=None: _Gen()
for name in ("rand", "empty", "zeros", "tensor", "arange"):
    def make(orig):
        def f(*a, **k):
            k.pop("device", None)
            return orig(*a, **k)
        return f
    setattr(torch, name, make(getattr(torch, name)))
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.is_cuda = property(lambda self: True)

import bench  # noqa: E402
bench.ClockSampler.start = lambda self: None
bench.run_reference_speed3d = lambda *a, **k: None
torch.cuda.empty_cache = lambda: None
bench.secondary_list = lambda args, ws: [dict(kind="r2c", size=(32, 16, 16), precision="double"), dict(kind="conv", size=(16, 16, 32), precision="double"),
                                         dict(kind="c2c", size=(16, 16, 16), precision="float", force_flush=True)]
_orig_flush = bench.flush_l2
def _small_flush(ctx):
    if ctx.flush is None:
        ctx.flush = torch.empty(1024)
    ctx.flush.fill_(1.0)
bench.flush_l2 = _small_flush
for kind in ("c2c", "r2c", "r2r", "conv"):
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--size", "32", "32", "32", "--kind", kind]
    bench.main()
# the default workload triggers the secondary list: shrink it through the argument parser's defaults
_parse = bench.parse_args
def _tiny_default():
    a = _parse()
    return a
sys.argv = ["bench.py", "--steps", "2", "--warmup", "1"]
bench.workload_default = None
import argparse
_orig_run = bench.run_workload
def _run(ctx, args, kind, size, *a, **k):
    if tuple(size) == (512, 512, 512):
        size = (32, 32, 32)
    return _orig_run(ctx, args, kind, size, *a, **k)
bench.run_workload = _run
bench.main()
