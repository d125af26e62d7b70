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
for name in ("rand", "empty", "zeros", "tensor"):
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
for kind in ("c2c", "r2c", "r2r", "conv"):
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--size", "32", "32", "32", "--kind", kind]
    bench.main()
