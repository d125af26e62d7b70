# Developer aid: runs the logic of tests/test_y_fullsize_gpu.py on the CPU at 32^3 against the emulated library (torch CPU tensors).
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.emul.build_emul_library import build
from heffte_b200 import _lib
_lib.LIB_PATH = build()
import heffte_b200 as hf
import tests.test_y_fullsize_gpu as T

# "cuda" -> cpu
_Gen = torch.Generator
torch.Generator = lambda device=None: _Gen()
for name in ("rand", "empty", "zeros"):
    orig = getattr(torch, name)
    def make(orig):
        def f(*a, **k):
            k.pop("device", None)
            return orig(*a, **k)
        return f
    setattr(torch, name, make(orig))
torch.Tensor.cuda = lambda self, *a, **k: self

class Shim:
    def __init__(self, plan): self.p = plan
    def __getattr__(self, k): return getattr(self.p, k)
    def _np(self, t): return t.numpy() if isinstance(t, torch.Tensor) else t
    def forward(self, a, b, scaling=0, batch=1): self.p.forward(self._np(a.contiguous()), self._np(b), scaling, batch)
    def backward(self, a, b, scaling=0, batch=1): self.p.backward(self._np(a.contiguous()), self._np(b), scaling, batch)
    def forward_buffered(self, a, b, w, scaling=0, batch=1):
        src = self._np(a).copy(); self.p.forward(src, self._np(b), scaling, batch)
    def backward_buffered(self, a, b, w, scaling=0, batch=1):
        src = self._np(a).copy(); self.p.backward(src, self._np(b), scaling, batch)
orig_plan = T._plan
T._plan = lambda kind, n, hf_: Shim(orig_plan(kind, n, hf_))
T.FULL = (32, 32, 32)
T.test_c2c_properties_at_full_size(None, (32, 32, 32), 1); print("c2c fp64 ok")
T.test_c2c_properties_at_full_size(None, (32, 16, 64), 0); print("c2c fp32 ok")
T.test_r2c_matches_the_complex_plan_at_512(None); print("r2c ok")
T.test_dct_properties_at_512(None); print("dct ok")
from oracle import ref_lib
for kind, n, prec in [("c2c", (32, 16, 64), 1), ("r2c", (32, 32, 32), 1), ("cos", (32, 32, 32), 1), ("c2c", (32, 32, 32), 0), ("sin", (32, 64, 32), 1), ("r2c", (64, 32, 32), 0)]:
    T.test_full_size_against_the_compiled_reference(None, ref_lib, kind, n, prec); print("reference parity", kind, n, prec, "ok")
