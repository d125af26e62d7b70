"""
TEST INFRASTRUCTURE (GPU).  The TMA-loaded tile of the 512-point strided kernel on small boxes: HEFFTE_B200_TMA=force takes it
whenever the box allows, both precisions, both directions, middle and slow axis, with and without a batch; compared with the
oracle.  Run by tests/test_gpu_lengths.py in a subprocess (the switch is read once per process).
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["HEFFTE_B200_TMA"] = "force"

import numpy as np   # noqa: E402
import torch         # noqa: E402

from heffte_b200 import _lib                      # noqa: E402
from heffte_b200._lib import b200_fft1d_desc, b200_line_geom   # noqa: E402
from oracle import heffte_oracle as O             # noqa: E402
from tests.helpers import TOL, line_geometry      # noqa: E402


def main():
    lib = _lib.load()
    before = lib.b200_launch_count()
    worst = 0.0
    for prec in (1, 0):
        cdt = np.complex128 if prec == 1 else np.complex64
        for shape, dim in (((16, 512, 3), 1), ((32, 2, 512), 2), ((8, 512, 1), 1), ((24, 512, 2), 1)):
            box = O.Box((0, 0, 0), tuple(v - 1 for v in shape))
            g, ca, cb = line_geometry(box, dim)
            d = b200_fft1d_desc(prec, 0, box.size[dim], ca, cb, b200_line_geom(*g), b200_line_geom(*g))
            plan = ctypes.c_void_p()
            assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(plan)) == 0, lib.b200_last_error()
            rng = np.random.default_rng(7)
            for batch in (1, 3):
                x = (rng.random(batch * box.count()) + 1j * rng.random(batch * box.count())).astype(cdt)
                for direction in (0, 1):
                    xin = torch.from_numpy(x).cuda()
                    out = torch.zeros_like(xin)
                    step = box.count() * x.itemsize
                    rc = lib.b200_fft1d_execute_batch(plan, direction, ctypes.c_void_p(xin.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_double(1.0), None,
                                                      batch, step, step)
                    assert rc == 0, lib.b200_last_error()
                    torch.cuda.synchronize()
                    got = out.cpu().numpy()
                    for b in range(batch):
                        seg = x[b * box.count():(b + 1) * box.count()]
                        ref = O.exec1d_c2c(seg, box, dim, backward=bool(direction))
                        err = O.rel_l2(got[b * box.count():(b + 1) * box.count()], ref)
                        worst = max(worst, err / TOL[prec])
                        assert err <= TOL[prec], (prec, shape, dim, batch, direction, err)
            lib.b200_fft1d_destroy(plan)
    print("gpu_tma_worker: ok, worst error / tolerance %.3g, launches %d" % (worst, lib.b200_launch_count() - before))


if __name__ == "__main__":
    main()
