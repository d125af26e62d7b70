"""
TEST INFRASTRUCTURE.  The paired kernel (fft_pair_kernel) and the fused spectral-operator kernel (fft_strided_conv_kernel)
through the C ABI of the EMULATED library (kernel source executed thread by thread on the CPU), against numpy.
Run by tests/test_emul_pair_conv.py in a subprocess.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests.emul.build_emul_library import build
    from heffte_b200 import _lib
    _lib.LIB_PATH = build()
    lib = _lib.load()
    from heffte_b200._lib import b200_fft1d_desc, b200_line_geom
    vp = ctypes.c_void_p
    failures = []

    def plan(prec, n, ca, cb, geom):
        d = b200_fft1d_desc(prec, 0, n, ca, cb, b200_line_geom(*geom), b200_line_geom(*geom))
        p = vp()
        assert lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(p)) == 0, _lib.last_error()
        return p

    rng = np.random.default_rng(5)
    # ---- paired kernel: box n x n x planes, contiguous axis + middle axis, both orders, both directions, batch 2 --------------
    for prec, n, planes in ((1, 128, 3), (0, 128, 2), (1, 256, 2), (0, 256, 1)):
        cdt = np.complex64 if prec == 0 else np.complex128
        tol = 2e-5 if prec == 0 else 1e-12
        pc = plan(prec, n, n, planes, (1, n, n * n))          # along the contiguous axis
        ps = plan(prec, n, n, planes, (n, 1, n * n))          # along the middle axis
        assert lib.b200_fft1d_pairable(pc, ps) == 1 and lib.b200_fft1d_pairable(ps, pc) == 1
        batch = 2
        count = n * n * planes
        x = (rng.random(batch * count) + 1j * rng.random(batch * count)).astype(cdt)
        counters = np.zeros(batch * planes + 8, dtype=np.uint32)
        for direction in (0, 1):
            ref = x.reshape(batch, planes, n, n).astype(np.complex128)
            ref = np.fft.fft2(ref, axes=(2, 3)) if direction == 0 else np.fft.ifft2(ref, axes=(2, 3)) * (n * n)
            ref = ref.reshape(-1) * 0.5
            for first, second in ((pc, ps), (ps, pc)):
                src = x.copy()
                mid = np.zeros_like(x)
                step = count * x.itemsize
                rc = lib.b200_fft1d_execute_pair(first, second, direction, vp(src.ctypes.data), vp(mid.ctypes.data), None, ctypes.c_double(0.5),
                                                 vp(counters.ctypes.data), 2, None, batch, step, step, 0, 0, 0)
                if rc != 0:
                    failures.append("pair n=%d prec=%d: rc %d %s" % (n, prec, rc, _lib.last_error()))
                    continue
                err = np.linalg.norm(mid - ref) / np.linalg.norm(ref)
                if not err <= tol:
                    failures.append("pair n=%d prec=%d direction=%d contig_first=%s: rel l2 %.3e" % (n, prec, direction, first is pc, err))
                if not np.array_equal(src, x):
                    failures.append("pair n=%d: the out-of-place input was modified" % n)
        for p in (pc, ps):
            lib.b200_fft1d_destroy(p)
        # lengths without a pair shape, and the two slow axes, are not pairable
        p64 = plan(prec, 64, 64, 1, (1, 64, 64 * 64))
        q64 = plan(prec, 64, 64, 1, (64, 1, 64 * 64))
        assert lib.b200_fft1d_pairable(p64, q64) == 0
        lib.b200_fft1d_destroy(p64); lib.b200_fft1d_destroy(q64)

    # ---- spectral operator along a strided axis: 1, 2, 3 and 4 passes; self-product and caller multiplier; batch -------------------
    for prec, n, lines in ((1, 16, 40), (0, 64, 24), (1, 128, 10), (0, 512, 20), (1, 2048, 6)):
        cdt = np.complex64 if prec == 0 else np.complex128
        tol = 5e-5 if prec == 0 else 1e-11
        p = plan(prec, n, lines, 1, (lines, 1, 0))            # lines adjacent in memory: x[k * lines + line]
        assert lib.b200_fft1d_convolvable(p) == 1
        batch = 2
        count = n * lines
        x = (rng.random(batch * count) + 1j * rng.random(batch * count)).astype(cdt)
        m = (rng.random(count) + 1j * rng.random(count)).astype(cdt)
        scale = 1.0 / n
        spec = np.fft.fft(x.reshape(batch, n, lines).astype(np.complex128), axis=1) * scale
        for mult in (None, m):
            prod = spec * (spec if mult is None else mult.reshape(1, n, lines).astype(np.complex128))
            ref = (np.fft.ifft(prod, axis=1) * n).reshape(-1)
            out = np.zeros_like(x)
            step = count * x.itemsize
            rc = lib.b200_fft1d_execute_convolve(p, vp(x.ctypes.data), vp(out.ctypes.data), None, None if mult is None else vp(mult.ctypes.data),
                                                 ctypes.c_double(scale), None, batch, step, step, 0, 0, 0)
            if rc != 0:
                failures.append("convolve n=%d: rc %d %s" % (n, rc, _lib.last_error()))
                continue
            err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
            if not err <= tol:
                failures.append("convolve n=%d prec=%d multiplier=%s: rel l2 %.3e" % (n, prec, mult is not None, err))
        lib.b200_fft1d_destroy(p)
    pc = plan(1, 64, 8, 1, (1, 64, 0))
    assert lib.b200_fft1d_convolvable(pc) == 0       # contiguous axis: the plan falls back to three launches
    lib.b200_fft1d_destroy(pc)

    if failures:
        print("emul_kernels_worker: FAILED", failures, flush=True)
        sys.exit(1)
    print("emul_kernels_worker: ok", flush=True)


if __name__ == "__main__":
    main()
