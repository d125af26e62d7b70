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

    # ---- any length: the composite engine (four-step over sub-plans, Bluestein for lengths that do not split) --------------------
    def run1d(kind, prec, n, lines, direction, x, contiguous):
        rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
        nc = n // 2 + 1
        if contiguous:
            gin, gout = (1, n, 0), (1, nc if kind == 1 else n, 0)
        else:
            gin, gout = (lines, 1, 0), (lines, 1, 0)
        d = b200_fft1d_desc(prec, kind, n, lines, 1, b200_line_geom(*gin), b200_line_geom(*gout))
        p = vp()
        rc = lib.b200_fft1d_create(ctypes.byref(d), ctypes.byref(p))
        assert rc == 0, (n, _lib.last_error())
        name = lib.b200_fft1d_kernel_name(p).decode()
        real_in = (kind == 1 and direction == 0) or kind >= 2
        real_out = (kind == 1 and direction == 1) or kind >= 2
        src = np.ascontiguousarray(x.astype(rdt if real_in else cdt))
        count_out = lines * (nc if (kind == 1 and direction == 0) else n)
        dst = np.zeros(count_out, dtype=rdt if real_out else cdt)
        rc = lib.b200_fft1d_execute(p, direction, vp(src.ctypes.data), vp(dst.ctypes.data), ctypes.c_double(1.0), None)
        assert rc == 0, (n, _lib.last_error())
        lib.b200_fft1d_destroy(p)
        return dst, name

    for prec, n, lines, expect in ((1, 8192, 3, "four-step"), (1, 7168, 2, "four-step"), (0, 7168, 2, "generic"), (1, 1021, 4, "Bluestein"), (1, 10007, 2, "Bluestein"),
                                   (0, 12289, 1, "Bluestein"), (1, 16384, 1, "four-step"), (1, 4099, 2, "Bluestein")):
        tol = 3e-5 if prec == 0 else 2e-11
        for contiguous in (True, False):
            x = rng.random(n * lines) + 1j * rng.random(n * lines)
            view = x.reshape(lines, n) if contiguous else x.reshape(n, lines).T
            for direction in (0, 1):
                ref = np.fft.fft(view, axis=1) if direction == 0 else np.fft.ifft(view, axis=1) * n
                ref = ref.reshape(-1) if contiguous else ref.T.reshape(-1)
                y, name = run1d(0, prec, n, lines, direction, x, contiguous)
                if expect not in name:
                    failures.append("length %d: kernel %s, expected %s" % (n, name, expect))
                err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
                if not err <= tol:
                    failures.append("composite c2c n=%d prec=%d contiguous=%s direction=%d: rel l2 %.3e" % (n, prec, contiguous, direction, err))
    # real transforms of awkward lengths: r2c / c2r of a prime length, DCT-II / DCT-III and DST-II / III of a prime length, long DCT-I
    from oracle import heffte_oracle as O
    for n, lines in ((1021, 3), (8200, 2)):
        box = O.Box((0, 0, 0), (n - 1, lines - 1, 0))
        xr = rng.random(n * lines)
        y, name = run1d(1, 1, n, lines, 0, xr, True)
        ref = O.exec1d_r2c(xr, box, 0)
        if not O.rel_l2(y, ref) <= 2e-11:
            failures.append("composite r2c n=%d (%s): %.3e" % (n, name, O.rel_l2(y, ref)))
        z, _ = run1d(1, 1, n, lines, 1, ref, True)
        if not O.rel_l2(z, O.exec1d_c2r(ref, box, 0)) <= 2e-11:
            failures.append("composite c2r n=%d: %.3e" % (n, O.rel_l2(z, O.exec1d_c2r(ref, box, 0))))
    for kind, kname, n in ((2, "cos", 509), (3, "sin", 509), (2, "cos", 9000), (4, "cos1", 4200)):
        lines = 2
        box = O.Box((0, 0, 0), (n - 1, lines - 1, 0))
        xr = rng.random(n * lines)
        f, name = run1d(kind, 1, n, lines, 0, xr, True)
        if "composite" not in name:
            failures.append("%s n=%d ran on %s" % (kname, n, name))
        if not O.rel_l2(f, O.r2r_forward(xr, box, 0, kname)) <= 1e-10:
            failures.append("composite %s forward n=%d: %.3e" % (kname, n, O.rel_l2(f, O.r2r_forward(xr, box, 0, kname))))
        b, _ = run1d(kind, 1, n, lines, 1, xr, True)
        if not O.rel_l2(b, O.r2r_backward(xr, box, 0, kname)) <= 1e-10:
            failures.append("composite %s backward n=%d: %.3e" % (kname, n, O.rel_l2(b, O.r2r_backward(xr, box, 0, kname))))

    if failures:
        print("emul_kernels_worker: FAILED", failures, flush=True)
        sys.exit(1)
    print("emul_kernels_worker: ok", flush=True)


if __name__ == "__main__":
    main()
