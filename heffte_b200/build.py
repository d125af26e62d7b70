"""
Builds libheffte_b200.so IN-TREE (heffte_b200/lib/) with nvcc for sm_100a.  nvcc cross-compiles without a GPU, so
this runs in the CPU-only container; the resulting .so travels to the GPU box with the repo snapshot.
Usage: python -m heffte_b200.build [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBNAME = "libheffte_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
CUDA_SOURCES = ["fft1d.cu", "pack.cu"] + ["fft_inst_%s_%s_%s.cu" % (f, t, m) for f in ("strided", "contig", "real", "sreal", "pair", "conv") for t in ("f32", "f64")
                                           for m in ("direct", "scatter")] + ["fft_inst_sreal2_f32_direct.cu", "fft_inst_sreal2_f64_direct.cu"]
HOST_SOURCES = ["plan_logic.cpp", "comm.cpp", "transform.cpp", "capi.cpp"]


def _sources():
    names = sorted(os.listdir(CSRC)) + ["../../include/heffte_b200.h", "../../include/heffte_b200_kernels.h"]
    return [os.path.join(CSRC, n) for n in names if n.endswith((".cu", ".cuh", ".cpp", ".h", ".inc"))]


def _fingerprint():
    h = hashlib.sha256()
    for path in _sources():
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode())   # not the absolute path: the tree is copied to the GPU box
            h.update(f.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def library_path():
    return os.path.join(LIBDIR, LIBNAME)


def build_library(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, ".fingerprint")
    fp = _fingerprint()
    if not force and os.path.exists(library_path()) and os.path.exists(stamp) and open(stamp).read() == fp:
        return library_path()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    objects = []
    jobs = []
    for src in CUDA_SOURCES + HOST_SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + ARCH + COMMON + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objects.append(obj)
    failed = False
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("libheffte_b200.so: compilation failed")
    link = [NVCC] + ARCH + ["-shared", "-o", library_path()] + objects + ["-ldl", "-lpthread"]
    out = subprocess.run(link, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libheffte_b200.so: link failed:\n" + out.stdout + out.stderr)
    with open(stamp, "w") as f:
        f.write(fp)
    # every source is recompiled whenever the fingerprint changes: the objects are of no further use, and the tree (objects
    # included) is what travels to the GPU box
    for obj in objects:
        try:
            os.remove(obj)
        except OSError:
            pass
    return library_path()


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
