"""
TEST INFRASTRUCTURE.  Builds tests/emul/_build/libheffte_b200_emul.so: the COMPLETE product library (same sources, same C ABI)
compiled for the host with B200_HOST_EMULATION -- kernels run thread-by-thread on the CPU (tests/emul/cuda_emul.h), the
CUDA runtime is a synchronous stand-in.  Lets the CPU-only test-suite run whole multi-rank plans (ranks as host threads),
including the peer-memory mode, before GPU time is spent.  Never shipped, never loaded by the product.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "heffte_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libheffte_b200_emul.so")
UNITS = ["fft1d.cu", "pack.cu", "plan_logic.cpp", "comm.cpp", "transform.cpp", "capi.cpp"] + \
        ["fft_inst_%s_%s_%s.cu" % (f, t, m) for f in ("strided", "contig", "real", "sreal", "pair", "conv") for t in ("f32", "f64") for m in ("direct", "scatter")] + ["fft_inst_sreal2_f32_direct.cu", "fft_inst_sreal2_f64_direct.cu"]


def build(force=False):
    deps = [os.path.join(CSRC, n) for n in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emul.h"), os.path.join(ROOT, "include", "heffte_b200.h"),
                                                                os.path.join(ROOT, "include", "heffte_b200_kernels.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    objdir = os.path.join(HERE, "_build", "obj")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O1", "-std=c++20", "-fPIC", "-pthread", "-DB200_HOST_EMULATION", "-I", CSRC, "-I", HERE]
    jobs, objects = [], []
    for unit in UNITS:
        obj = os.path.join(objdir, unit.rsplit(".", 1)[0] + ".o")
        objects.append(obj)
        jobs.append((unit, subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", os.path.join(CSRC, unit), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for unit, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("emulation build failed on %s:\n%s" % (unit, out[-4000:]))
    link = subprocess.run(["g++", "-shared", "-pthread", "-Wl,-Bsymbolic", "-o", OUT] + objects + ["-ldl"], capture_output=True, text=True)
    if link.returncode != 0:
        raise RuntimeError("emulation link failed:\n" + link.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
