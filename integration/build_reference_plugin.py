#!/usr/bin/env python
"""
Builds the REFERENCE's own test and benchmark programs against include/heffte_backend_b200.h -- the template plug-in
boundary of SURVEY.md section 8(b) -- so that heffte::fft3d<backend::b200> (reference templates, reference reshapes over
the thread-ranks MPI stand-in, b200 executors / packers / scaling from libheffte_b200.so) is compiled and run unchanged.

The reference tree is read-only and is never copied into this repository: the script makes a SCRATCH copy of
include/ src/ test/ benchmarks/ under a temporary directory, applies the reference-side edits a maintainer would make to
add the backend (listed in EDITS below and in INTEGRATION.md section 3 -- each one a one-line insertion or a copy of an
`#ifdef Heffte_ENABLE_CUDA` block with the CUDA names replaced), compiles with plain g++ (the backend header needs no CUDA
header: it calls the C ABI), and leaves ONLY the binaries in integration/_build/ (git-ignored; they travel to the GPU box
with the snapshot like oracle/_ref).  Runs only where /root/reference exists.

    python integration/build_reference_plugin.py [--keep] [--only name ...]
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HEFFTE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_build")
SHIM = os.path.join(ROOT, "oracle", "mpi_shim")

CONFIG = """#ifndef HEFFTE_CONFIG_H
#define HEFFTE_CONFIG_H
/* what the reference's CMake would generate from include/heffte_config.cmake.h with -DHeffte_ENABLE_B200=ON */
#define Heffte_VERSION_MAJOR 2
#define Heffte_VERSION_MINOR 4
#define Heffte_VERSION_PATCH 1
#define Heffte_GIT_HASH "b200-plugin-build"
#define Heffte_ENABLE_B200
#define Heffte_ENABLE_GPU
#endif
"""

TEST_COMMON_BLOCK = """#ifdef Heffte_ENABLE_B200
using gpu_backend = heffte::backend::b200;

heffte::b200::stream_t make_stream(backend::b200){
    void *result = nullptr;
    heffte::b200::check_error(b200_stream_create(&result), "b200_stream_create()");
    return static_cast<heffte::b200::stream_t>(result);
}
void sync_stream(heffte::b200::stream_t stream){ b200_stream_synchronize(stream); b200_stream_synchronize(nullptr); }
void free_stream(heffte::b200::stream_t stream){ b200_stream_destroy(stream); }
#endif
"""


def duplicate_cuda_blocks(text):
    """after every `#ifdef Heffte_ENABLE_CUDA ... #endif` block append the same block for the b200 backend"""
    lines = text.split("\n")
    out, i = [], 0
    while i < len(lines):
        if re.match(r"\s*#\s*ifdef\s+Heffte_ENABLE_CUDA\s*$", lines[i]):
            depth, j = 1, i + 1
            while j < len(lines) and depth > 0:
                if re.match(r"\s*#\s*if", lines[j]):
                    depth += 1
                elif re.match(r"\s*#\s*endif", lines[j]):
                    depth -= 1
                j += 1
            block = lines[i:j]
            out.extend(block)
            body = "\n".join(block)
            # blocks that call the CUDA runtime directly (includes, managed memory, stream creation) have hand-written twins
            if not re.search(r"#include\s*<cu|cudaMalloc|cudaStream|cudaMemcpy", body):
                twin = body.replace("Heffte_ENABLE_CUDA", "Heffte_ENABLE_B200").replace("cufft", "b200")
                out.extend(twin.split("\n"))
            i = j
        else:
            out.append(lines[i])
            i += 1
    return "\n".join(out)


def insert_after(text, anchor, addition):
    pos = text.find(anchor)
    if pos < 0:
        raise RuntimeError("anchor not found: %r" % anchor)
    end = text.find("\n", pos + len(anchor) - 1) + 1
    return text[:end] + addition + text[end:]


def edit(path, fn):
    with open(path) as f:
        text = f.read()
    new = fn(text)
    if new == text:
        raise RuntimeError("edit did not change %s" % path)
    with open(path, "w") as f:
        f.write(new)


def patch_tree(scratch):
    inc, test, bench = (os.path.join(scratch, d) for d in ("include", "test", "benchmarks"))
    with open(os.path.join(inc, "heffte_config.h"), "w") as f:
        f.write(CONFIG)
    # EDITS
    # 1. include/heffte_backends.h: the new backend header next to the CUDA one
    edit(os.path.join(inc, "heffte_backends.h"), lambda t: insert_after(t, '#include "heffte_backend_cuda.h"', '#include "heffte_backend_b200.h"\n'))
    # 2. include/heffte_fft3d.h:643: the type-I cosine scaling names its backends one by one
    edit(os.path.join(inc, "heffte_fft3d.h"), lambda t: t.replace(
        "std::is_same<backend_tag, backend::cufft_cos1>::value or", "std::is_same<backend_tag, backend::cufft_cos1>::value or std::is_same<backend_tag, backend::b200_cos1>::value or", 1))
    # 3. test/test_common.h: stream helpers of the tests, then the generic twin of every CUDA block of the tests and benchmarks
    edit(os.path.join(test, "test_common.h"), lambda t: insert_after(t, "void free_stream(cudaStream_t stream){ cudaStreamDestroy(stream); }\n#endif", TEST_COMMON_BLOCK))
    # the ranks of these programs are host threads of one process (oracle/mpi_shim): the per-process bookkeeping of the test
    # harness becomes per-thread (every rank assigns the name of the running test: a data race on one std::string otherwise)
    edit(os.path.join(test, "test_common.h"), lambda t: t.replace("std::string heffte_test_name;", "thread_local std::string heffte_test_name;", 1)
         .replace("bool heffte_test_pass  = true;", "thread_local bool heffte_test_pass  = true;", 1))
    for folder in (test, bench):
        for name in sorted(os.listdir(folder)):
            if name.endswith((".cpp", ".h")):
                path = os.path.join(folder, name)
                with open(path) as f:
                    text = f.read()
                new = duplicate_cuda_blocks(text)
                # 4. tests that pick their golden vectors by naming the backends one by one: the b200 tag next to the cufft tag
                new = re.sub(r"std::is_same<(\w+), backend::cufft(_\w+)?>::value(?! or std::is_same<\w+, backend::b200)",
                             lambda m: "std::is_same<%s, backend::cufft%s>::value or std::is_same<%s, backend::b200%s>::value" % (
                                 m.group(1), m.group(2) or "", m.group(1), m.group(2) or ""), new)
                if new != text:
                    with open(path, "w") as f:
                        f.write(new)


PROGRAMS = {
    # name: (source relative to the scratch tree, extra defines)
    "test_units_nompi": ("test/test_units_nompi.cpp", []),
    "test_fft3d_np1": ("test/test_fft3d_np1.cpp", []),
    "test_fft3d_np2": ("test/test_fft3d_np2.cpp", []),
    "test_fft3d_np4": ("test/test_fft3d_np4.cpp", []),
    "test_fft3d_np8": ("test/test_fft3d_np8.cpp", []),
    "test_fft3d_r2c": ("test/test_fft3d_r2c.cpp", []),
    "test_cos": ("test/test_cos.cpp", []),
    "test_reshape3d": ("test/test_reshape3d.cpp", []),
    "test_streams": ("test/test_streams.cpp", []),
    "test_longlong": ("test/test_longlong.cpp", []),
    "test_subcomm": ("test/test_subcomm.cpp", []),
    "speed3d_c2c": ("benchmarks/speed3d_c2c.cpp", []),
    "speed3d_r2c": ("benchmarks/speed3d_r2c.cpp", []),
    "speed3d_r2r": ("benchmarks/speed3d_r2r.cpp", []),
}


def run(cmd):
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), (out.stdout + out.stderr)[-6000:]))
    return out


def build(only=None, keep=False, verbose=False, emulated=False):
    """emulated=True links the programs against the CPU-emulated library of tests/emul (integration/_build/emul/): the
    reference's tests then exercise this header and the kernel source on the CPU-only development container"""
    if not os.path.exists(os.path.join(REF, "include", "heffte.h")):
        return False
    if emulated:
        sys.path.insert(0, ROOT)
        from tests.emul.build_emul_library import build as build_emulated
        lib = build_emulated()
        out_dir = os.path.join(OUT, "emul")
    else:
        from heffte_b200 import build as product
        lib = product.build_library()
        out_dir = OUT
    os.makedirs(out_dir, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="heffte_b200_plugin_")
    try:
        for d in ("include", "src", "test", "benchmarks"):
            shutil.copytree(os.path.join(REF, d), os.path.join(scratch, d))
        patch_tree(scratch)
        flags = ["-O2", "-std=c++14", "-pthread", "-Wno-deprecated-declarations", "-I", os.path.join(scratch, "include"), "-I", os.path.join(ROOT, "include"),
                 "-I", SHIM, "-I", os.path.join(scratch, "test"), "-I", os.path.join(scratch, "benchmarks")]
        flags += os.environ.get("HEFFTE_B200_PLUGIN_CXXFLAGS", "").split()      # e.g. -fsanitize=address -g when chasing a bug
        objects = []
        jobs = []
        for src in ("src/heffte_plan_logic.cpp", "src/heffte_reshape3d.cpp", "src/heffte_compute_transform.cpp"):
            obj = os.path.join(scratch, os.path.basename(src) + ".o")
            jobs.append(subprocess.Popen(["g++"] + flags + ["-c", os.path.join(scratch, src), "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            objects.append(obj)
        shim_obj = os.path.join(scratch, "mpi_threads.o")
        jobs.append(subprocess.Popen(["g++"] + flags + ["-c", os.path.join(SHIM, "mpi_threads.cpp"), "-o", shim_obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        objects.append(shim_obj)
        for j in jobs:
            text, _ = j.communicate()
            if j.returncode != 0:
                raise RuntimeError("reference source failed to compile against the b200 backend:\n" + text[-6000:])
        names = [n for n in PROGRAMS if (only is None or n in only)]
        if emulated:
            link = [lib, "-Wl,-rpath," + os.path.dirname(lib), "-ldl"]
        else:
            link = ["-L", os.path.dirname(lib), "-lheffte_b200", "-Wl,-rpath,$ORIGIN/../../heffte_b200/lib", "-ldl"]
        pending = []
        for name in names:
            src, defines = PROGRAMS[name]
            obj = os.path.join(scratch, name + ".o")
            cmd = ["g++"] + flags + defines + ["-Dmain=shim_rank_main", "-c", os.path.join(scratch, src), "-o", obj]
            pending.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        failed = []
        for name, obj, proc in pending:
            text, _ = proc.communicate()
            if proc.returncode != 0:
                failed.append((name, text[-5000:]))
                continue
            run(["g++"] + flags + [os.path.join(SHIM, "shim_main.cpp"), obj] + objects + link + ["-o", os.path.join(out_dir, name)])
            if verbose:
                print("built", os.path.join(out_dir, name))
        if failed:
            raise RuntimeError("\n".join("%s:\n%s" % f for f in failed))
        with open(os.path.join(out_dir, ".done"), "w") as f:
            f.write("\n".join(names) + "\n")
        return True
    finally:
        if keep:
            print("scratch tree kept at", scratch)
        else:
            shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    only = None
    if "--only" in sys.argv:
        only = sys.argv[sys.argv.index("--only") + 1:]
    if "--only" in sys.argv:
        only = [a for a in only if not a.startswith("--")]
    ok = build(only=only, keep="--keep" in sys.argv, verbose=True, emulated="--emulated" in sys.argv)
    print("built" if ok else "reference tree not present: nothing built")
