"""
TEST INFRASTRUCTURE.  Whole multi-rank plans on the CPU: loads the EMULATED build of the product library
(tests/emul/build_emul_library.py: same sources and C ABI, kernels executed thread-by-thread on the host, synchronous
stand-in for the CUDA runtime) and runs the reference-style distributed test matrix with the ranks as host threads, in the
peer-memory mode (fused FFT + reshape, scatter maps into the other ranks' buffers) and in the pack / exchange / unpack mode.
Run by tests/test_emul_distributed.py in a subprocess:   python tests/emul_worker.py <nranks> <peer|exchange> [stride]
"""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nranks, mode = int(sys.argv[1]), sys.argv[2]
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if mode == "exchange":
        os.environ["HEFFTE_B200_DISABLE_P2P"] = "1"
    else:
        os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
    # emulated launches are synchronous, so the two-stream overlap of a local and a fused transform (kept off for ranks that
    # share a real GPU) is safe to exercise here
    os.environ["HEFFTE_B200_OVERLAP_ON_SHARED_DEVICE"] = "1"
    from tests.emul.build_emul_library import build
    from heffte_b200 import _lib
    _lib.LIB_PATH = build()          # the emulated library stands in for libheffte_b200.so in THIS process only
    import heffte_b200 as hf
    from tests.multi_rank_worker import HostArrays, configs, grids_for, run_config

    todo = [(c, 1) for c in configs(nranks, quick=True, subcomm=True) if max(c["n"]) <= 32][::stride]
    gin, gout = grids_for(nranks)[0]
    todo.append((dict(kind="c2c", n=(8, 9, 10), prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 2))
    # power-of-two lines of 32 points: the real-data fast kernels (contiguous and strided, plain and fused-reshape stores)
    for kind, extra, prec in (("r2c", dict(r2c_dir=0), 1), ("r2c", dict(r2c_dir=2), 0), ("cos", {}, 1), ("sin", {}, 0)):
        if nranks in (1, 2, 8):
            todo.append((dict(kind=kind, n=(32, 32, 32), prec=prec, reorder=(kind in ("cos", "sin")), pencils=True, alg=0, gin=gin, gout=gout,
                              order_out=(0, 1, 2), **extra), 1))
    # lengths with factors 3 and 5: the mixed-radix kernels inside whole plans (contiguous, strided, fused reshapes, real engine)
    if nranks in (1, 2) and stride != 6:
        todo.append((dict(kind="c2c", n=(96, 80, 6), prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 1))
        todo.append((dict(kind="r2c", n=(160, 96, 4), prec=0, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2), r2c_dir=0), 1))
        todo.append((dict(kind="cos", n=(160, 192, 4), prec=1, reorder=True, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 1))
    # 128-point lines along the two fast axes: the paired kernel (a local transform inside the persistent kernel of the fused
    # stage behind it; on one rank the single-precision pair through the L2 cache), batched
    if nranks in (1, 2) and stride != 6:
        todo.append((dict(kind="c2c", n=(128, 128, 4), prec=0, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 1))
        todo.append((dict(kind="c2c", n=(128, 128, 4), prec=1, reorder=False, pencils=False, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 2))
    comms = hf.comm_threads(nranks)
    gate = threading.Barrier(nranks)
    failures = [None] * nranks
    done = [0] * nranks
    stop = threading.Event()

    def body(rank):
        for index, (c, batch) in enumerate(todo):
            try:
                run_config(hf, None, comms[rank], rank, c, batch, expect_peer=(mode == "peer"), arrays=HostArrays())
                done[rank] += 1
            except Exception as e:  # noqa: BLE001
                failures[rank] = "config %d: %r" % (index, e)
                stop.set()
            gate.wait()
            if stop.is_set():
                break
            gate.wait()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1500)
    if any(t.is_alive() for t in threads):
        print("emul_worker: hung; failures:", [f for f in failures if f], "done:", done, flush=True)
        os._exit(3)
    if any(failures):
        print("emul_worker: FAILED", [f for f in failures if f], flush=True)
        sys.exit(1)
    print("emul_worker: ranks=%d mode=%s configs=%d ok" % (nranks, mode, min(done)), flush=True)


if __name__ == "__main__":
    main()
