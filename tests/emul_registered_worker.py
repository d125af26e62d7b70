"""
TEST INFRASTRUCTURE.  The contract of register_buffer on the emulated library (two thread-ranks): a transform whose output is
the registered array on every rank stores straight into it; with HEFFTE_B200_CHECK_REGISTERED=1 a call in which only some
ranks pass their registered array is refused on every rank.  Run by tests/test_emul_distributed.py in a subprocess.
"""
import ctypes
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    os.environ["HEFFTE_B200_CHECK_REGISTERED"] = "1"
    os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
    from tests.emul.build_emul_library import build
    from heffte_b200 import _lib
    _lib.LIB_PATH = build()
    import numpy as np
    import heffte_b200 as hf
    from oracle import heffte_oracle as O
    from tests.multi_rank_worker import bricks, to_h

    n = (16, 16, 16)
    world = O.world_box(n)
    boxes = bricks(world, (1, 1, 2))
    rng = np.random.default_rng(5)
    x = (rng.random(world.count()) + 1j * rng.random(world.count())).astype(np.complex128)
    expect = O.fft3d_forward(x, n, "c2c", scaling="none")
    comms = hf.comm_threads(2)
    results = [None, None]

    def body(rank):
        fft = hf.fft3d(hf.backend.b200, to_h(boxes[rank]), to_h(boxes[rank]), comms[rank], hf.plan_options(hf.backend.b200))
        lib = _lib.load()
        d = O.get_subbox(world, boxes[rank], x).copy()
        other = np.zeros_like(d)
        registered = fft.register_buffer(d)

        def run(src, dst):
            return lib.heffte_execute(fft.plan, 1, 0, 1, ctypes.c_void_p(src.ctypes.data), ctypes.c_void_p(dst.ctypes.data), None, 0)
        bad = run(d, d if rank == 0 else other)            # rank 1 breaks the contract: refused everywhere, nothing written
        message = _lib.last_error()
        good = run(d, d)                                   # in place into the registered arrays
        err = O.rel_l2(d, O.get_subbox(world, boxes[rank], expect))
        fft.unregister_buffer(d)
        results[rank] = (registered, bad, message, good, err)

    threads = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    [t.start() for t in threads]
    [t.join(timeout=300) for t in threads]
    for rank, r in enumerate(results):
        assert r is not None, "rank %d hung" % rank
        registered, bad, message, good, err = r
        assert registered, "the array was not registered"
        assert bad == 1 and "registered" in message, (bad, message)
        assert good == 0 and err < 1e-12, (good, err)
    print("emul_registered_worker: ok")


if __name__ == "__main__":
    main()
