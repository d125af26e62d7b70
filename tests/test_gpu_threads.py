"""
Multi-rank plans on ONE GPU: the ranks are host threads of this process (heffte_comm_create_threads), each with its own CUDA
stream, all driving the same device.  Exercises exactly the code the multi-GPU job runs -- plan-time allgather, peer-memory
mode (scatter maps into the other ranks' buffers, stream-ordered peer barriers, fused FFT + reshape kernels) and, with
HEFFTE_B200_DISABLE_P2P=1, the pack / exchange / unpack path -- against the oracle, in the spirit of the reference's
mpirun -np 2/4/6/8/12 tests (test/test_fft3d_np*.cpp, test_fft3d_r2c.cpp, test_cos.cpp, test_subcomm.cpp geometry).
"""
import os
import threading
import time

import pytest

from tests.multi_rank_worker import configs, grids_for, run_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _run_group(nranks, todo, expect_peer, budget_s=240, label=""):
    import heffte_b200 as hf
    comms = hf.comm_threads(nranks)
    gate = threading.Barrier(nranks)
    failures = [None] * nranks
    worst = [0.0] * nranks
    done = [0] * nranks
    stop = threading.Event()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "threads_%d_%s%s.log" % (nranks, "peer" if expect_peer else "exchange", label)), "w")

    def body(rank):
        try:
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for index, (c, batch) in enumerate(todo):
                    try:
                        worst[rank] = max(worst[rank], run_config(hf, torch, comms[rank], rank, c, batch, stream=stream.cuda_stream, expect_peer=expect_peer))
                        done[rank] += 1
                    except Exception as e:  # noqa: BLE001
                        failures[rank] = "config %d %s: %r" % (index, c, e)
                        stop.set()
                    if rank == 0 or failures[rank]:
                        log.write("rank %d config %d %s: %s\n" % (rank, index, c, failures[rank] or "ok"))
                        log.flush()
                    stream.synchronize()
                    gate.wait(timeout=60)    # nobody enters the next collective plan creation after a failure
                    if stop.is_set():
                        break
                    gate.wait(timeout=60)
        except Exception as e:  # noqa: BLE001
            failures[rank] = failures[rank] or "rank %d died: %r" % (rank, e)
            stop.set()
            gate.abort()

    torch.cuda.synchronize()
    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(nranks)]
    deadline = time.time() + budget_s
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=max(1.0, deadline - time.time()))
    log.close()
    if any(t.is_alive() for t in threads):
        # a rank is stuck inside a collective: nothing can be recovered in this process
        print("thread-ranks hung; failures so far:", [f for f in failures if f], "configs done per rank:", done, flush=True)
        import faulthandler
        import sys
        faulthandler.dump_traceback(file=sys.stdout, all_threads=True)
        sys.stdout.flush()
        os._exit(3)
    assert not any(failures), [f for f in failures if f]
    return min(done), max(worst)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_peer_memory_mode(lib, nranks):
    os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
    todo = [(c, 1) for c in configs(nranks, quick=True)]
    gin, gout = grids_for(nranks)[0]
    todo.append((dict(kind="c2c", n=(16, 18, 20), prec=1, reorder=False, pencils=True, alg=0, gin=gin, gout=gout, order_out=(0, 1, 2)), 3))
    done, worst = _run_group(nranks, todo, expect_peer=True)
    assert done == len(todo)


@pytest.mark.parametrize("nranks", [3, 6, 12])
def test_peer_memory_odd_rank_counts(lib, nranks):
    os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
    todo = [(c, 1) for c in configs(nranks, quick=True)][::3]
    # 12 slabs along one axis exceed the 8 cells per axis of a scatter map: such plans take the exchange path by design
    done, _ = _run_group(nranks, todo, expect_peer=True if nranks < 12 else None)
    assert done == len(todo)


@pytest.mark.parametrize("nranks", [2, 4])
def test_exchange_path(lib, nranks):
    os.environ["HEFFTE_B200_DISABLE_P2P"] = "1"
    try:
        todo = [(c, 1) for c in configs(nranks, quick=True)][::2]
        done, _ = _run_group(nranks, todo, expect_peer=False)
        assert done == len(todo)
    finally:
        os.environ.pop("HEFFTE_B200_DISABLE_P2P", None)
