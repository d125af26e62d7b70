"""Developer tool: per-rank NVLink bytes of every reshape of a plan (pure host planning, no GPU).
usage: python tools/plan_traffic.py N [n] [--io-pencils] [--reorder] [--slabs]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heffte_b200 as hf
from heffte_b200 import heffte as H


def traffic(N, n=512, io_pencils=False, reorder=False, pencils=True, elem=16, executed=False):
    world = hf.box3d((0, 0, 0), (n - 1,) * 3)
    if io_pencils:
        g2 = H.make_procgrid(N)
        gin, gout = [1, g2[0], g2[1]], [g2[0], g2[1], 1]
    else:
        gin = gout = H.proc_setup_min_surface(world, N)
    planner = H.execution_plan if executed else H.logic_plan
    shapes, fdir, _ = planner(H.split_world(world, gin), H.split_world(world, gout), use_reorder=reorder, use_pencils=pencils)
    out = []
    for s in range(4):
        ins, outs = shapes[s], shapes[4 + s]
        sent = []
        for r in range(N):
            a, tot = ins[r], 0
            for q in range(N):
                if q == r:
                    continue
                b, ov = outs[q], 1
                for d in range(3):
                    ov *= max(0, min(a[3 + d], b[3 + d]) - max(a[d], b[d]) + 1)
                tot += ov
            sent.append(tot * elem)
        out.append(sent)
    return gin, fdir, out


if __name__ == "__main__":
    N = int(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 512
    for executed in (False, True):
        gin, fdir, t = traffic(N, n, "--io-pencils" in sys.argv, "--reorder" in sys.argv, "--slabs" not in sys.argv, executed=executed)
        print("ranks", N, "grid", gin, "fft_direction", fdir, "EXECUTED plan (no reorder, balanced)" if executed else "REFERENCE plan")
        for s, sent in enumerate(t):
            print("  reshape %d: sent MB per rank %s   max %.1f  mean %.1f" % (s, [round(x / 1e6, 1) for x in sent], max(sent) / 1e6, sum(sent) / len(sent) / 1e6))
        print("  sum of per-reshape max: %.1f MB ; mean: %.1f MB" % (sum(max(s) for s in t) / 1e6, sum(sum(s) / len(s) for s in t) / 1e6))
