"""GPU box: mwf_wfa_chain / mwf_wfa_auto on large synthetic pairs, wall time per call and the phase times of mwf_chain.c
(MWF_B200_CHAIN_TIMING), device front end against host front end.  Usage: python tools/chain_big.py [n ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniwfa_b200 as mw
from miniwfa_b200 import synth

os.environ["MWF_B200_CHAIN_TIMING"] = "1"
sizes = [int(a) for a in sys.argv[1:]] or [1000000, 5000000]
for n in sizes:
    t, q = synth.make_pair(n, 0.03, 0)
    res = {}
    for front in ("gpu", "gpu", "host"):
        os.environ["MWF_B200_CHAIN_FRONT"] = front
        for flag in (0, 1):
            o = mw.opt_init(flag=flag, step=5000 if flag else 0)
            t0 = time.perf_counter()
            r = mw.wfa_chain(o, t, q)
            dt = time.perf_counter() - t0
            print("chain n=%d front=%s flag=%d: s=%d n_cigar=%d %.1f ms" % (n, front, flag, r[0], r[1], dt * 1e3), flush=True)
            res.setdefault(flag, []).append((r[0], r[1], r[3]))
    for flag, v in res.items():
        assert all(x == v[0] for x in v), "front ends disagree"
    del os.environ["MWF_B200_CHAIN_FRONT"]
    o = mw.opt_init(flag=1)
    t0 = time.perf_counter()
    r = mw.wfa_auto(o, t, q)
    print("auto  n=%d flag=1: s=%d n_cigar=%d %.1f ms" % (n, r[0], r[1], (time.perf_counter() - t0) * 1e3), flush=True)
    # the first leg of mwf_wfa_auto alone: exact with a budget of 1e8 cells, phase by phase
    for rep in range(2):
        o = mw.opt_init(flag=1, max_iter=100000000)
        tm = [time.perf_counter()]
        b = mw.Batch(o, [(t, q)]); tm.append(time.perf_counter())
        b.upload(); tm.append(time.perf_counter())
        b.run(); b.wait(); tm.append(time.perf_counter())
        rr = b.fetch(); tm.append(time.perf_counter())
        kms = b.kernel_ms
        b.close(); tm.append(time.perf_counter())
        print("exact, max_iter=1e8, n=%d rep %d: create %.1f upload %.1f run %.1f (kernel %.1f) fetch %.1f destroy %.1f ms -> s=%d n_iter=%d"
              % ((n, rep) + tuple((tm[i + 1] - tm[i]) * 1e3 for i in range(3)) + (kms,) + tuple((tm[i + 1] - tm[i]) * 1e3 for i in (3, 4)) + (rr[0][0], rr[0][2])), flush=True)
