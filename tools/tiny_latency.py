"""Dev tool (GPU box): wall time of repeated one-shot calls -- a 72 bp pair (streaming kernel) and the 150 kb pair (tile engine), CIGAR."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
o = mw.opt_init(flag=1)
t, q = synth.make_pair(150000, 0.038, 900000)
for rnd in range(2):
    tt = []
    for _ in range(200):
        t0 = time.perf_counter(); mw.wfa_exact(o, b"ACGTACGTACGTTTGACA" * 4, b"ACGTACGAACGTTTGACA" * 4); tt.append(time.perf_counter() - t0)
    tt_us = sorted(x * 1e6 for x in tt)
    print("tiny: min %.0f p10 %.0f median %.0f p90 %.0f max %.0f us" % (tt_us[0], tt_us[20], tt_us[100], tt_us[180], tt_us[-1]), flush=True)
    bb = []
    for _ in range(6):
        t0 = time.perf_counter(); mw.wfa_exact(o, t, q); bb.append(time.perf_counter() - t0)
    print("150 kb: " + " ".join("%.1f" % (x * 1e3) for x in bb) + " ms", flush=True)
# host jitter check: a fixed pure-Python loop, no GPU, no library
jj = []
for _ in range(300):
    t0 = time.perf_counter(); x = 0
    for i in range(5000): x += i
    jj.append((time.perf_counter() - t0) * 1e6)
jj.sort()
print("pure-python loop: min %.0f median %.0f p90 %.0f max %.0f us; load average %s" % (jj[0], jj[150], jj[270], jj[-1], open("/proc/loadavg").read().strip()), flush=True)
