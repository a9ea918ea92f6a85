"""Dev tool (GPU box): wall time of repeated one-shot mwf_wfa_exact calls on the 150 kb pair (CIGAR)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
t, q = synth.make_pair(150000, 0.038, 900000)
for kw in ({"flag": 1}, {}, {"flag": 1, "step": 5000}):
    o = mw.opt_init(**kw)
    for rep in range(4):
        t0 = time.perf_counter()
        r = mw.wfa_exact(o, t, q)
        print(kw, "rep", rep, "%.1f ms" % ((time.perf_counter() - t0) * 1e3), r[:3], flush=True)
