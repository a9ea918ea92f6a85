"""Dev tool (GPU box): one mwf_wfa_chain call on a 5 Mb / 3 % pair with CIGAR (for the ncu launch list of the k-mer front end
and the gap-fill batch).  Usage: [N=5000000] python tools/prof_chain.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
t, q = synth.make_pair(int(os.environ.get("N", "5000000")), 0.03, 424242)
o = mw.opt_init(flag=1, step=5000)
t0 = time.perf_counter()
r = mw.wfa_chain(o, t, q)
print("chain s=%d n_cigar=%d %.1f ms" % (r[0], r[1], (time.perf_counter() - t0) * 1e3))
