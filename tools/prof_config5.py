"""Dev tool (GPU box): the config-5 pair (5 Mb, 3 %, CIGAR) with the phase times of the segmented traceback."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MWF_B200_BATCH_TIMING"] = "1"
import miniwfa_b200 as mw
from miniwfa_b200 import synth
t, q = synth.make_pair(5000000, float(os.environ.get("DIV", "0.03")), 424242)
o = mw.opt_init(flag=1, step=int(os.environ.get("STEP", "0")))
for rep in range(2):
    t0 = time.perf_counter()
    r = mw.wfa_exact(o, t, q)
    print("rep", rep, "%.3f s" % (time.perf_counter() - t0), r[:3], flush=True)
