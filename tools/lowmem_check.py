"""Dev tool (GPU box): low-memory mode (-cp5000) on a large pair vs the unmodified reference on the host, with timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
from oracle import orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0097
step = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
t, q = synth.make_pair(n, p, 424242)
kw = {"flag": 1, "step": step}
t0 = time.perf_counter()
got = mw.wfa_exact(mw.opt_init(**kw), t, q)
t1 = time.perf_counter()
print("gpu low-mem: s=%d n_cigar=%d n_iter=%d  %.3f s" % (got[0], got[1], got[2], t1 - t0), flush=True)
sc = mw.wfa_exact(mw.opt_init(), t, q)
t2 = time.perf_counter()
print("gpu score-only: s=%d n_iter=%d  %.3f s" % (sc[0], sc[2], t2 - t1), flush=True)
if os.environ.get("HIGHMEM"):
    hm = mw.wfa_exact(mw.opt_init(flag=1), t, q)
    t3 = time.perf_counter()
    print("gpu high-mem: s=%d n_cigar=%d n_iter=%d  %.3f s  same cigar as low-mem: %s" % (hm[0], hm[1], hm[2], t3 - t2, hm[3] == got[3]), flush=True)
if not os.environ.get("NOCPU"):
    t4 = time.perf_counter()
    want = orc.reference_exact(orc.make_opt(**kw), t, q)
    print("cpu reference: s=%d n_cigar=%d n_iter=%d  %.3f s  equal: %s" % (want[0], want[1], want[2], time.perf_counter() - t4, want == got), flush=True)
