"""Dev tool (GPU box): one score-only pass over a single large pair (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
n = int(os.environ.get("LEN", "5000000")); p = float(os.environ.get("DIV", "0.0097"))
kw = {"flag": 1} if os.environ.get("TB") else {}
t, q = synth.make_pair(n, p, 424242)
with mw.Batch(mw.opt_init(**kw), [(t, q)]) as b:
    b.upload()
    for _ in range(int(os.environ.get("REPS", "1"))):
        b.run(); b.wait()
    r = b.fetch()[0]
    print("kernel_ms %.2f launches %d" % (b.kernel_ms, b.launches), r[:3])
