"""Dev tool (GPU box): one pass over the single 150 kb pair (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
kw = {"flag": 1} if os.environ.get("TB") else {}
t, q = synth.make_pair(150000, 0.038, 900000)
with mw.Batch(mw.opt_init(**kw), [(t, q)]) as b:
    b.upload()
    for _ in range(int(os.environ.get("REPS", "1"))):
        b.run(); b.wait()
    r = b.fetch()
    print("kernel_ms %.2f launches %d" % (b.kernel_ms, b.launches), r[0][:3])
