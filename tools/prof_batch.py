"""Dev tool (GPU box): one pass of the config-3 batch (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
npairs = int(os.environ.get("NP", "128"))
n = int(os.environ.get("LEN", "100000"))
kw = {"flag": 1} if os.environ.get("TB") else {}
pairs = synth.make_batch(npairs, n, 0.05, 0)
with mw.Batch(mw.opt_init(**kw), pairs) as b:
    b.upload()
    for _ in range(int(os.environ.get("REPS", "1"))):
        b.run(); b.wait()
    r = b.fetch()
    print("kernel_ms %.2f launches %d cells %.4e" % (b.kernel_ms, b.launches, sum(x[2] for x in r)))
