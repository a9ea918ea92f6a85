"""Dev tool (GPU box): a CIGAR batch larger than the arena (waves sized by the expected traceback bytes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pairs = synth.make_batch(n, 100000, 0.05, 0)
o = mw.opt_init(flag=1)
with mw.Batch(o, pairs) as b:
    b.upload()
    t0 = time.perf_counter()
    b.run(); b.wait()
    dt = time.perf_counter() - t0
    r = b.fetch()
ok = all(mw.cigar2score(o, x[3]) == (x[0], len(t), len(q)) for x, (t, q) in zip(r[:8], pairs[:8]))
print("pairs", n, "run %.3f s" % dt, "launches", b.launches if False else "", "cells/s %.3e" % (sum(x[2] for x in r) / dt), "cigars ok:", ok, r[0][:3])
