"""Dev tool (GPU box): batch time under several values of one environment variable."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
mw.set_kernel(mw.KERNEL_TILE)
pairs = synth.make_batch(128, 100000, 0.05, 0)
var = sys.argv[1]
ref = None
for v in sys.argv[2:]:
    os.environ[var] = v
    with mw.Batch(mw.opt_init(), pairs) as b:
        b.upload(); b.run(); b.wait(); b.run(); b.wait()
        r = [(x[0], x[2]) for x in b.fetch()]
        ref = ref or r
        print(var, v, "%.2f ms" % b.kernel_ms, "same" if r == ref else "DIFFERENT", flush=True)
