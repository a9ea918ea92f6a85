"""Dev tool (GPU box): the 8 x 300 kb batch with the persistent kernel's debug line (geometry switches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
prs = synth.make_batch(8, 300000, 0.03, 77)
with mw.Batch(mw.opt_init(), prs) as b:
    b.upload(); b.run(); b.wait(); b.run(); b.wait()
    print("kernel_ms %.2f launches %d" % (b.kernel_ms, b.launches))
