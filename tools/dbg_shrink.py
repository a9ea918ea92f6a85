import os, sys
sys.path.insert(0, "/root/repo")
os.environ["MWF_B200_DEBUG"] = "1"
import miniwfa_b200 as mw
from miniwfa_b200 import synth
mw.set_kernel(mw.KERNEL_TILE)
for n, p in ((150000, 0.038), (20000, 0.05)):
    prs = [synth.make_pair(n, p, 900000)]
    with mw.Batch(mw.opt_init(), prs) as b:
        b.upload(); b.run(); b.wait()
        print(n, b.kernel_ms, flush=True)
