"""Dev tool (GPU box): does a fifth CTA per SM pay?  Penalties with a short ring (o2 = 9: 21 rows, 43 KB of shared memory per
tile) let five CTAs of the throughput geometry fit; MWF_B200_TILE_CTAS_PER_SM caps them at four for the comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
pairs = synth.make_batch(128, 100000, 0.05, 0)
for o2 in (9, 15):
    for cps in (4, 5, 8):
        os.environ["MWF_B200_TILE_CTAS_PER_SM"] = str(cps)
        with mw.Batch(mw.opt_init(o2=o2), pairs) as b:
            b.upload(); b.run(); b.wait(); b.run(); b.wait()
            r = b.fetch()
            ni = sum(x[2] for x in r)
            print("o2=%d ctas/SM<=%d: %.2f ms, %.3e cells/s" % (o2, cps, b.kernel_ms, ni / b.kernel_ms * 1e3), flush=True)
