"""Dev tool (GPU box): A/B of the geometry switch threshold on the batch, the 150 kb pair and a 1 Mb pair."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
mw.set_kernel(mw.KERNEL_TILE)
work = [("batch128x100k", synth.make_batch(128, 100000, 0.05, 0), {}),
        ("single150k", [synth.make_pair(150000, 0.038, 900000)], {}),
        ("single150k-tb", [synth.make_pair(150000, 0.038, 900000)], {"flag": 1}),
        ("single1M", [synth.make_pair(1000000, 0.0097, 424242)], {}),
        ("8x300k", synth.make_batch(8, 300000, 0.03, 77), {})]
for sw in sys.argv[1:]:
    os.environ["MWF_B200_TILE_SWITCH"] = sw
    out = []
    for name, prs, kw in work:
        with mw.Batch(mw.opt_init(**kw), prs) as b:
            b.upload()
            b.run(); b.wait()
            b.run(); b.wait()
            out.append("%s %.2f" % (name, b.kernel_ms))
    print("switch=%s :: %s" % (sw, " | ".join(out)), flush=True)
