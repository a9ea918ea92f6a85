"""Dev tool (GPU box): forced tile geometries on single / few large pairs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
mw.set_kernel(mw.KERNEL_TILE)
work = [("single150k", [synth.make_pair(150000, 0.038, 900000)], {}),
        ("single150k-tb", [synth.make_pair(150000, 0.038, 900000)], {"flag": 1}),
        ("single1M", [synth.make_pair(1000000, 0.0097, 424242)], {}),
        ("8x300k", synth.make_batch(8, 300000, 0.03, 77), {})]
for cfg in sys.argv[1:]:
    for k in ("MWF_B200_TILE_T", "MWF_B200_TILE_THREADS", "MWF_B200_TILE_CPT"):
        os.environ.pop(k, None)
    if cfg != "auto":
        T, NT, CPT = cfg.split(",")
        os.environ.update(MWF_B200_TILE_T=T, MWF_B200_TILE_THREADS=NT, MWF_B200_TILE_CPT=CPT)
    out = []
    for name, prs, kw in work:
        with mw.Batch(mw.opt_init(**kw), prs) as b:
            if b.kernel_used != mw.KERNEL_TILE:
                out.append("%s n/a" % name); continue
            b.upload(); b.run(); b.wait(); b.run(); b.wait()
            out.append("%s %.2f" % (name, b.kernel_ms))
    print(cfg, "::", " | ".join(out), flush=True)
