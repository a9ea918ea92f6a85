"""Dev tool (GPU box): tile-engine parameter sweep on the config-3 batch and the single 150 kb pair."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth

mw.set_kernel(mw.KERNEL_TILE)
npairs = int(os.environ.get("NP", "128"))
pairs = synth.make_batch(npairs, 100000, 0.05, 0)
single = [synth.make_pair(150000, 0.038, 900000)]
ref = None
configs = [tuple(int(x) for x in c.split(",")) for c in sys.argv[1:]] or [(64, 256, 8, 4)]
for cfg in configs:
    T, NT, CPS = cfg[:3]
    CPT = cfg[3] if len(cfg) > 3 else 4
    os.environ["MWF_B200_TILE_CPT"] = str(CPT)
    os.environ["MWF_B200_TILE_T"] = str(T)
    os.environ["MWF_B200_TILE_THREADS"] = str(NT)
    os.environ["MWF_B200_TILE_CTAS_PER_SM"] = str(CPS)
    out = []
    for name, prs, kw in (("batch", pairs, {}), ("single", single, {}), ("single-tb", single, {"flag": 1})):
        with mw.Batch(mw.opt_init(**kw), prs) as b:
            if b.kernel_used != mw.KERNEL_TILE:
                out.append("%s: not eligible" % name)
                continue
            b.upload()
            b.run(); b.wait()
            b.run(); b.wait()
            r = b.fetch()
            ni = sum(x[2] for x in r)
            out.append("%s %.2f ms (%.3e c/s, %d launches)" % (name, b.kernel_ms, ni / b.kernel_ms * 1e3, b.launches))
            if name == "batch":
                key = [(x[0], x[2]) for x in r]
                if ref is None:
                    ref = key
                assert key == ref or os.environ.get("NOASSERT")
    print("T=%d NT=%d CPS=%d CPT=%d :: %s" % (T, NT, CPS, CPT, " | ".join(out)), flush=True)
