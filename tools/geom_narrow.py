"""Dev tool (GPU box): a single pair whose band stays narrow (low divergence), default tile geometries against forced ones.
Usage: python tools/geom_narrow.py [n] [p]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
t, q = synth.make_pair(n, p, 31337)
ref = None
for name, env in (("default", {}),
                  ("1x512 T64", dict(MWF_B200_TILE_CPT=1, MWF_B200_TILE_THREADS=512, MWF_B200_TILE_T=64)),
                  ("1x256 T32", dict(MWF_B200_TILE_CPT=1, MWF_B200_TILE_THREADS=256, MWF_B200_TILE_T=32)),
                  ("1x384 T48", dict(MWF_B200_TILE_CPT=1, MWF_B200_TILE_THREADS=384, MWF_B200_TILE_T=48)),
                  ("1x320 T40", dict(MWF_B200_TILE_CPT=1, MWF_B200_TILE_THREADS=320, MWF_B200_TILE_T=40)),
                  ("2x256 T64", dict(MWF_B200_TILE_CPT=2, MWF_B200_TILE_THREADS=256, MWF_B200_TILE_T=64)),
                  ("1x128 T16", dict(MWF_B200_TILE_CPT=1, MWF_B200_TILE_THREADS=128, MWF_B200_TILE_T=16))):
    for k in ("MWF_B200_TILE_CPT", "MWF_B200_TILE_THREADS", "MWF_B200_TILE_T"):
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    for kw in ({}, {"flag": 1}):
        try:
            with mw.Batch(mw.opt_init(**kw), [(t, q)]) as b:
                b.upload()
                ms = []
                for _ in range(3):
                    b.run(); b.wait(); ms.append(b.kernel_ms)
                r = b.fetch()[0]
                if ref is None:
                    ref = r[:3]
                print("%-10s %-10s s=%d kernel %.2f ms, %d launches, kernel family %d%s" % (name, "cigar" if kw else "score", r[0], min(ms), b.launches, b.kernel_used,
                      "" if r[0] == ref[0] else "  MISMATCH"), flush=True)
        except Exception as e:
            print(name, "failed", e)
