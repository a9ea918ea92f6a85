"""Dev tool (GPU box): segmented traceback vs the all-at-once high-memory traceback on a large pair."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
n = int(sys.argv[1]); p = float(sys.argv[2]); segp = sys.argv[3] if len(sys.argv) > 3 else "4096"
t, q = synth.make_pair(n, p, 31337)
o = mw.opt_init(flag=1)
res = {}
for mode in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["full", "seg"]):
    if mode == "seg":
        os.environ["MWF_B200_TILE_SEGP"] = segp
    else:
        os.environ.pop("MWF_B200_TILE_SEGP", None)
    with mw.Batch(o, [(t, q)]) as b:
        b.upload()
        t0 = time.perf_counter()
        b.run(); b.wait()
        dt = time.perf_counter() - t0
        r = b.fetch()[0]
        res[mode] = r
        print("%s: s=%d n_cigar=%d n_iter=%d  run %.3f s  launches %d  cigar2score %s" % (mode, r[0], r[1], r[2], dt, b.launches, mw.cigar2score(o, r[3]) == (r[0], len(t), len(q))), flush=True)
if len(res) == 2:
    print("identical:", res["full"] == res["seg"])
