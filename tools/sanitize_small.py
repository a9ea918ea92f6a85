"""Dev tool (GPU box): a few small cases through every tile-engine path, to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py
    compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
from oracle import orc
mw.set_kernel(mw.KERNEL_TILE)
batch = synth.make_batch(16, 2500, 0.06, 900)                       # throughput geometry, 4 cells per thread
single = [synth.make_pair(6000, 0.05, 950)]                          # latency geometry, 2 cells per thread
nbatch = [(t.replace(b"A", b"N", 3), q) for t, q in batch[:16]]     # four-bit codes
bad = 0
for name, pairs in (("batch", batch), ("single", single), ("batch-N", nbatch)):
    for kw in ({}, {"flag": 1}, {"flag": 1, "step": 300}):
        want = [orc.oracle_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
        got = mw.wfa_exact_batch(mw.opt_init(**kw), pairs)
        ok = got == want
        bad += not ok
        print(name, kw, "ok" if ok else "MISMATCH", flush=True)
os.environ["MWF_B200_TILE_SEGP"] = "256"
want = [orc.oracle_exact(orc.make_opt(flag=1), t, q) for t, q in single]
ok = mw.wfa_exact_batch(mw.opt_init(flag=1), single) == want
bad += not ok
print("segmented", "ok" if ok else "MISMATCH")
sys.exit(1 if bad else 0)
