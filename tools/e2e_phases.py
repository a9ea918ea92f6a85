"""Dev tool (GPU box): wall time of each phase of the one-shot batch call on the config-3 batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
pairs = synth.make_batch(128, 100000, 0.05, 0)
opt = mw.opt_init()
arrs = mw.api.host_arrays(pairs)
for rep in range(4):
    t = [time.perf_counter()]
    b = mw.Batch(opt, pairs, arrays=arrs); t.append(time.perf_counter())
    b.upload(); t.append(time.perf_counter())
    b.run(); t.append(time.perf_counter())
    r = b.fetch(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create(+ctypes arrays)", "upload", "run", "fetch", "destroy"]
    print("rep", rep, " ".join("%s %.1f ms" % (n, (t[i + 1] - t[i]) * 1e3) for i, n in enumerate(names)), "total %.1f" % ((t[-1] - t[0]) * 1e3), flush=True)
