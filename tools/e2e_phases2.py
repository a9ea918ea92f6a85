"""Dev tool (GPU box): as e2e_phases.py but after the same prelude as bench.py (torch stream, a device-resident batch first)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import miniwfa_b200 as mw
from miniwfa_b200 import synth
pairs = synth.make_batch(128, 100000, 0.05, 0)
opt = mw.opt_init()
arrs = mw.api.host_arrays(pairs)
mode = sys.argv[1] if len(sys.argv) > 1 else "torchstream"
torch.cuda.set_device(0)
if mode != "none":
    stream = torch.cuda.Stream()
    b = mw.Batch(opt, pairs)
    if mode == "torchstream":
        b.set_stream(stream.cuda_stream)
    b.upload()
    for _ in range(2):
        b.run(); b.wait()
    b.fetch()
    b.close()
for rep in range(4):
    t = [time.perf_counter()]
    b = mw.Batch(opt, pairs, arrays=arrs); t.append(time.perf_counter())
    b.upload(); t.append(time.perf_counter())
    b.run(); t.append(time.perf_counter())
    r = b.fetch(); t.append(time.perf_counter())
    b.close(); t.append(time.perf_counter())
    names = ["create", "upload", "run", "fetch", "destroy"]
    print(mode, "rep", rep, " ".join("%s %.1f" % (n, (t[i + 1] - t[i]) * 1e3) for i, n in enumerate(names)), "total %.1f" % ((t[-1] - t[0]) * 1e3), flush=True)
