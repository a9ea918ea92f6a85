"""Dev tool (GPU box): batch time with 4-, 5- and 256-symbol alphabets (two-bit, four-bit, raw probes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
mw.set_kernel(mw.KERNEL_TILE)
base = synth.make_batch(128, 100000, 0.05, 0)
def with_n(s):
    b = bytearray(s)
    b[::997] = b"N" * len(b[::997])
    return bytes(b)
def spread(s):  # more than 16 distinct bytes: every 2000th base gets one of 32 other values
    b = bytearray(s)
    for k, i in enumerate(range(0, len(b), 2000)):
        b[i] = 128 + (k % 32)
    return bytes(b)
for name, f in (("ACGT (2-bit)", lambda x: x), ("ACGT+N (4-bit)", with_n), ("40 symbols (raw)", spread)):
    pairs = [(f(t), f(q)) for t, q in base]
    with mw.Batch(mw.opt_init(), pairs) as b:
        b.upload(); b.run(); b.wait(); b.run(); b.wait()
        r = b.fetch()
        print(name, "%.2f ms" % b.kernel_ms, r[0][:3], flush=True)
