"""Dev tool (GPU box): tile engine vs the CPU checker on a ladder of sizes, then timings."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
from oracle import orc

mw.set_kernel(mw.KERNEL_TILE)
bad = 0
def check(kw, t, q, tag):
    global bad
    if os.environ.get("V"): print("case", tag, kw, len(t), len(q), flush=True)
    want = orc.checker_exact(orc.make_opt(**kw), t, q)
    got = mw.wfa_exact(mw.opt_init(**kw), t, q)
    ok = got == want
    if not ok:
        bad += 1
        print("MISMATCH", tag, kw, "want", want[:3], "got", got[:3], flush=True)
    return ok

t3 = (b"CAGGGGCAGACTGACACTTCACACGGCCGGGTACTCTAACAGACCTGCAGCTGAGGGTCCT",
      b"TAGGGGCAGACTGACACCTCACACGGCCGGGTACTCCTCTGAGACAAAACTTCCAGAGGAACGATCAGACAGCAGCATTCGCGGTTCATGAAAATCCGCTGTTCTG"
      b"CAGCCACCGCTGCTGGTACCCAGGCAAACAGGGTCTAGAGTGGACCTTTAGCAAACTCCAACAGACCTGCAGCTGAGGGTCCT")
for kw in ({}, {"flag": 1}, {"flag": 1, "x": 1, "o1": 0, "o2": 0, "e1": 1, "e2": 1}, {"flag": 1, "o2": 4, "e2": 2}):
    check(kw, t3[0], t3[1], "t3")
    check(kw, t3[1], t3[0], "t3swap")
for t, q in ((b"ACGT", b"ACGT"), (b"ACGT", b"ACCT"), (b"ACGT", b""), (b"", b"ACGT"), (b"A", b"C"), (b"A" * 30, b"A" * 10), (b"AAAA", b"CCCC")):
    for kw in ({}, {"flag": 1}, {"max_s": 3}, {"max_iter": 10}):
        check(kw, t, q, "kat")
print("small done, bad =", bad, flush=True)
for n, p in ((300, 0.1), (1000, 0.05), (3000, 0.15), (3000, 0.4), (10000, 0.05), (20000, 0.02), (30000, 0.1)):
    for i in range(2):
        t, q = synth.make_pair(n, p, 50 + i)
        for kw in ({}, {"flag": 1}, {"max_iter": 200000}, {"flag": 1, "x": 2, "o1": 3, "e1": 1, "o2": 9, "e2": 1}):
            t0 = time.time()
            ok = check(kw, t, q, "n%d p%g" % (n, p))
    print("size", n, p, "bad =", bad, flush=True)
if bad:
    sys.exit(1)
# timings
for npairs, n, p in ((128, 100000, 0.05),):
    pairs = synth.make_batch(npairs, n, p, 0)
    for fam in (mw.KERNEL_TILE, mw.KERNEL_CTA):
        mw.set_kernel(fam)
        with mw.Batch(mw.opt_init(), pairs) as b:
            b.upload()
            for _ in range(2):
                b.run(); b.wait()
            r = b.fetch()
            ni = sum(x[2] for x in r)
            print("batch", npairs, n, "family", fam, "kernel_ms %.2f" % b.kernel_ms, "launches", b.launches,
                  "wavefront cells/s %.3e" % (ni / b.kernel_ms * 1e3), "s0", r[0][:3], flush=True)
t, q = synth.make_pair(150000, 0.038, 900000)
for fam in (mw.KERNEL_TILE, mw.KERNEL_GRID):
    mw.set_kernel(fam)
    for kw in ({}, {"flag": 1}):
        with mw.Batch(mw.opt_init(**kw), [(t, q)]) as b:
            b.upload()
            for _ in range(2):
                b.run(); b.wait()
            r = b.fetch()[0]
            print("single 150k", kw, "family", fam, "kernel_ms %.2f" % b.kernel_ms, r[:3], flush=True)
