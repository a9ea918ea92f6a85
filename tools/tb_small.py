"""GPU box: a batch of many tiny pairs (what mwf_wfa_chain's gap fills look like): wall time of create / upload / run+wait /
fetch / destroy and the engine's kernel time, score-only, CIGAR and low-memory CIGAR.  Usage: python tools/tb_small.py [n_pairs]"""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miniwfa_b200 as mw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rng = random.Random(5)
pairs = []
for i in range(n):
    L = rng.randint(14, 60)
    t = bytearray(rng.choice(b"ACGT") for _ in range(L))
    q = bytearray(t)
    for _ in range(rng.randint(1, 3)):
        p = rng.randrange(len(q))
        r = rng.random()
        if r < 0.7:
            q[p] = rng.choice(b"ACGT")
        elif r < 0.85:
            del q[p]
        else:
            q.insert(p, rng.choice(b"ACGT"))
    pairs.append((bytes(t), bytes(q)))
arrays = mw.api.host_arrays(pairs)
for name, kw in (("score", {}), ("cigar", {"flag": 1}), ("cigar step=5000", {"flag": 1, "step": 5000})):
    for rep in range(2):
        o = mw.opt_init(**kw)
        tm = [time.perf_counter()]
        b = mw.Batch(o, pairs, arrays); tm.append(time.perf_counter())
        b.upload(); tm.append(time.perf_counter())
        b.run(); b.wait(); tm.append(time.perf_counter())
        r = (mw.MwfRst * b.n)()
        mw.lib().mwf_b200_batch_fetch(b.h, None, r); tm.append(time.perf_counter())
        kms, launches, used = b.kernel_ms, b.launches, b.kernel_used
        b.close(); tm.append(time.perf_counter())
        for i in range(b.n):
            if r[i].cigar:
                mw.lib().kfree(None, r[i].cigar)
        d = [(tm[i + 1] - tm[i]) * 1e3 for i in range(5)]
        print("%-16s rep %d: create %.1f upload %.1f run %.1f (kernel %.1f ms, %d launches, kernel %d) fetch %.1f destroy %.1f ms; sum s = %d"
              % (name, rep, d[0], d[1], d[2], kms, launches, used, d[3], d[4], sum(r[i].s for i in range(b.n))), flush=True)

# read-sized pairs: 150 bp, ~2 % differences
pairs = []
for i in range(n):
    t = bytes(rng.choice(b"ACGT") for _ in range(150))
    q = bytearray(t)
    for _ in range(3):
        q[rng.randrange(len(q))] = rng.choice(b"ACGT")
    pairs.append((t, bytes(q)))
arrays = mw.api.host_arrays(pairs)
for name, kw in (("150 bp score", {}), ("150 bp cigar", {"flag": 1})):
    o = mw.opt_init(**kw)
    for rep in range(2):
        t0 = time.perf_counter()
        with mw.Batch(o, pairs, arrays) as b:
            b.upload(); b.run(); b.wait()
            kms = b.kernel_ms
            r = (mw.MwfRst * b.n)()
            mw.lib().mwf_b200_batch_fetch(b.h, None, r)
        dt = time.perf_counter() - t0
        for i in range(n):
            if r[i].cigar:
                mw.lib().kfree(None, r[i].cigar)
        print("%-16s rep %d: %d pairs, kernel %.2f ms, end to end %.1f ms = %.2f M pairs/s" % (name, rep, n, kms, dt * 1e3, n / dt / 1e6), flush=True)
