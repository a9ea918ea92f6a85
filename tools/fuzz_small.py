"""Dev tool (GPU box): large batches of short pairs (what mwf_wfa_chain's gap fills and read-sized inputs look like) through
the one-CTA-per-pair streaming kernel with its small-CTA geometries (64 / 128 / 256 threads, several CTAs per SM), against the
unmodified reference (oracle/_ref) or the oracle.  usage: fuzz_small.py <seed> <n_batches>"""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from oracle import orc

seed, n_batches = int(sys.argv[1]), int(sys.argv[2])
rng = random.Random(seed)


def mutate(t, p):
    q = bytearray()
    for ch in t:
        u = rng.random()
        if u < p * 0.7:
            q.append(rng.choice(b"ACGT"))
        elif u < p * 0.85:
            q.extend(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 6))))
            q.append(ch)
        elif u < p:
            pass
        else:
            q.append(ch)
    return bytes(q)


bad = 0
t0 = time.time()
for it in range(n_batches):
    max_n = rng.choice([40, 150, 250, 900, 1900])
    n_pairs = rng.choice([300, 700, 1500, 4000]) if max_n <= 250 else rng.choice([300, 700])
    pairs = []
    for _ in range(n_pairs):
        n = rng.randint(0, max_n)
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        q = mutate(t, rng.choice([0, 0.01, 0.05, 0.2])) if rng.random() < 0.9 else bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, max_n)))
        pairs.append((t, q))
    kw = {}
    mode = rng.choice(["s", "c", "c", "p", "stop"])
    if mode in ("c", "p"):
        kw["flag"] = 1
    if mode == "p":
        kw["step"] = rng.choice([3, 17, 100, 5000])
    if mode == "stop":
        kw[rng.choice(["max_s", "max_iter"])] = rng.randint(1, 3000)
    if rng.random() < 0.3:
        kw.update(x=rng.randint(1, 9), o1=rng.randint(0, 8), e1=rng.randint(1, 4), o2=rng.randint(0, 40), e2=rng.randint(1, 3))
    if kw.get("flag"):
        pairs = [(t, q) for t, q in pairs if len(t) + len(q) > 0]
    mw.set_kernel(rng.choice([mw.KERNEL_AUTO, mw.KERNEL_CTA]))
    want = [orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
    got = mw.wfa_exact_batch(mw.opt_init(**kw), pairs)
    if got != want:
        bad += 1
        i = next(i for i in range(len(pairs)) if got[i] != want[i])
        print("MISMATCH batch", it, "pair", i, "lens", len(pairs[i][0]), len(pairs[i][1]), "opt", kw, "want", want[i][:3], "got", got[i][:3], flush=True)
print("done", n_batches, "batches, bad =", bad, "%.0f s" % (time.time() - t0))
sys.exit(1 if bad else 0)
