"""Dev tool (GPU box): randomized differential test of mwf_wfa_chain / mwf_wfa_auto (device k-mer front end, host LIS, batched
gap fills) against the unmodified reference (oracle/_ref): repeats, low-complexity stretches, N runs, soft-masking, large
insertions of unrelated sequence, random k / max_occ / min_len / penalties.  usage: fuzz_chain.py <seed> <n_cases>"""
import ctypes, os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from oracle import orc

seed, n_cases = int(sys.argv[1]), int(sys.argv[2])
rng = random.Random(seed)
ref = orc.reference()
assert ref is not None, "oracle/_ref missing"


def rand_seq(n):
    out = bytearray()
    while len(out) < n:
        r = rng.random()
        m = rng.randint(1, max(1, n // 5))
        if r < 0.6:
            out += bytes(rng.choice(b"ACGT") for _ in range(m))
        elif r < 0.7:
            out += bytes([rng.choice(b"ACGT")]) * min(m, 400)              # homopolymer
        elif r < 0.8:
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(2, 40)))
            out += unit * rng.randint(2, 30)                                # tandem repeat
        elif r < 0.9 and len(out) > 50:
            a = rng.randrange(len(out) - 20)
            out += out[a:a + rng.randint(20, 3000)]                         # dispersed copy of an earlier stretch
        elif r < 0.95:
            out += b"N" * rng.randint(1, 200)
        else:
            out += bytes(rng.choice(b"acgt") for _ in range(min(m, 500)))  # soft-masked
    return bytes(out[:n])


def mutate(t, p):
    q = bytearray()
    for ch in t:
        u = rng.random()
        if u < p * 0.7:
            q.append(rng.choice(b"ACGT"))
        elif u < p * 0.85:
            q.extend(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 8))))
            q.append(ch)
        elif u < p:
            pass
        else:
            q.append(ch)
    return q


bad = 0
t0 = time.time()
for case in range(n_cases):
    n = rng.choice([50, 500, 3000, 5000, 20000, 60000, 150000])
    t = rand_seq(n)
    q = mutate(t, rng.choice([0.0, 0.003, 0.02, 0.06, 0.15]))
    if rng.random() < 0.3:   # a long unrelated insert / a long deletion (mwf_ksim's branch needs >= 10 kb on both sides)
        cut = rng.randrange(len(q) + 1)
        q[cut:cut] = rand_seq(rng.choice([300, 12000, 25000]))
        if rng.random() < 0.5 and len(t) > 30000:
            a = rng.randrange(len(t) - 15000)
            t = t[:a] + rand_seq(rng.choice([11000, 20000])) + t[a:]
    q = bytes(q)
    kw = {"flag": rng.choice([0, 1, 1])}
    if kw["flag"]:
        kw["step"] = rng.choice([0, 5000, 5000, 100])
    if rng.random() < 0.5:
        kw.update(kmer=rng.choice([5, 8, 11, 13, 15]), max_occ=rng.choice([1, 2, 3, 10]), min_len=rng.choice([0, 20, 30, 100]))
    if rng.random() < 0.2:
        kw.update(x=rng.randint(1, 6), o1=rng.randint(0, 6), e1=rng.randint(1, 3), o2=rng.randint(0, 30), e2=rng.randint(1, 2))
    fn = rng.choice(["mwf_wfa_chain", "mwf_wfa_chain", "mwf_wfa_auto"])
    os.environ["MWF_B200_CHAIN_FRONT"] = rng.choice(["gpu", "gpu", "host"])
    o = mw.opt_init(**kw)
    r = mw.MwfRst()
    getattr(mw.lib(), fn)(None, ctypes.byref(o), len(t), t, len(q), q, ctypes.byref(r))
    got = (r.s, r.n_cigar, r.cigar[:r.n_cigar] if r.n_cigar > 0 else [])
    if r.cigar:
        mw.lib().kfree(None, r.cigar)
    ro, rr = orc.make_opt(**kw), orc.Rst()
    getattr(ref, fn)(None, ctypes.byref(ro), len(t), t, len(q), q, ctypes.byref(rr))
    want = (rr.s, rr.n_cigar, rr.cigar[:rr.n_cigar] if rr.n_cigar > 0 else [])
    if got != want:
        bad += 1
        print("MISMATCH case", case, fn, "lens", len(t), len(q), "opt", kw, "front", os.environ["MWF_B200_CHAIN_FRONT"], "want", want[:2], "got", got[:2], flush=True)
    if case % 50 == 49:
        print("case", case + 1, "bad", bad, "%.0f s" % (time.time() - t0), flush=True)
print("done", n_cases, "cases, bad =", bad)
sys.exit(1 if bad else 0)
