"""Dev tool (GPU box): randomized differential test of the whole engine against the unmodified reference (oracle/_ref).
Random sizes, divergences, penalties, modes (score / CIGAR / low-memory / stops), kernel families, tile geometries,
segmented traceback periods and wave sizes.  usage: fuzz.py <seed> <n_cases> [max_len]"""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from oracle import orc

seed, n_cases = int(sys.argv[1]), int(sys.argv[2])
max_len = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
rng = random.Random(seed)
ENV = ["MWF_B200_TILE_CPT", "MWF_B200_TILE_THREADS", "MWF_B200_TILE_T", "MWF_B200_TILE_SEGP", "MWF_B200_TILE_WAVE", "MWF_B200_TILE_ARENA_MAX", "MWF_B200_TILE_SWITCH", "MWF_B200_LOWMEM_STREAMING", "MWF_B200_TILE_PERSIST", "MWF_B200_TILE_FAST"]


def mutate(t, p):
    q = bytearray()
    for ch in t:
        u = rng.random()
        if u < p * 0.8:
            q.append(rng.choice(b"ACGT"))
        elif u < p * 0.9:
            q.extend(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 40))))
            q.append(ch)
        elif u < p:
            pass
        else:
            q.append(ch)
    return bytes(q)


def rand_pair():
    n = rng.choice([0, 1, 3, 30, 200, 1000, 3000, max_len // 2, max_len])
    alpha = rng.choice([b"ACGT", b"ACGT", b"ACGT", b"AC", b"ACGTN", bytes(range(256))])
    t = bytes(rng.choice(alpha) for _ in range(n))
    r = rng.random()
    if r < 0.8:
        q = mutate(t, rng.choice([0, 0.005, 0.02, 0.05, 0.15, 0.4]))
    elif r < 0.9:
        q = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 500)))
    else:  # a large block missing / inserted
        cut = rng.randint(0, max(0, n - 1))
        q = t[:cut] + t[min(n, cut + rng.randint(1, 800)):] if rng.random() < 0.5 else t[:cut] + bytes(rng.choice(alpha) for _ in range(rng.randint(1, 800))) + t[cut:]
    return t, q


def rand_opt():
    kw = {}
    pre = rng.choice("dddaerrr")
    if pre == "a":
        kw.update(o2=4, e2=2)
    elif pre == "e":
        kw.update(x=1, o1=0, o2=0, e1=1, e2=1)
    elif pre == "r":
        kw.update(x=rng.randint(1, 9), o1=rng.randint(0, 8), e1=rng.randint(1, 4), o2=rng.randint(0, 40), e2=rng.randint(1, 3))
    mode = rng.choice(["s", "c", "c", "c", "p", "p", "stop", "stopc"])
    if mode in ("c", "p", "stopc"):
        kw["flag"] = 1
    if mode == "p":
        kw["step"] = rng.choice([1, 2, 3, 5, 7, 16, 17, 37, 64, 100, 255, 256, 500, 5000])
    if mode in ("stop", "stopc"):
        kw[rng.choice(["max_s", "max_iter"])] = rng.randint(1, 30000)
    return kw


bad = 0
t_start = time.time()
for case in range(n_cases):
    for k in ENV:
        os.environ.pop(k, None)
    fam = rng.choice([mw.KERNEL_TILE] * 6 + [mw.KERNEL_AUTO, mw.KERNEL_CTA, mw.KERNEL_GRID])
    mw.set_kernel(fam)
    env = {}
    if rng.random() < 0.6:
        cpt, nt, T = rng.choice([(4, 128, 32), (4, 256, 64), (4, 64, 16), (2, 256, 32), (2, 128, 24), (1, 512, 64), (1, 256, 32), (1, 384, 48), (4, 96, 20)])
        env.update(MWF_B200_TILE_CPT=cpt, MWF_B200_TILE_THREADS=nt, MWF_B200_TILE_T=T)
    elif rng.random() < 0.5:
        env.update(MWF_B200_TILE_SWITCH=rng.choice([0, 1, 2, 5, 1000000]))
    if rng.random() < 0.35:
        env.update(MWF_B200_TILE_SEGP=rng.choice([256, 256, 512, 1024, 4096]))
    if rng.random() < 0.2:
        env.update(MWF_B200_TILE_ARENA_MAX=rng.choice([70000, 300000, 2000000]))
    if rng.random() < 0.3:
        env.update(MWF_B200_TILE_WAVE=rng.randint(1, 4))
    if rng.random() < 0.15:
        env.update(MWF_B200_LOWMEM_STREAMING=1)
    if rng.random() < 0.25:
        env.update(MWF_B200_TILE_PERSIST=0)  # one plan + one tile launch per block instead of the persistent kernel
    if rng.random() < 0.3:
        env.update(MWF_B200_TILE_FAST=rng.choice([0, 1]))  # the interior step with all rows in shared memory / the first register-resident step
    for k, v in env.items():
        os.environ[k] = str(v)
    kw = rand_opt()
    npairs = rng.choice([1, 1, 2, 5, 20])
    pairs = [rand_pair() for _ in range(npairs)]
    pairs = [(t, q) for t, q in pairs if not (len(t) + len(q) == 0 and kw.get("flag"))]
    if kw.get("step") == 1:
        pairs = [(t[:400], q[:400]) for t, q in pairs]
    elif 0 < kw.get("step", 0) < 64:  # the reference keeps a snapshot of the whole ring every `step` scores: host memory
        pairs = [(t[:6000], q[:6000]) for t, q in pairs]
    if not pairs:
        continue
    if os.environ.get("FUZZ_ONLY") and case != int(os.environ["FUZZ_ONLY"]):
        continue
    if case < int(os.environ.get("FUZZ_FROM", "0")):
        continue
    if os.environ.get("FUZZ_V"):
        print("case", case, "fam", fam, "env", env, "opt", kw, "lens", [(len(t), len(q)) for t, q in pairs][:6], flush=True)
    if any(len(set(t + q)) >= 255 for t, q in pairs):  # the reference needs two unused byte values (miniwfa.c:189-201)
        want = [orc.oracle_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
    else:
        want = [orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
    got = mw.wfa_exact_batch(mw.opt_init(**kw), pairs)
    if got != want:
        bad += 1
        for i, (g, w) in enumerate(zip(got, want)):
            if g != w:
                if os.environ.get("FUZZ_DUMP"):
                    open(os.environ["FUZZ_DUMP"], "wb").write(len(pairs[i][0]).to_bytes(4, "little") + pairs[i][0] + pairs[i][1])
                print("MISMATCH case", case, "pair", i, "lens", len(pairs[i][0]), len(pairs[i][1]), "fam", fam, "env", env, "opt", kw,
                      "want", w[:3], "got", g[:3], flush=True)
                break
    if case % 100 == 99:
        print("case", case + 1, "bad", bad, "%.0f s" % (time.time() - t_start), flush=True)
print("done", n_cases, "cases, bad =", bad)
sys.exit(1 if bad else 0)
