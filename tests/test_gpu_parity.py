"""GPU: the CUDA path, called through the C-ABI (mwf_wfa_exact / mwf_wfa_exact_batch / mwf_b200_batch_*),
against the reference's golden vectors, the oracle (and oracle/_ref when it travelled with the snapshot) on
seeded random inputs, and size-independent properties at full size.  Bit-exact: s, n_iter, n_cigar, CIGAR words."""
import random

import pytest

from conftest import case_inputs, chain_case_inputs
import miniwfa_b200 as mw
from miniwfa_b200 import synth
from miniwfa_b200.api import cigar_string
from oracle import orc

pytestmark = pytest.mark.gpu

KERNELS = [("tile", mw.KERNEL_TILE), ("cta", mw.KERNEL_CTA), ("grid", mw.KERNEL_GRID)]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(product_lib):
    assert mw.device_count() > 0, "no CUDA device visible: the product has no CPU fallback"
    yield
    mw.set_kernel(mw.KERNEL_AUTO)


def exact_cases(golden, big=False):
    out = []
    for c in golden:
        if c.get("fn", "mwf_wfa_exact") != "mwf_wfa_exact":
            continue
        if c["name"].startswith(("n100k", "n20k")) != big:
            continue
        out.append(c)
    return out


def expect(c):
    e = c["expect"]
    return (e["s"], e["n_cigar"], e["n_iter"], e["cigar"])


def got(r):
    return (r[0], r[1], r[2], cigar_string(r[3]))


@pytest.mark.parametrize("kname,kernel", KERNELS)
def test_golden_single_pair_api(golden, kname, kernel):
    """mwf_wfa_exact(), one call per golden case, with each kernel family."""
    mw.set_kernel(kernel)
    n = 0
    for c in exact_cases(golden):
        t, q = case_inputs(c)
        if len(t) + len(q) == 0 and c["opt"].get("flag", 0) & 1:
            continue  # the reference itself is undefined here (reads tb.a[-1])
        r = mw.wfa_exact(mw.opt_init(**c["opt"]), t, q)
        assert got(r) == expect(c), (kname, c["name"])
        n += 1
    assert n > 250


def test_golden_batched(golden):
    """The same golden cases submitted as batches (one batch per option set) through mwf_wfa_exact_batch."""
    mw.set_kernel(mw.KERNEL_AUTO)
    groups = {}
    for c in exact_cases(golden):
        t, q = case_inputs(c)
        if len(t) + len(q) == 0 and c["opt"].get("flag", 0) & 1:
            continue
        groups.setdefault(tuple(sorted(c["opt"].items())), []).append((c, t, q))
    assert len(groups) > 20
    for key, items in groups.items():
        rs = mw.wfa_exact_batch(mw.opt_init(**dict(key)), [(t, q) for _, t, q in items])
        for (c, _, _), r in zip(items, rs):
            assert got(r) == expect(c), c["name"]


def test_empty_pair_with_cigar_is_defined():
    r = mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR), b"", b"")
    assert r == (0, 0, 0, [])
    assert mw.wfa_exact_batch(mw.opt_init(), []) == []


@pytest.mark.parametrize("kname,kernel", KERNELS)
def test_golden_large(golden, kname, kernel):
    mw.set_kernel(kernel)
    for c in exact_cases(golden, big=True):
        if kname == "cta" and c["name"].startswith("n100k"):
            continue
        t, q = case_inputs(c)
        r = mw.wfa_exact(mw.opt_init(**c["opt"]), t, q)
        assert got(r) == expect(c), (kname, c["name"])


def mutate(rng, t, p):
    q = bytearray()
    for ch in t:
        u = rng.random()
        if u < p * 0.8:
            q.append(rng.choice(b"ACGT"))
        elif u < p * 0.9:
            q.extend(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 30))))
            q.append(ch)
        elif u < p:
            pass
        else:
            q.append(ch)
    return bytes(q)


def random_opt(rng):
    kw = {}
    pre = rng.choice(["d", "d", "a", "e", "r", "r"])
    if pre == "a":
        kw.update(o2=4, e2=2)
    elif pre == "e":
        kw.update(x=1, o1=0, o2=0, e1=1, e2=1)
    elif pre == "r":
        kw.update(x=rng.randint(1, 9), o1=rng.randint(0, 8), e1=rng.randint(1, 4), o2=rng.randint(0, 30), e2=rng.randint(1, 3))
    mode = rng.choice(["s", "c", "c", "p", "p", "stop"])
    if mode != "s":
        kw["flag"] = mw.F_CIGAR
    if mode == "p":
        kw["step"] = rng.choice([1, 2, 5, 7, 37, 500, 5000])
    if mode == "stop":
        kw[rng.choice(["max_s", "max_iter"])] = rng.randint(1, 20000)
    return kw


@pytest.mark.parametrize("kname,kernel", KERNELS)
def test_random_differential(kname, kernel):
    """Seeded random pairs and random valid penalty sets against the CPU checker."""
    mw.set_kernel(kernel)
    rng = random.Random({"cta": 42, "grid": 43}.get(kname, 44))
    for it in range(160):
        n = rng.choice([0, 1, 2, 5, 17, 60, 200, 300, 1000, 3000]) if it % 20 else 9000
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        if rng.random() < 0.85:
            q = mutate(rng, t, rng.choice([0, 0.01, 0.05, 0.15, 0.4]))
        else:
            q = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(0, 400)))
        kw = random_opt(rng)
        if kw.get("step") == 1 and len(t) > 400:
            kw["step"] = 3
        if len(t) + len(q) == 0 and kw.get("flag"):
            continue
        want = orc.checker_exact(orc.make_opt(**kw), t, q)
        assert mw.wfa_exact(mw.opt_init(**kw), t, q) == want, (kname, it, n, kw)


def test_random_batch_ragged():
    """One ragged batch (lengths 0..5000, including empty and unrelated pairs) per mode."""
    mw.set_kernel(mw.KERNEL_AUTO)
    rng = random.Random(99)
    pairs = []
    for i in range(200):
        n = rng.choice([0, 1, 10, 100, 500, 2000, 5000])
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        q = mutate(rng, t, rng.choice([0, 0.02, 0.1, 0.3])) if i % 7 else bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 300)))
        if len(t) + len(q):
            pairs.append((t, q))
    for kw in ({}, {"flag": mw.F_CIGAR}, {"flag": mw.F_CIGAR, "step": 50}, {"flag": mw.F_CIGAR, "max_s": 300},
               {"x": 1, "o1": 0, "o2": 0, "e1": 1, "e2": 1, "flag": mw.F_CIGAR}):
        o = orc.make_opt(**kw)
        want = [orc.checker_exact(o, t, q) for t, q in pairs]
        assert mw.wfa_exact_batch(mw.opt_init(**kw), pairs) == want, kw


def test_bytes_and_case_sensitivity():
    mw.set_kernel(mw.KERNEL_AUTO)
    o = mw.opt_init(flag=mw.F_CIGAR)
    assert got(mw.wfa_exact(o, b"ACGTNNNNACGT", b"acgtNNNNacgt"))[:2] == (32, 3)
    # all 256 byte values in use: the reference aborts (no free sentinel); the engine does not need sentinels
    t = bytes(range(256)) * 2
    q = t[:100] + b"\x00" + t[100:300] + t[305:]
    r = mw.wfa_exact(o, t, q)
    want = orc.oracle_exact(orc.make_opt(flag=1), t, q)
    assert r == want and mw.cigar2score(o, r[3]) == (r[0], len(t), len(q))


def test_full_size_properties():
    """100 kb / 5 % (the BASELINE config-3 pair shape): score vs golden, CIGAR re-scoring, low-memory == high-memory."""
    mw.set_kernel(mw.KERNEL_AUTO)
    t, q = synth.make_pair(100000, 0.05, 0)
    s0 = mw.wfa_exact(mw.opt_init(), t, q)
    assert s0[:3] == (23932, 0, 572070878)
    o = mw.opt_init(flag=mw.F_CIGAR)
    hi = mw.wfa_exact(o, t, q)
    assert hi[0] == s0[0] and hi[2] == s0[2]
    assert mw.cigar2score(o, hi[3]) == (hi[0], len(t), len(q))
    lo = mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR, step=5000), t, q)
    assert lo[0] == hi[0] and lo[3] == hi[3] and lo[2] < hi[2]
    # symmetry: swapping the sequences swaps I and D
    sw = mw.wfa_exact(o, q, t)
    flip = {1: 2, 2: 1}
    assert sw[0] == hi[0] and mw.cigar2score(o, sw[3]) == (hi[0], len(q), len(t))
    # (co-optimal alignments may differ between the two orientations -- the tie-breaks are not symmetric -- so only the
    # operation counts of the swapped CIGAR are compared: it must consume the swapped lengths, checked above, at the same score)
    n_ins = sum(c >> 4 for c in hi[3] if c & 0xf == 1), sum(c >> 4 for c in sw[3] if c & 0xf == flip[1])
    assert abs(n_ins[0] - n_ins[1]) <= hi[0]


def _sha1_words(words):
    import hashlib
    import struct
    return hashlib.sha1(struct.pack("<%dI" % len(words), *words)).hexdigest()


def _large_case(golden_large, name):
    for c in golden_large:
        if c["name"] == name:
            return c
    pytest.skip("tests/golden/golden_large.json has no case %s yet" % name)


def _check_large(c, r):
    e = c["expect"]
    assert (r[0], r[1], r[2]) == (e["s"], e["n_cigar"], e["n_iter"]), c["name"]
    assert _sha1_words(r[3]) == e["cigar_sha1"], c["name"]


def test_config2_full_size_against_reference(golden_large, monkeypatch):
    """BASELINE config 2 at its size (150 kb pair, s = 27 362): s, n_iter, n_cigar and every CIGAR word (sha1 over the words)
    equal to the unmodified reference's (tests/golden/make_golden_large.py), high-memory and -cp5000, and through the
    segmented traceback."""
    mw.set_kernel(mw.KERNEL_AUTO)
    for name in ("config2-c", "config2-cp5000"):
        c = _large_case(golden_large, name)
        t, q = synth.make_pair(*c["synth"])
        assert (len(t), len(q)) == (c["tl"], c["ql"])
        _check_large(c, mw.wfa_exact(mw.opt_init(**c["opt"]), t, q))
    c = _large_case(golden_large, "config2-c")
    t, q = synth.make_pair(*c["synth"])
    monkeypatch.setenv("MWF_B200_TILE_SEGP", "4096")
    _check_large(c, mw.wfa_exact(mw.opt_init(**c["opt"]), t, q))


def test_config3_shard_against_reference(golden_large):
    """The first 128 pairs of the BASELINE config-3 batch (rank 0's shard at 8 GPUs): (s, n_iter) of every pair equal to the
    reference's, and the sha1 of the committed list is the one the generator recorded."""
    import hashlib
    c = _large_case(golden_large, "config3")
    rows = c["expect"]["s_n_iter"]
    assert hashlib.sha1(("".join("%d,%d;" % (s, ni) for s, ni in rows)).encode()).hexdigest() == c["expect"]["sha1"]
    mw.set_kernel(mw.KERNEL_AUTO)
    pairs = synth.make_batch(128, 100000, 0.05, 0)
    rs = mw.wfa_exact_batch(mw.opt_init(), pairs)
    assert [[r[0], r[2]] for r in rs] == rows[:128]


def test_one_megabase_pair_against_reference(golden_large, monkeypatch):
    """1 Mb / 3 % (s = 142 199, 2e10 traceback bytes): high-memory CIGAR all at once and through the segmented traceback
    (arena capped at 2 GB), and low-memory mode -cp5000, against the reference's result."""
    mw.set_kernel(mw.KERNEL_AUTO)
    hi, lo = _large_case(golden_large, "n1m-p3-c"), _large_case(golden_large, "n1m-p3-cp5000")
    t, q = synth.make_pair(*hi["synth"])
    _check_large(hi, mw.wfa_exact(mw.opt_init(**hi["opt"]), t, q))
    _check_large(lo, mw.wfa_exact(mw.opt_init(**lo["opt"]), t, q))
    monkeypatch.setenv("MWF_B200_TILE_ARENA_MAX", str(2 << 30))
    _check_large(hi, mw.wfa_exact(mw.opt_init(**hi["opt"]), t, q))
    _check_large(lo, mw.wfa_exact(mw.opt_init(**lo["opt"]), t, q))
    mw.release_cache()


def test_config4_full_size_against_reference(golden_large):
    """BASELINE config 4 surrogate at its size (5 Mb pair, s = 231 245, -cp5000) against the reference's result (11 minutes of
    one host core in the build container)."""
    mw.set_kernel(mw.KERNEL_AUTO)
    c = _large_case(golden_large, "config4-cp5000")
    t, q = synth.make_pair(*c["synth"])
    _check_large(c, mw.wfa_exact(mw.opt_init(**c["opt"]), t, q))
    mw.release_cache()


def test_config5_full_size_against_reference(golden_large):
    """BASELINE config 5 surrogate at its size (5 Mb pair at 3 %, s = 712 856): the high-memory CIGAR (segmented traceback: its
    5e11 traceback bytes fit no memory) must be the CIGAR the reference gives with -cp5000 --
    s, n_cigar and the sha1 of the words; n_iter differs by construction (the reference's is pass 2's)."""
    mw.set_kernel(mw.KERNEL_AUTO)
    c = _large_case(golden_large, "config5-cp5000")
    t, q = synth.make_pair(*c["synth"])
    r = mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR), t, q)
    e = c["expect"]
    assert (r[0], r[1]) == (e["s"], e["n_cigar"]) and _sha1_words(r[3]) == e["cigar_sha1"]
    mw.release_cache()


def test_batch_object_reuse_and_timers():
    pairs = synth.make_batch(8, 20000, 0.05, 100)
    o = mw.opt_init()
    want = [orc.checker_exact(orc.make_opt(), t, q) for t, q in pairs]
    with mw.Batch(o, pairs) as b:
        for _ in range(2):
            b.upload()
            b.run()
            b.wait()
            assert b.fetch() == want
        assert b.kernel_ms > 0 and b.launches >= 1 and b.h2d_bytes > 8 * 40000


@pytest.mark.parametrize("front", ["default", "gpu", "host"])
def test_chain_and_auto_golden(golden, golden_chain, monkeypatch, front):
    """mwf_wfa_chain / mwf_wfa_auto (k-mer front end on the device or, for small inputs, the host; LIS on the host; one GPU
    batch of gap fills) against the reference's own output: score, CIGAR words and the n_iter the reference leaves in the result."""
    mw.set_kernel(mw.KERNEL_AUTO)
    if front != "default":
        monkeypatch.setenv("MWF_B200_CHAIN_FRONT", front)
    import ctypes
    L = mw.lib()
    n = 0
    for c in list(golden_chain) + [c for c in golden if c.get("fn", "mwf_wfa_exact") != "mwf_wfa_exact"]:
        t, q = chain_case_inputs(c) if "recipe" in c else case_inputs(c)
        o = mw.opt_init(**c["opt"])
        r = mw.MwfRst()
        getattr(L, c["fn"])(None, ctypes.byref(o), len(t), t, len(q), q, ctypes.byref(r))
        cig = [r.cigar[i] for i in range(r.n_cigar)]
        if r.cigar:
            L.kfree(None, r.cigar)
        e = c["expect"]
        assert (r.s, r.n_cigar, r.n_iter, cigar_string(cig)) == (e["s"], e["n_cigar"], e["n_iter"], e["cigar"]), c["name"]
        if c["opt"].get("flag", 0) & 1:
            assert mw.cigar2score(o, cig)[1:] == (len(t), len(q)), c["name"]
        n += 1
    assert n >= 40


_NT4 = {ord(c): v for c, v in zip("ACGTUacgtu", (0, 1, 2, 3, 3, 0, 1, 2, 3, 3))}
_NT4.update({0: 0, 1: 1, 2: 2, 3: 3})


def _py_kmers(s, k):
    """(k-mer word, position of its last base) of every valid k-mer: mg_fc_kmer (miniwfa.c:718-730), plain Python."""
    out, word, run, mask = [], 0, 0, (1 << 2 * k) - 1
    for i, ch in enumerate(s):
        c = _NT4.get(ch, 4)
        if c < 4:
            word, run = (word << 2 | c) & mask, run + 1
            if run >= k:
                out.append((word, i))
        else:
            word, run = 0, 0
    return out


def _py_hits_and_shared(t, q, k, max_occ):
    from collections import defaultdict
    dt, dq = defaultdict(list), defaultdict(list)
    kt, kq = _py_kmers(t, k), _py_kmers(q, k)
    for w, i in kt:
        dt[w].append(i)
    for w, i in kq:
        dq[w].append(i)
    hits = sorted((tp, qp) for w, tps in dt.items() if w in dq and len(tps) <= max_occ and len(dq[w]) <= max_occ
                  for tp in tps for qp in dq[w])
    shared = sum(min(len(v), len(dq[w])) for w, v in dt.items() if w in dq)
    return [qp << 32 | tp for tp, qp in hits], (len(kt), len(kq), shared)


def test_kmer_front_end_against_python_restatement():
    """mwf_b200_kmer_hits / mwf_b200_kmer_shared (list, sort, match, sort on the device) against a plain-Python restatement
    of mg_fc_kmer + the match loop of mg_chain (miniwfa.c:718-770) and of mwf_ksim's counts (:786-812): soft-masked and
    N-containing sequences, raw 0..3 codes, homopolymer runs (large groups), every k from 2 to 15, several max_occ."""
    rng = random.Random(99)

    def mutate(s, p):
        out = bytearray()
        for ch in s:
            r = rng.random()
            if r < p * 0.4:
                continue
            if r < p * 0.8:
                out.append(rng.choice(b"ACGT"))
            out.append(ch)
        return bytes(out)

    cases = []
    for k in range(2, 16):
        n = 200 if k < 5 else 6000
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        cases.append((t, mutate(t, 0.05), k, rng.choice((1, 2, 3, 50))))
    t = bytes(rng.choice(b"ACGTacgtNnU") for _ in range(20000))
    cases.append((t, mutate(t, 0.03), 13, 2))
    cases.append((t, mutate(t, 0.03), 7, 4))
    t = bytes(rng.choice((0, 1, 2, 3)) for _ in range(5000))  # raw codes
    cases.append((t, mutate(t, 0.02), 11, 2))
    t = b"A" * 3000 + bytes(rng.choice(b"ACGT") for _ in range(3000)) + b"CA" * 1500 + b"T" * 5  # huge groups
    cases.append((t, mutate(t, 0.02), 13, 2))
    cases.append((t, mutate(t, 0.02), 4, 3000))
    cases.append((b"ACGTACGTACGTAC", b"ACGTACGTACGTAC", 13, 2))   # two k-mers each
    cases.append((b"ACGTACGTACGTA", b"NNNNNNNNNNNNNNNN", 13, 2))    # nothing valid in the query
    cases.append((b"ACGTACGTACGT", b"ACGTACGTACGTACGT", 13, 2))    # target shorter than k
    t = bytes(rng.choice(b"ACGT") for _ in range(70000))              # many tiles, more than one sort pass
    cases.append((t, mutate(t, 0.05), 13, 2))
    for t, q, k, occ in cases:
        hits, counts = _py_hits_and_shared(t, q, k, occ)
        assert mw.kmer_hits(t, q, k, occ) == hits, (len(t), len(q), k, occ)
        if len(t) >= k and len(q) >= k:
            assert mw.kmer_shared(t, q, k) == counts, (len(t), len(q), k)
    assert mw.lib().mwf_b200_kmer_launches() > 0


@pytest.mark.parametrize("n,p,flag", [(300000, 0.03, 1), (1000000, 0.03, 0)])
def test_chain_large_pair_against_reference(n, p, flag):
    """mwf_wfa_chain on pairs large enough for the device front end and tens of thousands of gap fills, against the
    unmodified reference (oracle/_ref) on the same input; plus a long unrelated insert so that mwf_ksim's branch runs."""
    ref = orc.reference()
    if ref is None:
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    import ctypes
    t, q = synth.make_pair(n, p, 7)
    junk = synth.make_pair(15000, 0.0, 8)[0]
    t = t[:n // 2] + synth.make_pair(12000, 0.0, 9)[0] + t[n // 2:]
    q = q[:len(q) // 2] + junk + q[len(q) // 2:]
    o = mw.opt_init(flag=flag, step=5000 if flag else 0)
    r = mw.wfa_chain(o, t, q)
    ro, rr = orc.make_opt(flag=flag, step=5000 if flag else 0), orc.Rst()
    ref.mwf_wfa_chain(None, ctypes.byref(ro), len(t), t, len(q), q, ctypes.byref(rr))
    assert (r[0], r[1]) == (rr.s, rr.n_cigar)
    assert r[3] == [rr.cigar[i] for i in range(rr.n_cigar)]
    if flag:
        assert mw.cigar2score(o, r[3])[1:] == (len(t), len(q))


def test_cli_matches_reference_output(tmp_path):
    """test-mwf over two FASTA files (two records each): PAF-like lines as the reference prints them (main.c:73-80)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(mw.api.LIB_PATH), "test-mwf")
    assert os.path.exists(exe)
    t3 = ("CAGGGGCAGACTGACACTTCACACGGCCGGGTACTCTAACAGACCTGCAGCTGAGGGTCCT",
          "TAGGGGCAGACTGACACCTCACACGGCCGGGTACTCCTCTGAGACAAAACTTCCAGAGGAACGATCAGACAGCAGCATTCGCGGTTCATGAAAATCCGCTGTTCTG"
          "CAGCCACCGCTGCTGGTACCCAGGCAAACAGGGTCTAGAGTGGACCTTTAGCAAACTCCAACAGACCTGCAGCTGAGGGTCCT")
    f1, f2 = tmp_path / "a.fa", tmp_path / "b.fa"
    f1.write_text(">t3-0 first\n%s\n%s\n>x\nACGT\n" % (t3[0][:30], t3[0][30:]))
    f2.write_text(">t3-1\n%s\n>y\nACCT\n" % t3[1])
    want = {"": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t155", "x\t4\t0\t4\t+\ty\t4\t0\t4\t4"],
            "-c": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t155\t1X16=1X14=128I4=1X24=", "x\t4\t0\t4\t+\ty\t4\t0\t4\t4\t2=1X1="],
            "-cp5": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t155\t1X16=1X14=128I4=1X24=", "x\t4\t0\t4\t+\ty\t4\t0\t4\t4\t2=1X1="],
            "-cu": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t155\t1X16=1X18=128I1X24="],
            "-ct": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t155\t1X16=1X14=128I4=1X24="],
            "-ca": ["t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t272\t1X16=1X18=118I1=10I24="]}
    for flags, lines in want.items():
        for env in ({}, {"MWF_CLI_CHUNK_PAIRS": "1"}):  # one batch for the file; one batch per pair (reader thread ahead of the GPU)
            out = subprocess.run([exe] + ([flags] if flags else []) + [str(f1), str(f2)], capture_output=True, text=True, timeout=120,
                                 env=dict(os.environ, **env))
            assert out.returncode == 0, out.stderr
            got = out.stdout.strip().split("\n")
            assert got[:len(lines)] == lines, (flags, env, got)


def test_tile_geometry_variants(monkeypatch):
    """The tile engine under every cells-per-thread variant, several tile widths / block lengths, and with the batch cut into
    waves: same bits as the CPU checker (the engine reads these settings when a batch is created)."""
    mw.set_kernel(mw.KERNEL_TILE)
    rng = random.Random(7)
    pairs = []
    for i in range(20):
        n = rng.choice([50, 400, 1500, 4000, 9000])
        t = bytes(rng.choice(b"ACGT" if i % 3 else b"ACGTN") for _ in range(n))  # every third pair takes four-bit codes
        pairs.append((t, mutate(rng, t, rng.choice([0.01, 0.05, 0.2]))))
    modes = ({}, {"flag": mw.F_CIGAR}, {"flag": mw.F_CIGAR, "x": 2, "o1": 3, "e1": 1, "o2": 9, "e2": 1}, {"max_iter": 300000})
    want = {i: [orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs] for i, kw in enumerate(modes)}
    for cpt, nt, T, wave in ((4, 256, 64, 0), (4, 128, 32, 7), (2, 256, 32, 0), (2, 256, 64, 3), (1, 512, 64, 0), (1, 256, 24, 0), (4, 64, 16, 0), (2, 96, 12, 5)):
        monkeypatch.setenv("MWF_B200_TILE_CPT", str(cpt))
        monkeypatch.setenv("MWF_B200_TILE_THREADS", str(nt))
        monkeypatch.setenv("MWF_B200_TILE_T", str(T))
        monkeypatch.setenv("MWF_B200_TILE_WAVE", str(wave))
        for i, kw in enumerate(modes):
            with mw.Batch(mw.opt_init(**kw), pairs) as b:
                assert b.kernel_used == mw.KERNEL_TILE, (cpt, nt, T)
                b.upload()
                b.run()
                assert b.fetch() == want[i], (cpt, nt, T, wave, kw)


def test_edge_tiles_on_the_register_resident_step(monkeypatch):
    """Tiles at the band's edges grow the band from the recurrence itself (MWF_B200_TILE_FASTEDGE, default on) or with round 1's
    per-score masks: same bits as the CPU checker both ways -- on pairs of unequal length (the band reaches a corner of the matrix
    and trims take diagonals away), with stops, with traceback, in low-memory mode, on two- and four-bit codes."""
    mw.set_kernel(mw.KERNEL_TILE)
    rng = random.Random(11)
    pairs = []
    for i, (n, m, p) in enumerate(((12000, 7000, 0.03), (5000, 9000, 0.05), (8000, 8000, 0.1), (3000, 2900, 0.02), (700, 6000, 0.05), (6000, 6100, 0.3))):
        t = bytes(rng.choice(b"ACGT" if i % 2 else b"ACGTN") for _ in range(n))
        q = mutate(rng, t, p)
        q = q[:m] if len(q) >= m else q + bytes(rng.choice(b"ACGT") for _ in range(m - len(q)))
        pairs.append((t, q))
    for kw in ({}, {"flag": mw.F_CIGAR}, {"max_iter": 2000000}, {"max_s": 3000}, {"flag": mw.F_CIGAR, "step": 700}):
        want = [orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
        for fe in ("1", "0"):
            monkeypatch.setenv("MWF_B200_TILE_FASTEDGE", fe)
            for T in ("32", "64"):
                monkeypatch.setenv("MWF_B200_TILE_T", T)
                with mw.Batch(mw.opt_init(**kw), pairs) as b:
                    assert b.kernel_used == mw.KERNEL_TILE
                    b.upload()
                    b.run()
                    assert b.fetch() == want, (kw, fe, T)


def test_lowmem_when_the_highmem_pass_does_not_fit(monkeypatch):
    """Low-memory requests run on the tile engine through an unbanded high-memory pass; when its s^2 traceback bytes do not
    fit the arena the checkpoints come from the segmented walk (snapshots + recompute), or, on request, from the reference's
    two-stripe pass 1 on the streaming kernels.  Same bits every way."""
    mw.set_kernel(mw.KERNEL_TILE)
    pairs = synth.make_batch(3, 6000, 0.08, 4242)
    for kw in ({"flag": mw.F_CIGAR, "step": 300}, {"flag": mw.F_CIGAR, "step": 37, "max_s": 1500}, {"flag": mw.F_CIGAR, "step": 5000}):
        want = [orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs]
        for cap, streaming, segp in ((0, 0, None), (200000, 0, None), (200000, 1, None), (0, 0, 256), (0, 0, 1024)):
            monkeypatch.setenv("MWF_B200_TILE_ARENA_MAX", str(cap))
            monkeypatch.setenv("MWF_B200_LOWMEM_STREAMING", str(streaming))
            if segp is None:
                monkeypatch.delenv("MWF_B200_TILE_SEGP", raising=False)
            else:
                monkeypatch.setenv("MWF_B200_TILE_SEGP", str(segp))
            with mw.Batch(mw.opt_init(**kw), pairs) as b:
                assert b.kernel_used == mw.KERNEL_TILE
                b.upload()
                b.run()
                assert b.fetch() == want, (kw, cap, streaming, segp)
                assert (b.kernel_used == mw.KERNEL_TILE) == (not (cap and streaming))


def test_segmented_traceback(monkeypatch):
    """High-memory CIGAR through the segmented traceback (snapshots every P scores, segments recomputed from the end), forced
    here with a tiny P, and reached on its own when the s^2 bytes do not fit the arena: the reference's CIGAR either way."""
    mw.set_kernel(mw.KERNEL_TILE)
    rng = random.Random(11)
    pairs = [(b"ACGTACGT", b"ACGTACGT"), (b"ACGT", b""), (b"A" * 700, b"A" * 300 + b"C" * 500)]
    for i in range(14):
        n = rng.choice([300, 2000, 6000, 12000])
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        pairs.append((t, mutate(rng, t, rng.choice([0.0, 0.02, 0.1, 0.3]))))
    modes = ({"flag": mw.F_CIGAR}, {"flag": mw.F_CIGAR, "x": 2, "o1": 3, "e1": 1, "o2": 9, "e2": 1}, {"flag": mw.F_CIGAR, "max_s": 900},
             {"flag": mw.F_CIGAR, "max_iter": 400000})
    want = [[orc.checker_exact(orc.make_opt(**kw), t, q) for t, q in pairs] for kw in modes]
    for segp, cap, wave in ((256, 0, 0), (512, 0, 5), (1024, 0, 0), (None, 300000, 0)):
        if segp is None:
            monkeypatch.delenv("MWF_B200_TILE_SEGP", raising=False)
        else:
            monkeypatch.setenv("MWF_B200_TILE_SEGP", str(segp))
        monkeypatch.setenv("MWF_B200_TILE_ARENA_MAX", str(cap))
        monkeypatch.setenv("MWF_B200_TILE_WAVE", str(wave))
        for kw, w in zip(modes, want):
            assert mw.wfa_exact_batch(mw.opt_init(**kw), pairs) == w, (segp, cap, wave, kw)
    # single pairs take the latency geometry, and several segments are recomputed per pass on virtual slots
    monkeypatch.setenv("MWF_B200_TILE_SEGP", "256")
    monkeypatch.setenv("MWF_B200_TILE_ARENA_MAX", "0")
    for (t, q), w in zip(pairs[3:9], want[0][3:9]):
        assert mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR), t, q) == w
    for par in (1, 2, 5):  # segments per pass; a batch of up to 8 pairs is spread over par x pairs virtual slots
        monkeypatch.setenv("MWF_B200_TILE_SEGPAR", str(par))
        assert mw.wfa_exact_batch(mw.opt_init(flag=mw.F_CIGAR), pairs[9:15]) == want[0][9:15], par
        assert mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR, max_s=900), *pairs[12]) == want[2][12], par


def test_concurrent_host_threads():
    """The library keeps no alignment state between calls (stateless, one km per thread): several host threads calling
    the entry points at once -- batches on the tile engine and on the streaming kernel, mwf_wfa_chain with the device front end,
    single pairs with CIGAR -- get the results of the same calls made one after the other."""
    import threading
    mw.set_kernel(mw.KERNEL_AUTO)
    jobs = []
    for i in range(8):
        if i % 4 == 0:
            pairs = synth.make_batch(6, 12000, 0.04, 100 + i)
            jobs.append((lambda pairs=pairs: mw.wfa_exact_batch(mw.opt_init(), pairs)))
        elif i % 4 == 1:
            pairs = synth.make_batch(300, 200, 0.05, 200 + i)
            jobs.append((lambda pairs=pairs: mw.wfa_exact_batch(mw.opt_init(flag=mw.F_CIGAR), pairs)))
        elif i % 4 == 2:
            t, q = synth.make_pair(60000, 0.03, 300 + i)
            jobs.append((lambda t=t, q=q: mw.wfa_chain(mw.opt_init(flag=mw.F_CIGAR, step=5000), t, q)))
        else:
            t, q = synth.make_pair(15000, 0.05, 400 + i)
            jobs.append((lambda t=t, q=q: mw.wfa_exact(mw.opt_init(flag=mw.F_CIGAR, step=700), t, q)))
    want = [j() for j in jobs]
    for _ in range(3):
        got = [None] * len(jobs)
        errs = []

        def work(k):
            try:
                got[k] = jobs[k]()
            except Exception as e:  # pragma: no cover
                errs.append(e)
        th = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
        for x in th:
            x.start()
        for x in th:
            x.join()
        assert not errs and got == want


def test_arena_overflow_predicted_from_shared_kmers(monkeypatch):
    """A few very long pairs whose s^2 traceback bytes cannot fit skip the all-at-once attempt: the shared 13-mer fraction
    (mwf_b200_kmer_shared) gives a low estimate of s beforehand.  Forced here on a 12 kb pair with a 300 kB arena: the launch
    count shows that no all-at-once attempt ran, the CIGAR is the reference's; a near-identical pair is not predicted to overflow."""
    mw.set_kernel(mw.KERNEL_TILE)
    monkeypatch.setenv("MWF_B200_TILE_PREDICT_MINLEN", "1000")
    monkeypatch.setenv("MWF_B200_TILE_ARENA_MAX", "300000")
    rng = random.Random(5)
    t = bytes(rng.choice(b"ACGT") for _ in range(12000))
    o = mw.opt_init(flag=mw.F_CIGAR)
    launches = {}
    for name, q in (("far", mutate(rng, t, 0.1)), ("near", mutate(rng, t, 0.002))):
        want = orc.checker_exact(orc.make_opt(flag=1), t, q)
        for predict in ("1", "0"):
            monkeypatch.setenv("MWF_B200_TILE_PREDICT", predict)
            with mw.Batch(o, [(t, q)]) as b:
                b.upload()
                b.run()
                assert b.fetch() == [want], (name, predict)
                launches[name, predict] = b.launches
    assert launches["far", "1"] < launches["far", "0"]      # the failed attempt is gone
    assert launches["near", "1"] == launches["near", "0"]   # s^2 fits: nothing changes


def test_reference_cli_linked_against_this_library(tmp_path):
    """the reference's own, unmodified main.c, compiled against include/miniwfa.h and linked with
    libminiwfa_b200.so in place of miniwfa.o kalloc.o mwf-dbg.o (oracle/_ref/test-mwf-dropin, built where /root/reference
    exists), prints the reference's lines."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "test-mwf-dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/test-mwf-dropin was not built (needs /root/reference at build time)")
    t3 = ("CAGGGGCAGACTGACACTTCACACGGCCGGGTACTCTAACAGACCTGCAGCTGAGGGTCCT",
          "TAGGGGCAGACTGACACCTCACACGGCCGGGTACTCCTCTGAGACAAAACTTCCAGAGGAACGATCAGACAGCAGCATTCGCGGTTCATGAAAATCCGCTGTTCTG"
          "CAGCCACCGCTGCTGGTACCCAGGCAAACAGGGTCTAGAGTGGACCTTTAGCAAACTCCAACAGACCTGCAGCTGAGGGTCCT")
    f1, f2 = tmp_path / "t3-0.fa", tmp_path / "t3-1.fa"
    f1.write_text(">t3-0\n%s\n" % t3[0])
    f2.write_text(">t3-1\n%s\n" % t3[1])
    head = "t3-0\t61\t0\t61\t+\tt3-1\t189\t0\t189\t"
    want = {"": "155", "-c": "155\t1X16=1X14=128I4=1X24=", "-cp5": "155\t1X16=1X14=128I4=1X24=", "-cK": "155\t1X16=1X14=128I4=1X24=",
            "-ct": "155\t1X16=1X14=128I4=1X24=", "-cu": "155\t1X16=1X18=128I1X24=", "-ca": "272\t1X16=1X18=118I1=10I24=",
            "-ce": "128\t21I2=6I2=9I1=1I1=10I1=2I1=1I1=1I2=2I1=9I1=1I1=9I1=6I1=4I1=3I1=1I1=5I2=6I1=4I1=2I1=1I3=3I1=3I1=4I1=1I2=2I1=1I1=2I1=6I2=2I24="}
    for flags, tail in want.items():
        out = subprocess.run([exe] + ([flags] if flags else []) + [str(f1), str(f2)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        assert out.stdout.strip() == head + tail, (flags, out.stdout)


def test_batch_spread_over_devices():
    """mwf_wfa_exact_batch() through the C ABI with the pairs dealt out over two devices by one host process: the
    results, CIGARs from the caller's allocator included, in input order and equal to the one-device call."""
    if mw.device_count() < 2:
        pytest.skip("needs two visible devices (gpurun --gpus 2)")
    mw.set_kernel(mw.KERNEL_AUTO)
    pairs = synth.make_batch(6, 30000, 0.05, 500) + synth.make_batch(5, 12000, 0.1, 600) + [(b"ACGT", b"ACGA"), (b"", b"A")]
    for kw in ({}, {"flag": mw.F_CIGAR}, {"flag": mw.F_CIGAR, "step": 2000}):
        mw.set_devices(1)
        one = mw.wfa_exact_batch(mw.opt_init(**kw), pairs)
        mw.set_devices(2)
        try:
            two = mw.wfa_exact_batch(mw.opt_init(**kw), pairs)
        finally:
            mw.set_devices(0)
        assert two == one, kw
    t, q = pairs[0]
    assert one[0] == orc.checker_exact(orc.make_opt(flag=1, step=2000), t, q)
