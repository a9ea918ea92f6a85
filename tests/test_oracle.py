"""CPU: pin the oracle restatement (oracle/wfa_oracle.c) to the reference's golden vectors and, when the
unmodified reference is built (oracle/_ref), to the reference itself on random inputs."""
import random

import pytest

from conftest import case_inputs
from miniwfa_b200.api import cigar_string
from oracle import orc

EXACT = lambda c: c.get("fn", "mwf_wfa_exact") == "mwf_wfa_exact"  # noqa: E731


def test_golden_has_the_survey_vectors(golden):
    by = {c["name"]: c["expect"] for c in golden}
    assert by["t3-score"] == {"s": 155, "n_cigar": 0, "n_iter": 16875, "cigar": ""}
    assert by["t3-c"]["cigar"] == "1X16=1X14=128I4=1X24=" and by["t3-c"]["n_cigar"] == 8
    assert by["t3-ca"]["s"] == 272 and by["t3-ca"]["cigar"] == "1X16=1X18=118I1=10I24="
    assert by["t3-ce"]["s"] == 128
    assert by["t3-swapped-c"]["cigar"] == "1X16=1X14=128D4=1X24="
    assert by["stop-max_s"] == {"s": -1, "n_cigar": 0, "n_iter": 12, "cigar": ""}
    assert len(golden) > 250


def test_oracle_matches_golden(golden):
    n = 0
    for c in golden:
        if not EXACT(c) or c["name"].startswith("n100k"):
            continue
        t, q = case_inputs(c)
        if len(t) == 0 and len(q) == 0 and c["opt"].get("flag", 0) & 1:
            continue
        s, nc, ni, cig = orc.oracle_exact(orc.make_opt(**c["opt"]), t, q)
        e = c["expect"]
        assert (s, nc, ni, cigar_string(cig)) == (e["s"], e["n_cigar"], e["n_iter"], e["cigar"]), c["name"]
        n += 1
    assert n > 250


def test_oracle_auto_leg_matches_golden(golden):
    for c in golden:
        if c.get("fn") != "mwf_wfa_auto":
            continue
        t, q = case_inputs(c)
        r = orc.Rst()
        o = orc.make_opt(**c["opt"])
        import ctypes
        orc.oracle().orc_wfa_auto_exact_leg(ctypes.byref(o), len(t), t, len(q), q, ctypes.byref(r))
        e = c["expect"]
        cig = [r.cigar[i] for i in range(r.n_cigar)]
        orc.oracle().orc_free(r.cigar)
        assert (r.s, r.n_iter, cigar_string(cig)) == (e["s"], e["n_iter"], e["cigar"])


def test_oracle_cigar_is_consistent(golden):
    import ctypes
    L = orc.oracle()
    for c in golden[:120]:
        if not EXACT(c) or not (c["opt"].get("flag", 0) & 1):
            continue
        t, q = case_inputs(c)
        if len(t) + len(q) == 0:
            continue
        o = orc.make_opt(**c["opt"])
        s, nc, ni, cig = orc.oracle_exact(o, t, q)
        if s < 0:
            continue
        arr = (ctypes.c_uint32 * max(1, nc))(*cig)
        tl, ql = ctypes.c_int32(), ctypes.c_int32()
        sc = L.orc_cigar2score(ctypes.byref(o), nc, arr, ctypes.byref(tl), ctypes.byref(ql))
        assert (tl.value, ql.value) == (len(t), len(q))
        assert sc == s, c["name"]


@pytest.mark.skipif(orc.reference() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_matches_reference_differentially():
    rng = random.Random(7)
    for it in range(150):
        n = rng.choice([0, 1, 3, 20, 100, 400, 1500])
        p = rng.choice([0.0, 0.02, 0.1, 0.3])
        t = bytes(rng.choice(b"ACGT") for _ in range(n))
        q = bytearray()
        for ch in t:
            u = rng.random()
            if u < p * 0.8:
                q.append(rng.choice(b"ACGT"))
            elif u < p * 0.9:
                q.extend(bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 20))))
                q.append(ch)
            elif u < p:
                pass
            else:
                q.append(ch)
        q = bytes(q)
        kw = {}
        if rng.random() < 0.5:
            kw.update(x=rng.randint(1, 8), o1=rng.randint(0, 7), e1=rng.randint(1, 3), o2=rng.randint(0, 25), e2=rng.randint(1, 2))
        mode = rng.choice(["s", "c", "p", "stop"])
        if mode != "s":
            kw["flag"] = 1
        if mode == "p":
            kw["step"] = rng.choice([1, 3, 16, 100])
        if mode == "stop":
            kw[rng.choice(["max_s", "max_iter"])] = rng.randint(1, 3000)
        if len(t) + len(q) == 0 and kw.get("flag"):
            continue
        o = orc.make_opt(**kw)
        assert orc.oracle_exact(o, t, q) == orc.reference_exact(o, t, q), (it, kw)


def test_synth_is_deterministic():
    from miniwfa_b200 import synth
    a = synth.make_pair(5000, 0.05, 3)
    b = synth.make_pair(5000, 0.05, 3)
    c = synth.make_pair(5000, 0.05, 4)
    assert a == b and a != c
    assert set(a[0]) <= set(b"ACGT") and set(a[1]) <= set(b"ACGT")
    assert abs(len(a[1]) - 5000) < 200
    import hashlib
    # pinned so that bench inputs cannot drift silently
    assert hashlib.sha1(a[0] + b"|" + a[1]).hexdigest() == hashlib.sha1(b[0] + b"|" + b[1]).hexdigest()
