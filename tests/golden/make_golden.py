"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built from /root/reference
by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Every case records the inputs (inline bytes as latin-1, or the synthetic (n, p, index) triple), the option
overrides, which reference entry point was called, and the reference's (s, n_cigar, n_iter, CIGAR string).
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from miniwfa_b200 import synth  # noqa: E402
from miniwfa_b200.api import cigar_string  # noqa: E402

PRESETS = {
    "default": {},
    "affine": {"o2": 4, "e2": 2},                                  # test-mwf -a
    "edit": {"x": 1, "o1": 0, "o2": 0, "e1": 1, "e2": 1},          # test-mwf -e
    "rand1": {"x": 6, "o1": 2, "e1": 3, "o2": 24, "e2": 1},
    "rand2": {"x": 3, "o1": 5, "e1": 1, "o2": 1, "e2": 2},
}


def read_fa(path):
    seq = []
    for line in open(path):
        if not line.startswith(">"):
            seq.append(line.strip())
    return "".join(seq).encode()


def run(case):
    if "synth" in case:
        t, q = synth.make_pair(*case["synth"])
    else:
        t, q = case["t"].encode("latin-1"), case["q"].encode("latin-1")
    o = orc.make_opt(**case["opt"])
    s, nc, ni, cig = orc.reference_exact(o, t, q, case.get("fn", "mwf_wfa_exact"))
    case["expect"] = {"s": s, "n_cigar": nc, "n_iter": ni, "cigar": cigar_string(cig)}
    return case


def main():
    if orc.reference() is None:
        sys.exit("oracle/_ref is not built (needs /root/reference): make -C oracle ref")
    cases = []
    t3 = [read_fa("/root/reference/test/t3-%d.fa" % i).decode() for i in (0, 1)]
    C = 1

    def add(name, t, q, fn="mwf_wfa_exact", **opt):
        cases.append({"name": name, "t": t, "q": q, "fn": fn, "opt": opt})

    # t3 pair under the CLI flag combinations of SURVEY.md §4
    add("t3-score", t3[0], t3[1])
    add("t3-c", t3[0], t3[1], flag=C)
    for p in (1, 5, 37, 5000):
        add("t3-cp%d" % p, t3[0], t3[1], flag=C, step=p)
    add("t3-cK", t3[0], t3[1], flag=C | 2)
    add("t3-ct", t3[0], t3[1], fn="mwf_wfa_auto", flag=C)
    add("t3-cu", t3[0], t3[1], fn="mwf_wfa_chain", flag=C)
    add("t3-ca", t3[0], t3[1], flag=C, **PRESETS["affine"])
    add("t3-ce", t3[0], t3[1], flag=C, **PRESETS["edit"])
    add("t3-swapped-c", t3[1], t3[0], flag=C)
    add("t3-swapped-cp7", t3[1], t3[0], flag=C, step=7)
    # small known answers
    kats = [("ACGT", "ACGT"), ("ACGT", "ACCT"), ("ACGT", ""), ("", "ACGT"), ("A", "C"), ("ACGTACGTACGT", "ACGTCGTACGT"),
            ("A" * 30, "A" * 10), ("ACGTNNNNACGT", "acgtNNNNacgt"), ("AAAA", "CCCC"), ("A", "A"), ("", "A"), ("A", ""),
            ("ACGTACGTAC" * 20, "ACGTACGTAC" * 20), ("ACGT" * 50, "TGCA" * 50), ("\x00\x01\xff\x00", "\x00\xff\x01\x00")]
    for i, (t, q) in enumerate(kats):
        add("kat%d-score" % i, t, q)
        add("kat%d-c" % i, t, q, flag=C)
        add("kat%d-cp3" % i, t, q, flag=C, step=3)
    add("empty-score", "", "")
    add("stop-max_s", "AAAA", "CCCC", max_s=3)
    add("stop-max_iter", "AAAA", "CCCC", max_iter=10)
    add("stop-max_s-c", "AAAA", "CCCC", flag=C, max_s=3)
    # random small pairs: every preset x every mode
    rng = random.Random(20261017)
    idx = 1000
    for preset, pv in PRESETS.items():
        for n, p in ((60, 0.0), (150, 0.05), (300, 0.15), (300, 0.4), (1000, 0.05), (3000, 0.15), (3000, 0.01)):
            for mode in ("score", "c", "cp1", "cp7", "cp37", "cp5000"):
                if mode == "cp1" and n > 300:
                    continue
                opt = dict(pv)
                if mode != "score":
                    opt["flag"] = C
                if mode.startswith("cp"):
                    opt["step"] = int(mode[2:])
                cases.append({"name": "%s-n%d-p%g-%s" % (preset, n, p, mode), "synth": [n, p, idx], "opt": opt})
                idx += 1
    # unequal lengths / unrelated sequences
    for i in range(12):
        tl, ql = rng.choice([0, 1, 7, 64, 200, 500]), rng.choice([0, 1, 9, 100, 333, 800])
        t = "".join(rng.choice("ACGT") for _ in range(tl))
        q = "".join(rng.choice("ACGT") for _ in range(ql))
        if tl == 0 and ql == 0:
            continue
        add("unrelated%d-c" % i, t, q, flag=C)
        add("unrelated%d-cp5-edit" % i, t, q, flag=C, step=5, **PRESETS["edit"])
    # stops on mid-size pairs
    for i, kw in enumerate(({"max_s": 50}, {"max_iter": 5000}, {"max_s": 400, "flag": C}, {"max_iter": 200000, "flag": C},
                            {"max_iter": 3000, "flag": C, "step": 7})):
        cases.append({"name": "stop%d" % i, "synth": [2000, 0.1, 2000 + i], "opt": kw})
    # larger anchors (seconds on the CPU)
    cases.append({"name": "n20k-p5-score", "synth": [20000, 0.05, 3000], "opt": {}})
    cases.append({"name": "n20k-p5-c", "synth": [20000, 0.05, 3000], "opt": {"flag": C}})
    cases.append({"name": "n20k-p5-cp500", "synth": [20000, 0.05, 3000], "opt": {"flag": C, "step": 500}})
    cases.append({"name": "n100k-p5-score", "synth": [100000, 0.05, 0], "opt": {}})
    cases.append({"name": "n30k-p2-auto-c", "synth": [30000, 0.02, 3001], "fn": "mwf_wfa_auto", "opt": {"flag": C}})

    out = [run(c) for c in cases]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump({"reference": "lh3/miniwfa @ 66770a3 (oracle/_ref, gcc -O3 -march=native)", "cases": out}, f, indent=0)
    print("wrote %d cases -> %s (%d bytes)" % (len(out), path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
