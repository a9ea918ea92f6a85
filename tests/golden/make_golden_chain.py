"""Generate tests/golden/golden_chain.json from the UNMODIFIED reference (oracle/_ref): mwf_wfa_chain and
mwf_wfa_auto on inputs that exercise every branch of the gap-fill loop (miniwfa.c:861-889).
Run in the build container only:  python tests/golden/make_golden_chain.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from miniwfa_b200 import synth  # noqa: E402
from miniwfa_b200.api import cigar_string  # noqa: E402


def rnd(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))


def build_inputs(case):
    """Inputs are described by a recipe so that the fixture stays small."""
    kind = case["recipe"][0]
    if kind == "synth":
        return synth.make_pair(*case["recipe"][1:])
    rng = random.Random(case["recipe"][1])
    if kind == "blocks":  # shared flanks around unrelated / missing middles
        _, _, flank, mid_t, mid_q, div = case["recipe"]
        a, b = synth.make_pair(flank, div, 7000 + case["recipe"][1]), synth.make_pair(flank, div, 7100 + case["recipe"][1])
        return a[0] + rnd(rng, mid_t) + b[0], a[1] + rnd(rng, mid_q) + b[1]
    if kind == "text":
        return case["recipe"][2].encode("latin-1"), case["recipe"][3].encode("latin-1")
    raise ValueError(kind)


def main():
    if orc.reference() is None:
        sys.exit("oracle/_ref is not built (needs /root/reference): make -C oracle ref")
    C = 1
    cases = []

    def add(name, recipe, fn, **opt):
        cases.append({"name": name, "recipe": recipe, "fn": fn, "opt": opt})

    for fn in ("mwf_wfa_chain", "mwf_wfa_auto"):
        tag = "chain" if fn.endswith("chain") else "auto"
        for n, p, idx in ((200, 0.05, 1), (3000, 0.02, 2), (3000, 0.15, 3), (20000, 0.05, 4), (50000, 0.01, 5)):
            add("%s-n%d-p%g" % (tag, n, p), ["synth", n, p, 5000 + idx], fn)
            add("%s-n%d-p%g-c" % (tag, n, p), ["synth", n, p, 5000 + idx], fn, flag=C)
    # the exact leg of auto runs out of its 10^8-cell budget: chaining with step = 5000 (miniwfa.c:903-907)
    add("auto-n100k-p5-c", ["synth", 100000, 0.05, 0], "mwf_wfa_auto", flag=C)
    add("auto-n100k-p5", ["synth", 100000, 0.05, 0], "mwf_wfa_auto")
    add("chain-n100k-p5-cp5000", ["synth", 100000, 0.05, 1], "mwf_wfa_chain", flag=C, step=5000)
    # two long unrelated stretches between anchors: the 2-gap shortcut (:869-874)
    add("chain-unrelated-12k-c", ["blocks", 11, 3000, 12000, 12500, 0.02], "mwf_wfa_chain", flag=C)
    add("chain-unrelated-12k", ["blocks", 11, 3000, 12000, 12500, 0.02], "mwf_wfa_chain")
    # pure deletion / pure insertion between anchors (:882-888), also recorded in score-only mode
    add("chain-del-c", ["blocks", 12, 2000, 500, 0, 0.0], "mwf_wfa_chain", flag=C)
    add("chain-del", ["blocks", 12, 2000, 500, 0, 0.0], "mwf_wfa_chain")
    add("chain-ins-c", ["blocks", 13, 2000, 0, 700, 0.0], "mwf_wfa_chain", flag=C)
    add("chain-ins", ["blocks", 13, 2000, 0, 700, 0.0], "mwf_wfa_chain")
    add("chain-short-unrelated-c", ["blocks", 14, 1500, 300, 200, 0.03], "mwf_wfa_chain", flag=C)
    # sequences shorter than k, no anchors at all, non-ACGT bytes, repeats above max_occ
    add("chain-tiny-c", ["text", 0, "ACGTAC", "ACGGTAC"], "mwf_wfa_chain", flag=C)
    add("chain-noanchor-c", ["text", 0, "ACGT" * 30, "TTGCA" * 25], "mwf_wfa_chain", flag=C)
    add("chain-N-c", ["text", 0, "ACGTTGCATGCAAGCTNNNNNNACGTGCATGCAGTCAGTCAGTACGTAGCTAGCTAGCATCGATCGATCAGCTAGCATGCATCGAT" * 3,
         "ACGTTGCATGCAAGCTNNNNNACGTGCATGCAGTCAGTCAGTACGTAGCTAGCTAGCATCGATCGATCAGCTAGCATGCATCGAT" * 3], "mwf_wfa_chain", flag=C)
    add("chain-lowercase-c", ["text", 0, "acgttgcatgcaagctacgtgcatgcagtcagtcagtacgtagctagc", "ACGTTGCATGCAAGCTACGTGCATGCAGTCAGTCAGTACGTAGCTAGC"],
        "mwf_wfa_chain", flag=C)
    add("chain-maxocc1-c", ["synth", 5000, 0.03, 5050], "mwf_wfa_chain", flag=C, max_occ=1, min_len=60)
    add("chain-k11-c", ["synth", 5000, 0.08, 5051], "mwf_wfa_chain", flag=C, kmer=11, min_len=20)
    add("chain-edit-c", ["synth", 5000, 0.05, 5052], "mwf_wfa_chain", flag=C, x=1, o1=0, o2=0, e1=1, e2=1)
    add("chain-nokalloc-c", ["synth", 5000, 0.05, 5053], "mwf_wfa_chain", flag=C | 2)

    for c in cases:
        t, q = build_inputs(c)
        o = orc.make_opt(**c["opt"])
        s, nc, ni, cig = orc.reference_exact(o, t, q, c["fn"])
        c["expect"] = {"s": s, "n_cigar": nc, "n_iter": ni, "cigar": cigar_string(cig), "tl": len(t), "ql": len(q)}
        print(c["name"], s, nc, ni, len(t), len(q))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_chain.json")
    with open(path, "w") as f:
        json.dump({"reference": "lh3/miniwfa @ 66770a3 (oracle/_ref)", "cases": cases}, f, indent=0)
    print("wrote %d cases -> %s (%d bytes)" % (len(cases), path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
