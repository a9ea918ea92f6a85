"""Generate tests/golden/golden_large.json from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile): the BASELINE.json configurations at their full sizes (BASELINE.md 3.3 / 3.5 "parity gate").  Build container
only -- these are minutes to hours of one host core each and up to ~40 GB of host memory:

    python tests/golden/make_golden_large.py [case ...]        # default: every case not yet in the file

A CIGAR of 1e4..3e5 words is recorded as n_cigar + sha1 over its little-endian uint32 words (`cigar_sha1`); the config-3
batch as the (s, n_iter) of each of its 1024 pairs plus a sha1 over the list.  Every entry records the reference's wall
time on this container's CPU (one thread), which is NOT the box of the bench -- a context number only.
"""
import hashlib
import json
import os
import re
import struct
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402
from miniwfa_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_large.json")

# name -> (synth (n, p, index), option overrides, what it pins)
CASES = {
    "config2-c": ((150000, 0.038, 900000), {"flag": 1},
                  "BASELINE config 2 surrogate: 150 kb pair, high-memory CIGAR (test-mwf -c)"),
    "config2-cp5000": ((150000, 0.038, 900000), {"flag": 1, "step": 5000}, "the same pair in low-memory mode (test-mwf -cp5000)"),
    "n1m-p3-c": ((1000000, 0.03, 77), {"flag": 1},
                 "1 Mb / 3 % pair, high-memory CIGAR: 2e10 traceback bytes in the reference; the segmented traceback must give this CIGAR"),
    "n1m-p3-cp5000": ((1000000, 0.03, 77), {"flag": 1, "step": 5000}, "the same pair, low-memory mode"),
    "config4-cp5000": ((5000000, 0.0097, 424242), {"flag": 1, "step": 5000},
                       "BASELINE config 4 surrogate: 5 Mb pair, s ~ 231 k, low-memory mode (test-mwf -cp5000)"),
    "config5-cp5000": ((5000000, 0.03, 424242), {"flag": 1, "step": 5000},
                       "BASELINE config 5 surrogate: 5 Mb pair at 3 %, s ~ 713 k.  The reference cannot hold its s^2 = 5e11 high-memory "
                       "traceback bytes on this host, so it is run with -cp5000 (BASELINE.md 3.3: same CIGAR, SURVEY 7.3-6); n_iter is "
                       "therefore pass 2's, not the high-memory count"),
}


def cigar_sha1(words):
    return hashlib.sha1(struct.pack("<%dI" % len(words), *words)).hexdigest()


def pairs_sha1(rows):
    return hashlib.sha1(("".join("%d,%d;" % (s, ni) for s, ni in rows)).encode()).hexdigest()


def run_case(name):
    spec, opt, what = CASES[name]
    t, q = synth.make_pair(*spec)
    o = orc.make_opt(**opt)
    t0 = time.perf_counter()
    s, nc, ni, cig = orc.reference_exact(o, t, q)
    dt = time.perf_counter() - t0
    return {"name": name, "what": what, "synth": list(spec), "opt": opt, "tl": len(t), "ql": len(q),
            "expect": {"s": s, "n_cigar": nc, "n_iter": ni, "cigar_sha1": cigar_sha1(cig)},
            "reference_seconds_here": round(dt, 2)}


def run_config3(n_pairs=1024, threads=6):
    """(s, n_iter) of every pair of the config-3 batch (100 kb, p = 0.05, indices 0..1023), score-only."""
    def one(i):
        t, q = synth.make_pair(100000, 0.05, i)
        r = orc.reference_exact(orc.make_opt(), t, q)
        return (r[0], r[2], max(len(t), len(q)))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        rows = list(ex.map(one, range(n_pairs)))
    dt = time.perf_counter() - t0
    return {"name": "config3", "what": "BASELINE config 3: 1024 x 100 kb pairs, p = 0.05, indices 0..1023, score-only",
            "synth": [100000, 0.05, 0], "n_pairs": n_pairs, "opt": {},
            "expect": {"s_n_iter": [[r[0], r[1]] for r in rows], "sha1": pairs_sha1([(r[0], r[1]) for r in rows]),
                       "sum_ns": sum(r[0] * r[2] for r in rows), "sum_n_iter": sum(r[1] for r in rows)},
            "reference_seconds_here": round(dt, 2), "threads": threads}


def main():
    if orc.reference() is None:
        sys.exit("oracle/_ref is not built (needs /root/reference): make -C oracle ref")
    doc = {"cases": []}
    if os.path.exists(OUT):
        doc = json.load(open(OUT))
    have = {c["name"] for c in doc["cases"]}
    want = sys.argv[1:] or [n for n in list(CASES) + ["config3"] if n not in have]
    for name in want:
        c = run_config3() if name == "config3" else run_case(name)
        doc["cases"] = [x for x in doc["cases"] if x["name"] != name] + [c]
        with open(OUT + ".tmp", "w") as f:
            txt = json.dumps(doc, indent=1)  # the 1024 (s, n_iter) rows of config 3 on one line
            txt = re.sub(r'"s_n_iter": \[(?:.|\n)*?\]\s*\]',
                         lambda m: '"s_n_iter": ' + json.dumps(json.loads(m.group(0)[len('"s_n_iter": '):]), separators=(",", ":")), txt)
            f.write(txt + "\n")
        os.replace(OUT + ".tmp", OUT)
        e = c["expect"]
        print(name, {k: v for k, v in e.items() if k != "s_n_iter"}, c["reference_seconds_here"], "s", flush=True)


if __name__ == "__main__":
    main()
