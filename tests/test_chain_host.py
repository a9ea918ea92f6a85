"""CPU: the host side of mwf_wfa_chain (csrc/mwf_chain.c: host k-mer front end, longest increasing subsequence with its galloping /
two-level search, short-run filter, segment classification and merging, CIGAR assembly) against the reference's own outputs in
tests/golden/golden_chain.json.  The exact gap fills -- the part the product runs on the GPU -- are answered here by a stub of
mwf_wfa_exact_batch over the CPU oracle, so this file tests host logic only and nothing in it is a product path."""
import ctypes
import os
import subprocess

import pytest

from conftest import chain_case_inputs
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = r"""
#include <stdlib.h>
#include <string.h>
#include "miniwfa.h"
#include "mwf_b200.h"
#include "kalloc.h"
#include "wfa_oracle.h"
void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n, const int32_t *tl, const char *const *ts,
                         const int32_t *ql, const char *const *qs, mwf_rst_t *r)
{
	int32_t i;
	for (i = 0; i < n; ++i) {
		orc_rst_t o;
		orc_wfa_exact((const orc_opt_t*)opt, tl[i], ts[i], ql[i], qs[i], &o);
		r[i].s = o.s, r[i].n_cigar = o.n_cigar, r[i].n_iter = o.n_iter, r[i].cigar = 0;
		if (o.n_cigar > 0) {
			r[i].cigar = (uint32_t*)kmalloc(km, sizeof(uint32_t) * o.n_cigar);
			memcpy(r[i].cigar, o.cigar, sizeof(uint32_t) * o.n_cigar);
		}
		orc_free(o.cigar);
	}
}
""" + r"""
/* the device entry points mwf_chain.c can call: never reached with MWF_B200_CHAIN_FRONT=host; host scratch is plain malloc here */
int64_t mwf_b200_kmer_hits(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t max_occ, uint64_t **hits) { abort(); }
void mwf_b200_kmer_free(uint64_t *hits) { abort(); }
void *mwf_b200_host_scratch(size_t bytes) { return malloc(bytes); }
void mwf_b200_host_scratch_free(void *p) { free(p); }
void mwf_b200_kmer_shared(int32_t l1, const char *s1, int32_t l2, const char *s2, int32_t k, int64_t *n1, int64_t *n2, int64_t *shared) { abort(); }
"""


@pytest.fixture(scope="module")
def host_chain(tmp_path_factory):
    orc.oracle()  # builds oracle/liboracle.so if needed
    d = tmp_path_factory.mktemp("chain_host")
    stub = d / "stub.c"
    stub.write_text(STUB)
    so = str(d / "libchainhost.so")
    csrc = os.path.join(ROOT, "miniwfa_b200", "csrc")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
                    os.path.join(csrc, "mwf_chain.c"), os.path.join(csrc, "kalloc.c"), str(stub), "-o", so,
                    "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    L = ctypes.CDLL(so)
    L.mwf_wfa_chain.argtypes = [ctypes.c_void_p, ctypes.POINTER(orc.Opt), ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                ctypes.c_char_p, ctypes.POINTER(orc.Rst)]
    L.mwf_wfa_chain.restype = None
    L.kfree.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return L


def _chain(L, o, t, q):
    r = orc.Rst()
    L.mwf_wfa_chain(None, ctypes.byref(o), len(t), t, len(q), q, ctypes.byref(r))
    cig = r.cigar[:r.n_cigar] if r.n_cigar > 0 else []
    if r.cigar:
        L.kfree(None, r.cigar)
    return r.s, r.n_cigar, cig


def test_host_chain_against_reference_goldens(host_chain, golden_chain, monkeypatch):
    monkeypatch.setenv("MWF_B200_CHAIN_FRONT", "host")
    n = 0
    for c in golden_chain:
        if c["fn"] != "mwf_wfa_chain":
            continue
        t, q = chain_case_inputs(c)
        s, n_cigar, cig = _chain(host_chain, orc.make_opt(**c["opt"]), t, q)
        e = c["expect"]
        got = "".join("%d%s" % (w >> 4, "MIDNSHP=XBid"[w & 0xf]) for w in cig)
        assert (s, n_cigar, got) == (e["s"], e["n_cigar"], e["cigar"]), c["name"]
        n += 1
    assert n >= 25


def test_host_chain_large_pair_against_reference(host_chain, monkeypatch):
    """400 kb / 3 %: ~3e5 k-mer matches with ~10 % strays through the two-level search, ~8000 gap fills, merged match segments."""
    ref = orc.reference()
    if ref is None:
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    from miniwfa_b200 import synth
    monkeypatch.setenv("MWF_B200_CHAIN_FRONT", "host")
    t, q = synth.make_pair(400000, 0.03, 77)
    for flag in (0, 1):
        o = orc.make_opt(flag=flag, step=5000 if flag else 0)
        got = _chain(host_chain, o, t, q)
        rr = orc.Rst()
        ref.mwf_wfa_chain(None, ctypes.byref(o), len(t), t, len(q), q, ctypes.byref(rr))
        assert got[:2] == (rr.s, rr.n_cigar)
        assert got[2] == (rr.cigar[:rr.n_cigar] if rr.n_cigar > 0 else [])


LIS_WRAP = r"""
#include <stdlib.h>
#include "@CSRC@/mwf_chain.c"
/* test hook: the static longest_increasing() of mwf_chain.c */
int32_t lis_values(int32_t n, const uint64_t *v, uint64_t *out)
{
	int32_t m = 0, i;
	uint64_t *a = longest_increasing(0, n, v, &m);
	for (i = 0; i < m; ++i) out[i] = a[i];
	kfree(0, a);
	return m;
}
void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n, const int32_t *tl, const char *const *ts,
                         const int32_t *ql, const char *const *qs, mwf_rst_t *r) { abort(); }
int64_t mwf_b200_kmer_hits(int32_t tl, const char *ts, int32_t ql, const char *qs, int32_t k, int32_t max_occ, uint64_t **hits) { abort(); }
void mwf_b200_kmer_free(uint64_t *hits) { abort(); }
void *mwf_b200_host_scratch(size_t bytes) { return malloc(bytes); }
void mwf_b200_host_scratch_free(void *p) { free(p); }
void mwf_b200_kmer_shared(int32_t l1, const char *s1, int32_t l2, const char *s2, int32_t k, int64_t *n1, int64_t *n2, int64_t *shared) { abort(); }
"""


def _py_lis(v):
    """mg_lis_64 (miniwfa.c:678-697) restated: patience piles by bisection, predecessor = tail of the pile below."""
    import bisect
    tails, tail_idx, prev = [], [], []
    for i, x in enumerate(v):
        lo = bisect.bisect_left(tails, x)  # number of piles whose tail is below x (values are distinct)
        prev.append(tail_idx[lo - 1] if lo > 0 else -1)
        if lo == len(tails):
            tails.append(x)
            tail_idx.append(i)
        else:
            tails[lo] = x
            tail_idx[lo] = i
    out, at = [], tail_idx[-1] if tail_idx else -1
    while at >= 0:
        out.append(v[at])
        at = prev[at]
    return out[::-1]


def test_longest_increasing_against_plain_patience(tmp_path):
    """The galloping / sampled two-level / prefetch-guess search of longest_increasing() picks exactly the elements plain
    bisection picks: random permutations, near-sorted chains with 1..30 % strays (the k-mer match pattern), descending runs,
    sizes on both sides of the pinned-scratch threshold."""
    import random
    src = tmp_path / "lis_wrap.c"
    src.write_text(LIS_WRAP.replace("@CSRC@", os.path.join(ROOT, "miniwfa_b200", "csrc")))
    so = str(tmp_path / "liblis.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src),
                    os.path.join(ROOT, "miniwfa_b200", "csrc", "kalloc.c"), "-o", so], check=True)
    L = ctypes.CDLL(so)
    L.lis_values.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    L.lis_values.restype = ctypes.c_int32
    rng = random.Random(17)

    def check(v):
        n = len(v)
        a = (ctypes.c_uint64 * max(1, n))(*v)
        out = (ctypes.c_uint64 * max(1, n))()
        m = L.lis_values(n, a, out)
        assert out[:m] == _py_lis(v), n

    check([])
    check([5])
    for n in (2, 3, 63, 64, 65, 127, 128, 129, 1000, 5000):
        for _ in range(6):
            check(rng.sample(range(10 * n), n))
        check(list(range(n, 0, -1)))
        check(list(range(1, n + 1)))
    for n, stray in ((20000, 0.01), (20000, 0.1), (20000, 0.3), (70000, 0.1), (150000, 0.1), (150000, 0.02)):
        # matches sorted by target: mostly query = target + drift (a chain), some strays anywhere, as query << 32 | target
        v, drift = [], 0
        for t in range(n):
            if rng.random() < 0.01:
                drift += rng.randint(-3, 3)
            q = rng.randrange(2 * n) if rng.random() < stray else max(0, t + 1000 + drift)
            v.append(q << 32 | t)
        check(v)

