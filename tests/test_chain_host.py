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
