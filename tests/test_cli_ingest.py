"""CPU: the ingest pipeline of the test-mwf CLI (csrc/main.c: reader thread, chunks of pairs, FASTA / FASTQ / gzip parsing),
linked against a stub of the alignment entry points so that it runs without a GPU.  The stub scores a pair as 1000 tl + ql,
which makes pairing, order and lengths visible in the output."""
import gzip
import os
import random
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = r"""
#include <string.h>
#include "miniwfa.h"
#include "mwf_b200.h"
void mwf_opt_init(mwf_opt_t *o) { memset(o, 0, sizeof(*o)); o->x = 4, o->o1 = 4, o->e1 = 2, o->o2 = 15, o->e2 = 1; }
void mwf_wfa_exact_batch(void *km, const mwf_opt_t *opt, int32_t n, const int32_t *tl, const char *const *ts,
                         const int32_t *ql, const char *const *qs, mwf_rst_t *r)
{
	int i;
	for (i = 0; i < n; ++i) memset(&r[i], 0, sizeof(r[i])), r[i].s = tl[i] * 1000 + ql[i];
}
void mwf_wfa_chain(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r) { memset(r, 0, sizeof(*r)); r->s = 1; }
void mwf_wfa_auto(void *km, const mwf_opt_t *opt, int32_t tl, const char *ts, int32_t ql, const char *qs, mwf_rst_t *r) { memset(r, 0, sizeof(*r)); r->s = 2; }
void mwf_assert_cigar(const mwf_opt_t *opt, int32_t n_cigar, const uint32_t *cigar, int32_t tl0, int32_t ql0, int32_t s0) {}
"""


def test_cli_reader_thread_and_chunks(tmp_path):
    stub = tmp_path / "stub.c"
    stub.write_text(STUB)
    exe = str(tmp_path / "cli")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "miniwfa_b200", "csrc", "main.c"),
                    str(stub), "-o", exe, "-lz", "-lpthread"], check=True)
    rng = random.Random(3)
    recs1 = [("t%d" % i, "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 300)))) for i in range(57)]
    recs2 = [("q%d" % i, "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 300)))) for i in range(60)]
    fa = ["\n"]  # leading blank line, multi-line records, CRLF, comments after the name, stray blank lines
    for name, s in recs1:
        fa.append(">%s some comment\n" % name)
        fa.extend(s[j:j + 60] + "\r\n" for j in range(0, len(s), 60))
        if rng.random() < 0.2:
            fa.append("\n")
    (tmp_path / "a.fa").write_text("".join(fa))
    with gzip.open(tmp_path / "b.fq.gz", "wt") as f:  # FASTQ whose quality lines start with '@'
        f.write("".join("@%s x\n%s\n+\n%s\n" % (n, s, "@" * len(s)) for n, s in recs2))
    want = ["%s\t%d\t0\t%d\t+\t%s\t%d\t0\t%d\t%d" % (a, len(s), len(s), b, len(t), len(t), len(s) * 1000 + len(t))
            for (a, s), (b, t) in zip(recs1, recs2)]
    for env in ({}, {"MWF_CLI_CHUNK_PAIRS": "1"}, {"MWF_CLI_CHUNK_PAIRS": "7"}, {"MWF_CLI_CHUNK_BASES": "500"}):
        e = dict(os.environ)
        e.update(env)
        out = subprocess.run([exe, str(tmp_path / "a.fa"), str(tmp_path / "b.fq.gz")], capture_output=True, text=True, env=e, timeout=60)
        assert out.returncode == 0, out.stderr
        assert out.stdout.strip().split("\n") == want, env
        assert out.stderr.count("T\t") == len(want)
    out = subprocess.run([exe, "-u", str(tmp_path / "a.fa"), str(tmp_path / "b.fq.gz")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and all(line.endswith("\t1") for line in out.stdout.strip().split("\n"))
