"""Dev tool (GPU box): pairs per second of the test-mwf CLI (csrc/main.c: reader thread + batches) on FASTA files of read-sized and
of 100 kb pairs, next to the reference's own CLI (oracle/_ref/test-mwf-ref, one core) on a sample of the same files."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miniwfa_b200 import synth

CLI = os.path.join(ROOT, "miniwfa_b200", "test-mwf")
REF = os.path.join(ROOT, "oracle", "_ref", "test-mwf-ref")


def write_fasta(path, seqs):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">s%d\n" % i + s + b"\n")


def timed(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, r.stdout


with tempfile.TemporaryDirectory() as td:
    for name, n_pairs, n, p, n_ref in (("150 bp reads, 2 %", 200000, 150, 0.02, 20000), ("150 bp reads, 2 %", 2000000, 150, 0.02, 20000), ("100 kb pairs, 5 %", 256, 100000, 0.05, 2)):
        pairs = [synth.make_pair(n, p, 7000 + i) for i in range(n_pairs)] if n > 1000 else None
        if pairs is None:  # read-sized pairs: windows of one long sequence, substitutions by a numpy mask
            import numpy as np
            t, _ = synth.make_pair(4000000, 0.0, 7000)
            ta = np.frombuffer(t, dtype=np.uint8)
            rs = np.random.RandomState(1)
            starts = rs.randint(0, len(t) - n, size=n_pairs)
            T = ta[starts[:, None] + np.arange(n)[None, :]]
            Q = T.copy()
            m = rs.random_sample(Q.shape) < p
            Q[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.randint(0, 4, size=int(m.sum()))]
            pairs = [(T[i].tobytes(), Q[i].tobytes()) for i in range(n_pairs)]
        fa, fb = os.path.join(td, "a.fa"), os.path.join(td, "b.fa")
        write_fasta(fa, [x[0] for x in pairs]); write_fasta(fb, [x[1] for x in pairs])
        for flags in ([], ["-c"]):
            timed([CLI] + flags + [fa, fb])  # warm-up: CUDA context, workspace cache is per process, so this only warms the page cache
            dt, out = timed([CLI] + flags + [fa, fb])
            line = "%s, test-mwf %s: %d pairs in %.3f s = %.0f pairs/s" % (name, " ".join(flags) or "(score)", n_pairs, dt, n_pairs / dt)
            if os.path.exists(REF):
                ra, rb = os.path.join(td, "ra.fa"), os.path.join(td, "rb.fa")
                write_fasta(ra, [x[0] for x in pairs[:n_ref]]); write_fasta(rb, [x[1] for x in pairs[:n_ref]])
                rdt, rout = timed([REF] + flags + [ra, rb])
                same = rout.splitlines() == out.splitlines()[:len(rout.splitlines())]
                line += "; reference CLI on the first %d pairs: %.3f s = %.0f pairs/s, output lines %s" % (n_ref, rdt, n_ref / rdt, "identical" if same else "DIFFER")
            print(line, flush=True)
