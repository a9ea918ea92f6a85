"""Dev tool (GPU box): a submission whose sequences exceed 2^28 bytes (2304 x 100 kb pairs, 461 MB).  The register-resident steps
address a batch's sequences by 32-bit bit positions, so mwf_wfa_exact_batch() cuts such a submission into parts of at most 200 MB
per device; the rate must stay that of the 128-pair batch.  The first 1024 results are compared with the reference's list
(tests/golden/golden_large.json: config3)."""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from miniwfa_b200 import synth
n = int(os.environ.get("NP", "2304"))
with ThreadPoolExecutor(16) as ex:
    pairs = list(ex.map(lambda i: synth.make_pair(100000, 0.05, i), range(n)))
gold = [c for c in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_large.json")))["cases"] if c["name"] == "config3"][0]["expect"]["s_n_iter"]
mw.wfa_exact_batch(mw.opt_init(), pairs[:256])  # warm-up: context, workspace cache
t0 = time.perf_counter()
r = mw.wfa_exact_batch(mw.opt_init(), pairs)
dt = time.perf_counter() - t0
ok = [[x[0], x[2]] for x in r[:1024]] == gold[:min(n, 1024)]
ns = sum(max(len(t), len(q)) * x[0] for (t, q), x in zip(pairs, r))
print("%d pairs through mwf_wfa_exact_batch(): %.3f s end to end, %.3e n*s cells/s, first %d results %s the reference's" % (n, dt, ns / dt, min(n, 1024), "equal" if ok else "DIFFER FROM"))
sys.exit(0 if ok else 1)
