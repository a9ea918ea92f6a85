import os, sys, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import miniwfa_b200 as mw
from conftest import *  # noqa
import conftest
gold = json.load(open("/root/repo/tests/golden/golden.json")) if os.path.exists("/root/repo/tests/golden/golden.json") else None
from test_gpu_parity import exact_cases, case_inputs, expect, got
mw.set_kernel(mw.KERNEL_TILE)
bad = 0
for c in exact_cases(gold["cases"] if isinstance(gold, dict) and "cases" in gold else gold):
    t, q = case_inputs(c)
    if len(t) + len(q) == 0 and c["opt"].get("flag", 0) & 1: continue
    r = mw.wfa_exact(mw.opt_init(**c["opt"]), t, q)
    if got(r) != expect(c):
        os.environ["MWF_B200_TILE_FASTEDGE"] = "0"
        r0 = mw.wfa_exact(mw.opt_init(**c["opt"]), t, q)
        os.environ.pop("MWF_B200_TILE_FASTEDGE")
        e = expect(c)
        print("BAD", c["name"], c["opt"], len(t), len(q), "exp", e[:3], "got", got(r)[:3], "slow-path", got(r0)[:3], flush=True)
        bad += 1
        if bad > 12: break
print("bad", bad)
