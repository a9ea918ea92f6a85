import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
mw.set_kernel(mw.KERNEL_TILE)
got = mw.wfa_exact(mw.opt_init(flag=1), b"ACGT", b"ACCT")
print("got", got, flush=True)
