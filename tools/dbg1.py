import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import miniwfa_b200 as mw
from oracle import orc
mw.set_kernel(mw.KERNEL_TILE)
t3 = (b"CAGGGGCAGACTGACACTTCACACGGCCGGGTACTCTAACAGACCTGCAGCTGAGGGTCCT",
      b"TAGGGGCAGACTGACACCTCACACGGCCGGGTACTCCTCTGAGACAAAACTTCCAGAGGAACGATCAGACAGCAGCATTCGCGGTTCATGAAAATCCGCTGTTCTG"
      b"CAGCCACCGCTGCTGGTACCCAGGCAAACAGGGTCTAGAGTGGACCTTTAGCAAACTCCAACAGACCTGCAGCTGAGGGTCCT")
mode = sys.argv[1]
kw = {} if mode == "s" else {"flag": 1}
print("calling", kw, flush=True)
want = orc.checker_exact(orc.make_opt(**kw), *t3)
print("want", want[:3], flush=True)
got = mw.wfa_exact(mw.opt_init(**kw), *t3)
print("got", got[:3], got == want, flush=True)
