"""Runs here (no GPU): per-kernel counts of the Blackwell-era mnemonics in libminiwfa_b200.so, as the table of profiles/r2_sass.md."""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "miniwfa_b200", "libminiwfa_b200.so")], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "VIMNMX ", "VIMNMX3", "VIADDMNMX", "VIMNMX.RELU", "LDG.E.64.CONSTANT", "SHFL", "FENCE.VIEW.ASYNC", "MEMBAR", "CCTL.IVALL", "STL", "LDL"]
print("| kernel | instructions | " + " | ".join(c.strip() for c in cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for blk in txt.split("Function : ")[1:]:
    name = blk.split("\n", 1)[0].strip()
    if name.startswith("_ZN3cub"): continue  # (CUB radix sort of the k-mer front end: library code)
    lines = [l for l in blk.split("\n") if re.search(r"/\*[0-9a-f]{4,5}\*/", l)]
    cnt = []
    for c in cols:
        if c == "VIMNMX ": cnt.append(sum(1 for l in lines if re.search(r"\bVIMNMX(\.U32)? ", l)))
        else: cnt.append(sum(1 for l in lines if c in l))
    print("| `%s` | %d | %s |" % (name, len(lines), " | ".join(str(x) for x in cnt)))
