"""Dev tool: key metrics + per-region instruction/stall breakdown from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for i, k in enumerate(h):
    if k in want or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.15):
        print("%-90s %s %s" % (k, v[i], rows[1][i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    iA, iS, iE, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][iA], 16)
    tot = sum(int(r[iE]) for r in data)
    mx = sorted(int(r[iE]) for r in data)[-20]
    print("total warp inst", tot, "hot per-line count", mx, "=> inst per warp-step ~ %.1f" % (tot / mx))
    if sys.argv[2] == "dump":
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        for r in data:
            a = int(r[iA], 16) - base
            if lo <= a < hi:
                print(hex(a), r[iS].strip()[:100], r[iE], r[iSm])
