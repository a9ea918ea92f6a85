"""Dev tool: sum gpu__time_duration / dram bytes per kernel name from an `ncu --csv` launch list -> profiles/traffic.json."""
import csv, json, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
iID = hdr.index("ID")
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.defaultdict(set)
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows[1:]:
    name = r[iK].split("(")[0].replace("void ", "")
    v = float(r[iV].replace(",", "")) * scale.get(r[iU], 1)
    agg[name][r[iM]] += v
    cnt[name].add(r[iID])
out = {}
for k, m in agg.items():
    out[k] = {"launches": len(cnt[k]), "time_ms": m.get("gpu__time_duration.sum", 0.0),
              "dram_bytes": m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)}
tot_t = sum(v["time_ms"] for v in out.values())
for v in out.values():
    v["time_share"] = v["time_ms"] / tot_t if tot_t else None
res = {"source": sys.argv[1], "passes": int(sys.argv[3]) if len(sys.argv) > 3 else 1, "kernels": out}
tile = [v for k, v in out.items() if "wfa_tile_kernel" in k]
if tile:
    res["bench_kernel_dram_bytes_per_launch"] = sum(v["dram_bytes"] for v in tile) / res["passes"]
json.dump(res, open(sys.argv[2], "w"), indent=1)
print(json.dumps(res, indent=1))
