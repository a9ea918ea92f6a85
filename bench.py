#!/usr/bin/env python
"""bench.py -- aligned cells/s (sum n*s) of the exact WFA hot path on N B200s, next to the reference CPU path.

Workload (BASELINE.json config 3, the one the 1/2/4/8-GPU metric is quoted on): a synthetic batch of 100 kb pairs at
~5 % divergence, score-only, default penalties.  One "step" = one pass of the hot path over this rank's pairs.

  default (weak scaling)     : 128 pairs per GPU at every N (1024 pairs at 8 GPUs, the literal config); the line also
                               carries `strong` = the literal 1024-pair batch split over the N ranks (1024 pairs on one GPU
                               at N = 1), so that a 1/2/4/8 run gives both curves
  --scaling strong           : the main line itself is the 1024-pair batch split over the N ranks

  value : sum over all ranks of n*s (n = max(tl,ql)) / device time, sequences already resident in HBM
  e2e   : the same through the C-ABI call a user makes (mwf_wfa_exact_batch semantics: create, stage host buffers
          through pinned memory, H2D, kernels, D2H of the results), host wall clock around the call
  roofline     : wavefront cells (r.n_iter) x 64 algorithmic bytes / kernel time against the measured HBM copy peak, plus --
                 measured on this build inside this run by a short ncu pass over the same workload (rank 0, N = 1) --
                 the DRAM bytes (`traffic`) and the warp instructions of a pass, hence an issue-slot roofline that binds
  cpu_baseline : the unmodified reference (oracle/_ref) or the oracle port timed on this box's host cores
  single_pair  : BASELINE config 2 (one 150 kb pair, CIGAR, high-memory): parity against the reference's golden result,
                 the reference timed beside it on one host core, e2e from host buffers
  large_pairs  : BASELINE configs 4 and 5 (5 Mb pairs): same three things (the CPU leg on a stated down-scaled sample)
  config5_multi: N > 1 only: one 5 Mb / 3 % pair per GPU (BASELINE config 5 at 8 GPUs)

`--impl reference` times the reference's own CPU implementation on all host threads (rank 0 only).
"""
import argparse
import ctypes
import hashlib
import json
import os
import statistics
import struct
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "aligned cells/s (sum n*s)"
UNIT = "cells/s"
BYTES_PER_CELL_SCORE = 64  # 28 B read + 20 B written by wf_next, 16 B first probe of wf_extend
BYTES_PER_CELL_TB = 65
STRONG_PAIRS = 1024        # BASELINE config 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pairs-per-gpu", type=int, default=128)
    ap.add_argument("--len", type=int, default=100000)
    ap.add_argument("--div", type=float, default=0.05)
    ap.add_argument("--no-cpu", action="store_true", help="skip every CPU leg")
    ap.add_argument("--no-single", action="store_true", help="skip the single 150 kb pair leg")
    ap.add_argument("--no-large", action="store_true", help="skip the 5 Mb pair legs (BASELINE configs 4 and 5)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-measurement of the default run")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu pass (roofline.traffic / issue slots fall back to profiles/)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--ncu-child", action="store_true", help=argparse.SUPPRESS)  # one pass of the workload, run under ncu by the parent
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.lines = index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append(line)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def sha1_words(words):
    return hashlib.sha1(struct.pack("<%dI" % len(words), *words)).hexdigest()


def golden_large():
    """Results of the unmodified reference at the BASELINE sizes (tests/golden/make_golden_large.py), by case name."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "golden_large.json")) as f:
            return {c["name"]: c for c in json.load(f)["cases"]}
    except Exception:
        return {}


def make_pairs_parallel(n_pairs, n, p, first):
    from concurrent.futures import ThreadPoolExecutor
    from miniwfa_b200 import synth
    with ThreadPoolExecutor(min(8, os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda i: synth.make_pair(n, p, first + i), range(n_pairs)))


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------------

def cpu_checker():
    """(name, function(opt, t, q) -> (s, n_cigar, n_iter, [cigar words])): the unmodified reference when oracle/_ref travelled
    with the snapshot, else the oracle port."""
    from oracle import orc
    if orc.reference() is not None:
        return "reference", orc.reference_exact
    return "port", orc.oracle_exact


def cpu_build_note():
    p = os.path.join(ROOT, "oracle", "_ref", "BUILD_FLAGS")
    try:
        return open(p).read().strip()
    except Exception:
        return "oracle/Makefile (gcc -O3; see its CFLAGS)"


def cpu_run(pairs, n_threads):
    """Align `pairs` score-only on n_threads host threads.  Returns (seconds, [(s, n_iter)], kind)."""
    from oracle import orc
    kind, fn = cpu_checker()
    opt = orc.make_opt()
    out = [None] * len(pairs)
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(pairs):
                return
            r = fn(opt, pairs[i][0], pairs[i][1])  # ctypes releases the GIL for the duration of the call
            out[i] = (r[0], r[2])

    ths = [threading.Thread(target=work) for _ in range(n_threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, out, kind


def cpu_one(t, q, **kw):
    """One pair on one host core: (seconds, result, kind)."""
    from oracle import orc
    kind, fn = cpu_checker()
    t0 = time.perf_counter()
    r = fn(orc.make_opt(**kw), t, q)
    return time.perf_counter() - t0, r, kind


def cpu_threads(args):
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    if args.cpu_threads > 0:
        n = args.cpu_threads
    return max(1, min(n, 64))


def run_reference_arm(args, rank, world):
    """--impl reference: the reference CPU implementation of the path, all host threads, rank 0 only."""
    if rank != 0:
        return
    from miniwfa_b200 import synth
    nt = cpu_threads(args)
    P = pairs_per_rank(args, world)
    n_sample = min(nt, P)  # one pair per thread per step: a few seconds of wall time per step
    pairs = synth.make_batch(n_sample, args.len, args.div, 0)
    times, kind, res = [], "port", None
    for it in range(args.warmup + args.steps):
        if it < args.warmup and it > 0:
            continue  # one warm-up pass is enough for a CPU loop; the rest would only burn minutes
        dt, res, kind = cpu_run(pairs, nt)
        if it >= args.warmup:
            times.append(dt)
    ns = sum(max(len(t), len(q)) * r[0] for (t, q), r in zip(pairs, res))
    ni = sum(r[1] for r in res)
    dt = sum(times) / len(times)
    val = ns / dt
    sample = "%d of the %d pairs/GPU (%d bp, p=%g), one per host thread" % (n_sample, P, args.len, args.div)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, world), "wavefront_cells_per_s": ni / dt,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nt, "kind": kind, "sample": sample, "build": cpu_build_note()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def pairs_per_rank(args, world):
    return STRONG_PAIRS // world if args.scaling == "strong" else args.pairs_per_gpu


def workload_config(args, world):
    P = pairs_per_rank(args, world)
    return {"workload": "BASELINE config 3: synthetic batch of %d x %d bp pairs, ~%g%% divergence, score-only, "
                        "default penalties x=4,o1=4,e1=2,o2=15,e2=1; %d pairs per GPU (%s scaling)"
                        % (P * world, args.len, args.div * 100, P, args.scaling),
            "pairs_per_gpu": P, "pairs_total": P * world, "seq_len": args.len,
            "divergence": args.div, "mode": "score-only", "parallelism": "pairs sharded across GPUs, no data-path collective",
            "l2": "no flush: the wavefront state a pass streams through (2 x 27 rows x ~50k diagonals x 4 B x 128 pairs ~ 1.4 GB per "
                  "time block, 5.5 GB allocated per 128 pairs) is far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------------
# ncu pass over the same workload (rank 0, N = 1): DRAM bytes and warp instructions of one pass of THIS build
# ------------------------------------------------------------------------------------------------------------------

NCU_METRICS = "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum"


def ncu_child(args):
    """One pass of this rank's workload (what the parent times), to be run under ncu."""
    import miniwfa_b200 as mw
    from miniwfa_b200 import synth
    pairs = synth.make_batch(args.pairs_per_gpu, args.len, args.div, 0)
    with mw.Batch(mw.opt_init(), pairs) as b:
        b.upload()
        b.run()
        b.wait()


def ncu_pass(args):
    """Run `bench.py --ncu-child` under ncu and sum its launch list per kernel.  Returns a dict or None."""
    import csv
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    with tempfile.TemporaryDirectory() as td:
        log = os.path.join(td, "launches.csv")
        cmd = [ncu, "--metrics", NCU_METRICS, "--clock-control", "none", "--csv", "--log-file", log,
               sys.executable, os.path.abspath(__file__), "--ncu-child", "--pairs-per-gpu", str(args.pairs_per_gpu),
               "--len", str(args.len), "--div", str(args.div)]
        try:
            t0 = time.perf_counter()
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
            dt = time.perf_counter() - t0
            if r.returncode != 0 or not os.path.exists(log):
                return None
            rows = [x for x in csv.reader(open(log, errors="replace")) if len(x) > 10]
        except Exception:
            return None
        keep = os.environ.get("MWF_BENCH_KEEP_NCU")
        if keep:
            shutil.copy(log, keep)
    hdr = rows[0]
    iK, iM, iU, iV, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
             "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    agg, ids = {}, {}
    for x in rows[1:]:
        name = x[iK].split("(")[0].replace("void ", "")
        try:
            v = float(x[iV].replace(",", "")) * scale.get(x[iU], 1)
        except ValueError:
            continue
        agg.setdefault(name, {}).setdefault(x[iM], 0.0)
        agg[name][x[iM]] += v
        ids.setdefault(name, set()).add(x[iID])
    kernels = {}
    for k, m in agg.items():
        kernels[k] = {"launches": len(ids[k]), "time_ms": m.get("gpu__time_duration.sum", 0.0),
                      "dram_bytes": m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0),
                      "warp_inst": m.get("smsp__inst_executed.sum", 0.0)}
    tot = sum(v["time_ms"] for v in kernels.values()) or 1.0
    for v in kernels.values():
        v["time_share"] = v["time_ms"] / tot
    return {"kernels": kernels, "seconds": dt, "command": "ncu --metrics %s --clock-control none --csv python bench.py --ncu-child" % NCU_METRICS}


# ------------------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------------------

def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.ncu_child:
        ncu_child(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if local_rank == 0:
        ge.build()
    import miniwfa_b200 as mw
    from miniwfa_b200 import synth, dist as mdist

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    mw.lib()
    mw.set_device(local_rank)
    gold = golden_large()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def all_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    opt = mw.opt_init()
    stream = torch.cuda.Stream()
    fam_names = {mw.KERNEL_CTA: "cta", mw.KERNEL_GRID: "grid", mw.KERNEL_TILE: "tile"}
    c3 = gold.get("config3") if (args.len, args.div) == (100000, 0.05) else None

    def check_config3(first, res):
        """(s, n_iter) of pairs first.. against the unmodified reference's list (tests/golden/golden_large.json)."""
        if c3 is None:
            return "not pinned (non-default workload or no golden file)"
        want = c3["expect"]["s_n_iter"][first:first + len(res)]
        assert [[r[0], r[2]] for r in res] == want, "GPU result differs from the reference's golden (s, n_iter) list"
        return "reference: (s, n_iter) of all %d pairs of this rank equal to golden_large.json:config3" % len(res)

    def measure_batch(pairs, first, steps, warmup, with_clocks):
        """value leg (device-resident, CUDA events on the launching stream) and e2e leg (host buffers) of one batch."""
        P = len(pairs)
        sampler = ClockSampler(local_rank)
        b = mw.Batch(opt, pairs)
        b.set_stream(stream.cuda_stream)
        b.upload()
        for _ in range(warmup):
            b.run()
            b.wait()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if with_clocks:
            sampler.start()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            b.run()
        ev1.record(stream)
        b.wait()
        torch.cuda.synchronize()
        dev_ms = ev0.elapsed_time(ev1)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        clocks = sampler.stop() if with_clocks else None
        b_launches = int(b.launches)
        res = b.fetch()
        kernel_used = fam_names.get(b.kernel_used, "?")
        b.run()  # per-launch kernel duration, measured live by the engine's own CUDA events around the kernels of one pass
        b.wait()
        kernel_ms = b.kernel_ms
        b.close()
        assert all(r[0] > 0 for r in res)
        parity = check_config3(first, res)
        ns_local = float(sum(max(len(t), len(q)) * r[0] for (t, q), r in zip(pairs, res)))
        ni_local = float(sum(r[2] for r in res))
        ns_total, ni_total = sum_over_ranks(ns_local), sum_over_ranks(ni_local)
        ms_per_step = max_over_ranks(dev_ms / steps)
        per_rank_ms = all_ranks(dev_ms / steps)

        host_bufs = mw.api.host_arrays(pairs)  # the caller's host buffers (plain C arrays of pointers and lengths), built once

        def e2e_step():
            with mw.Batch(opt, pairs, arrays=host_bufs) as bb:   # exactly what mwf_wfa_exact_batch() does, kept open to read its byte counters
                bb.upload()
                bb.run()
                rr = bb.fetch()
                if world > 1:  # the one collective of the path: mwf_rst_t records of every shard -> rank 0
                    mdist.gather_to_root(list(range(first, first + P)), rr, P * world, fixed=True)  # score-only: one gather of equal-sized records
                return rr, bb.h2d_bytes, bb.d2h_bytes

        for _ in range(min(warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rr, h2d, d2h = e2e_step()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / steps)
        barrier()
        assert [(r[0], r[2]) for r in rr] == [(r[0], r[2]) for r in res]
        return {"res": res, "value": ns_total / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "per_rank_ms": per_rank_ms,
                "wall_ms_per_step": wall_ms / steps, "clocks": clocks, "ns_total": ns_total, "ni_total": ni_total, "ni_local": ni_local,
                "launches_per_pass": b_launches, "gpu_launches": int(sum_over_ranks(float(b_launches * steps))),
                "kernel_used": kernel_used, "kernel_ms": kernel_ms, "parity": parity,
                "e2e": {"value": ns_total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "api": "mwf_b200_batch_create/upload/run/fetch/destroy (= mwf_wfa_exact_batch) with host buffers"}}

    # ---- the main line ------------------------------------------------------------------------------------------------
    P = pairs_per_rank(args, world)
    first = rank * P
    pairs = make_pairs_parallel(P, args.len, args.div, first)
    M = measure_batch(pairs, first, args.steps, args.warmup, True)
    res = M["res"]

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------
    peak, peak_src = peaks()
    kernel_ms = M["kernel_ms"]
    achieved = M["ni_local"] * BYTES_PER_CELL_SCORE / (kernel_ms * 1e-3) / 1e9
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    sm_mhz = (M["clocks"] or {}).get("sm_mhz") or (M["clocks"] or {}).get("sm_max_mhz") or 1965.0
    traffic, issue, ncu_info = None, None, None
    if rank == 0 and world == 1 and args.scaling == "weak" and not args.no_ncu:
        ncu_info = ncu_pass(args)
    if ncu_info is not None:
        tk = {k: v for k, v in ncu_info["kernels"].items() if "wfa_tile_persist_kernel" in k or "wfa_tile_kernel" in k or "wfa_cta_kernel" in k or "wfa_grid_kernel" in k}
        allk = ncu_info["kernels"]
        traffic = sum(v["dram_bytes"] for k, v in allk.items() if "pack" not in k)
        winst = sum(v["warp_inst"] for k, v in allk.items() if "pack" not in k)
        slots = n_sm * 4 * sm_mhz * 1e6  # one warp instruction per scheduler per cycle, 4 schedulers per SM
        issue = {"warp_inst_per_pass": winst, "thread_inst_per_cell": winst * 32 / M["ni_local"],
                 "achieved": winst / (kernel_ms * 1e-3), "peak": slots, "unit": "warp-inst/s", "frac": winst / (kernel_ms * 1e-3) / slots,
                 "peak_source": "%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % (n_sm, sm_mhz),
                 "dominant_kernel_time_share_under_ncu": max((v["time_share"] for v in tk.values()), default=None),
                 "source": "ncu pass inside this run: " + ncu_info["command"]}
    else:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("bench_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
    roofline = {"bound": "hbm", "kernel": "wfa_tile_persist_kernel<0, 4>" if M["kernel_used"] == "tile" else "wfa_%s_kernel" % M["kernel_used"], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": ("ncu pass inside this run (dram__bytes_read.sum + dram__bytes_write.sum over every launch of one pass)"
                                   if ncu_info is not None else "profiles/traffic.json (committed ncu launch list; no ncu pass in this run)"),
                "traffic_GBps": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                "traffic_frac_of_peak": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None,
                "issue_slot": issue,
                "peak_source": peak_src, "algorithmic_bytes_per_cell": BYTES_PER_CELL_SCORE, "cells_per_pass": M["ni_local"],
                "launches_per_pass": M["launches_per_pass"],
                "kernel_ms": kernel_ms, "kernel_ms_max_over_ranks": max_over_ranks(kernel_ms),
                "note": "one pass over the batch = launches_per_pass launches (score-0 init, the queue's first items, ONE persistent tile kernel "
                        "that also runs the planner between blocks of 32 scores); kernel_ms is the CUDA-event time over all of them on the launching stream. frac > 1 on the "
                        "algorithmic bytes is expected and is not a skipped-work artefact: the 64 B/cell are what the reference's "
                        "formulation streams per cell (SURVEY 8d), while the tile engine keeps the live ring rows in shared memory for "
                        "32 scores, so its measured DRAM traffic (`traffic`, bytes per pass) is ~6 B/cell; every cell is computed "
                        "(r.n_iter and r.s equal the CPU reference's). The roofline that binds this kernel is instruction issue: "
                        "`issue_slot` (warp instructions of a pass / kernel time against SMs x 4 schedulers x clock)"}

    line = {"metric": METRIC, "value": M["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": M["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
            "wavefront_cells_per_s": M["ni_total"] / (M["ms_per_step"] * 1e-3), "wall_ms_per_step": M["wall_ms_per_step"],
            "ms_per_step_per_rank": M["per_rank_ms"], "parity": M["parity"],
            "clocks": M["clocks"], "e2e": M["e2e"], "gpu_launches": M["gpu_launches"], "roofline": roofline}

    # ---- strong scaling: the literal 1024-pair batch split over the N ranks ------------------------------------------------
    if args.scaling == "weak" and not args.no_strong and (args.len, args.div, args.pairs_per_gpu) == (100000, 0.05, 128):
        Ps = STRONG_PAIRS // world
        if Ps == P:
            S = M
        else:
            spairs = make_pairs_parallel(Ps, args.len, args.div, rank * Ps)
            S = measure_batch(spairs, rank * Ps, 2, 1, False)
            del spairs
        line["strong"] = {"workload": "BASELINE config 3, the whole batch: %d x 100 kb pairs split over %d GPU(s), %d pairs per GPU"
                                      % (STRONG_PAIRS, world, Ps),
                          "scaling": "strong", "pairs_total": STRONG_PAIRS, "pairs_per_gpu": Ps, "value": S["value"], "unit": UNIT,
                          "ms_per_step": S["ms_per_step"], "ms_per_step_per_rank": S["per_rank_ms"],
                          "steps": args.steps if S is M else 2, "warmup": args.warmup if S is M else 1,
                          "e2e": S["e2e"], "parity": S["parity"]}

    # ---- single 150 kb pair (config 2), rank 0 -------------------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_single:  # (the single-pair legs belong to the N = 1 line; at N > 1 the other ranks would only wait)
        c2 = gold.get("config2-c")
        t, q = synth.make_pair(150000, 0.038, 900000)
        o2 = mw.opt_init(flag=mw.F_CIGAR)
        with mw.Batch(o2, [(t, q)]) as sb:
            sb.upload()
            for _ in range(2):
                sb.run()
                sb.wait()
            r1 = sb.fetch()[0]
            kms = sb.kernel_ms
            fam = fam_names.get(sb.kernel_used, "?")
            launches1 = int(sb.launches)
        e2e_t = []
        for _ in range(3):  # the call a user makes: host buffers in, mwf_rst_t with the CIGAR out (H2D, kernels, traceback, D2H inside)
            t0 = time.perf_counter()
            r2 = mw.wfa_exact(o2, t, q)
            e2e_t.append(time.perf_counter() - t0)
        assert r2 == r1 and mw.cigar2score(o2, r1[3]) == (r1[0], len(t), len(q))
        n1 = max(len(t), len(q))
        sp = {"workload": "BASELINE config 2 surrogate: one synthetic 150 kb pair (p=0.038), CIGAR, high-memory (test-mwf -c)",
              "s": r1[0], "n_iter": r1[2], "n_cigar": r1[1], "kernel": fam, "kernel_ms": kms, "gpu_launches": launches1,
              "value": n1 * r1[0] / (kms * 1e-3), "unit": UNIT,
              "e2e": {"value": n1 * r1[0] / min(e2e_t), "unit": UNIT, "seconds": min(e2e_t), "seconds_all": e2e_t,
                      "h2d_bytes_per_step": len(t) + len(q), "d2h_bytes_per_step": 4 * r1[1] + 48,
                      "api": "mwf_wfa_exact() with host buffers (create, H2D, kernels, traceback, D2H, CIGAR into the caller's km)"},
              "roofline_frac": r1[2] * BYTES_PER_CELL_TB / (kms * 1e-3) / 1e9 / peak,
              "note": "one pair = one dependency chain of s scores: latency-bound, not HBM-bound"}
        # latency of one tiny call (there is no CPU path, so a t3-sized pair pays create / upload / run / fetch).
        # Taken right after the GPU-bound calls above: after seconds of an idle GPU (the CPU leg below) the same call was measured
        # at 0.8-1.3 ms until the device had been busy again (tools/tiny_latency.py).
        tt = []
        for _ in range(50):
            t0 = time.perf_counter()
            rt = mw.wfa_exact(o2, b"ACGTACGTACGTTTGACA" * 4, b"ACGTACGAACGTTTGACA" * 4)
            tt.append(time.perf_counter() - t0)
        tt.sort()
        line["tiny_pair_latency"] = {"workload": "mwf_wfa_exact() on a 72 bp pair with CIGAR, host buffers, 50 calls after a busy GPU", "s": rt[0],
                                     "median_us": statistics.median(tt) * 1e6, "min_us": tt[0] * 1e6, "p90_us": tt[44] * 1e6}
        if c2 is not None:
            e = c2["expect"]
            assert (r1[0], r1[1], r1[2], sha1_words(r1[3])) == (e["s"], e["n_cigar"], e["n_iter"], e["cigar_sha1"]), "config 2 differs from the reference"
            sp["parity"] = "reference: s, n_iter, n_cigar and sha1 of the CIGAR words equal to golden_large.json:config2-c"
        if not args.no_cpu:
            dt, rc, kind = cpu_one(t, q, flag=1)
            assert rc == r1, "config 2: GPU result differs from the CPU checker run beside it"
            sp["cpu_baseline"] = {"value": n1 * rc[0] / dt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": dt,
                                  "sample": "the whole pair, one host thread; s, n_iter and all CIGAR words equal to the GPU's",
                                  "build": cpu_build_note()}
            sp["parity"] = "reference: every CIGAR word, s, n_iter equal to the %s run beside it%s" % (kind, " and to golden_large.json:config2-c" if c2 else "")
        line["single_pair"] = sp

    # ---- 5 Mb pairs (config 4 and config 5 surrogates, one pair each), rank 0 --------------------------------------------
    if rank == 0 and world == 1 and not args.no_large and not args.no_single:
        large = {}
        for name, p, kw, gname, cpu_spec, what in (
                ("config4", 0.0097, {"flag": mw.F_CIGAR, "step": 5000}, "config4-cp5000", (1000000, 0.0097, 424242, {"flag": 1, "step": 5000}),
                 "BASELINE config 4 surrogate: one synthetic 5 Mb pair (p=0.0097, s ~ 231 k), low-memory mode (test-mwf -cp5000)"),
                ("config5", 0.03, {"flag": mw.F_CIGAR}, "config5-cp5000", (400000, 0.03, 424242, {"flag": 1}),
                 "BASELINE config 5 surrogate: one synthetic 5 Mb pair (p=0.03, s ~ 713 k, 5e11 cells), high-memory CIGAR (test-mwf -c); "
                 "its 505 GB of traceback bytes do not fit HBM -- predicted from the shared 13-mer fraction of the pair before any "
                 "alignment work -- so the engine goes straight to the segmented traceback (snapshots + recompute)")):
            t, q = synth.make_pair(5000000, p, 424242)
            oo = mw.opt_init(**kw)
            runs = []
            for _ in range(2):  # the first call also sizes the workspace cache
                t0 = time.perf_counter()
                rl = mw.wfa_exact(oo, t, q)  # host buffers in, CIGAR out: create, H2D, every pass, traceback, D2H
                runs.append(time.perf_counter() - t0)
            dt = min(runs)
            assert rl[0] > 0 and mw.cigar2score(oo, rl[3]) == (rl[0], len(t), len(q))
            nL = max(len(t), len(q))
            ent = {"workload": what, "s": rl[0], "n_iter": rl[2], "n_cigar": rl[1],
                   "e2e": {"value": nL * rl[0] / dt, "unit": UNIT, "seconds": dt, "seconds_all": runs,
                           "h2d_bytes_per_step": len(t) + len(q), "d2h_bytes_per_step": 4 * rl[1] + 48,
                           "api": "mwf_wfa_exact() with host buffers"},
                   "value": nL * rl[0] / dt, "unit": UNIT, "seconds": dt,
                   "check": "CIGAR re-scored with mwf_cigar2score: score == s and it consumes both sequences"}
            g = gold.get(gname)
            if g is not None:
                e = g["expect"]
                got = (rl[0], rl[1], sha1_words(rl[3])) + ((rl[2],) if name == "config4" else ())
                want = (e["s"], e["n_cigar"], e["cigar_sha1"]) + ((e["n_iter"],) if name == "config4" else ())
                assert got == want, "%s differs from the reference" % name
                ent["parity"] = ("reference: s, n_cigar%s and sha1 of the CIGAR words equal to golden_large.json:%s (the unmodified reference, "
                                 "%.0f s of one core of the build container%s)"
                                 % (", n_iter" if name == "config4" else "", gname, g.get("reference_seconds_here", 0),
                                    "; the reference can only run this pair with -cp5000, which gives the same CIGAR (BASELINE.md 3.3)" if name == "config5" else ""))
                ent["reference_seconds_build_container"] = g.get("reference_seconds_here")
            else:
                ent["parity"] = "self-consistency only: golden_large.json has no %s yet" % gname
            if not args.no_cpu:  # bounded CPU sample: the same mode on a down-scaled pair of the same divergence, checked against the GPU
                cn, cp, ci, ckw = cpu_spec
                ct, cq = synth.make_pair(cn, cp, ci)
                cdt, rc, kind = cpu_one(ct, cq, **ckw)
                rg = mw.wfa_exact(mw.opt_init(**ckw), ct, cq)
                assert rg == rc, "%s CPU sample: GPU result differs from the CPU checker" % name
                cN = max(len(ct), len(cq))
                t0 = time.perf_counter()
                mw.wfa_exact(mw.opt_init(**ckw), ct, cq)
                gdt = time.perf_counter() - t0
                ent["cpu_baseline"] = {"value": cN * rc[0] / cdt, "unit": UNIT, "cores": 1, "kind": kind, "seconds": cdt,
                                       "sample": "down-scaled pair of the same divergence and mode: %d bp, p=%g, %s (s = %d, n_iter = %d), one host "
                                                 "thread; every CIGAR word, s and n_iter equal to the GPU's; the GPU takes %.3f s end to end on this sample"
                                                 % (cn, cp, "-cp5000" if ckw.get("step") else "-c", rc[0], rc[2], gdt),
                                       "gpu_seconds_same_sample": gdt, "build": cpu_build_note()}
                del ct, cq
            large[name] = ent
            del t, q
        # mwf_wfa_auto on the config-5 pair: exact with a budget of 1e8 cells, then the chaining heuristic (k-mer front end on
        # the device, all gap fills as one batch); the unmodified reference's mwf_wfa_auto on one host core beside it
        t, q = synth.make_pair(5000000, 0.03, 424242)
        oa = mw.opt_init(flag=mw.F_CIGAR)
        k0 = int(mw.lib().mwf_b200_kmer_launches())
        ta = []
        for _ in range(3):
            t0 = time.perf_counter()
            ra = mw.wfa_auto(oa, t, q)
            ta.append(time.perf_counter() - t0)
        assert mw.cigar2score(oa, ra[3])[1:] == (len(t), len(q))
        auto = {"workload": "mwf_wfa_auto (miniwfa.c:898-908) on the config-5 pair, CIGAR: host buffers in, CIGAR out", "s": ra[0],
                "n_cigar": ra[1], "seconds_first_call": ta[0], "seconds": min(ta[1:]),
                "kmer_front_end_launches_per_call": (int(mw.lib().mwf_b200_kmer_launches()) - k0) // 3}
        if not args.no_cpu:
            try:
                from oracle import orc
                ref = orc.reference()
            except Exception:
                ref = None
            if ref is not None:
                ro, rr = orc.make_opt(flag=1), orc.Rst()
                t0 = time.perf_counter()
                ref.mwf_wfa_auto(None, ctypes.byref(ro), len(t), t, len(q), q, ctypes.byref(rr))
                auto["reference_cpu_seconds"] = time.perf_counter() - t0
                auto["check"] = "score and all CIGAR words equal to the reference's"
                assert (rr.s, rr.n_cigar) == (ra[0], ra[1]) and rr.cigar[:rr.n_cigar] == ra[3]
        large["auto"] = auto
        del t, q
        large["reference_published"] = "README.md:98-99 (Xeon 6230, 1 thread): MHC (s = 229 868) 385 s high-memory / 544 s low-memory"
        line["large_pairs"] = large
        mw.lib().mwf_b200_release_cache()

    # ---- config 5 across the GPUs: one 5 Mb / 3 % pair per GPU (N > 1) ----------------------------------------------------
    if world > 1 and not args.no_large:
        t, q = synth.make_pair(5000000, 0.03, 424242 + rank)
        o5 = mw.opt_init(flag=mw.F_CIGAR)
        barrier()
        t0 = time.perf_counter()
        r5 = mw.wfa_exact(o5, t, q)
        mdist.gather_to_root([rank], [r5], world)
        dt5 = time.perf_counter() - t0
        assert r5[0] > 0 and mw.cigar2score(o5, r5[3]) == (r5[0], len(t), len(q))
        g = gold.get("config5-cp5000")
        if rank == 0 and g is not None:
            assert (r5[0], r5[1], sha1_words(r5[3])) == (g["expect"]["s"], g["expect"]["n_cigar"], g["expect"]["cigar_sha1"])
        ns5 = sum_over_ranks(float(max(len(t), len(q)) * r5[0]))
        per_rank = all_ranks(dt5)
        line["config5_multi"] = {"workload": "BASELINE config 5: %d synthetic 5 Mb pairs (p=0.03), high-memory CIGAR, one pair per GPU, "
                                             "mwf_wfa_exact() with host buffers on every rank + the gather of the mwf_rst_t records to rank 0" % world,
                                 "value": ns5 / max(per_rank), "unit": UNIT, "seconds": max(per_rank), "seconds_per_rank": per_rank,
                                 "check": "every rank: CIGAR re-scored with mwf_cigar2score (score == s, both sequences consumed); "
                                          "rank 0's pair: s, n_cigar and CIGAR sha1 equal to the reference's golden" if g is not None else
                                          "every rank: CIGAR re-scored with mwf_cigar2score"}
        del t, q
        mw.lib().mwf_b200_release_cache()

    # ---- cpu_baseline: rank 0 at N=1 only ---------------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        nt = cpu_threads(args)
        n_sample = min(nt, P)
        dt, cres, kind = cpu_run(pairs[:n_sample], nt)
        assert [(r[0], r[2]) for r in res[:n_sample]] == cres, "GPU result differs from the CPU checker"
        cns = sum(max(len(t), len(q)) * r[0] for (t, q), r in zip(pairs[:n_sample], cres))
        line["cpu_baseline"] = {"value": cns / dt, "unit": UNIT, "cores": nt, "kind": kind,
                                "sample": "first %d of the %d pairs, one per host thread, %.1f s wall; s and n_iter equal to the GPU's"
                                          % (n_sample, P, dt),
                                "wavefront_cells_per_s": sum(r[1] for r in cres) / dt, "build": cpu_build_note()}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
