#!/usr/bin/env python
"""bench.py -- aligned cells/s (sum n*s) of the exact WFA hot path on N B200s, next to the reference CPU path.

Workload (BASELINE.json config 3, the one the 1/2/4/8-GPU metric is quoted on): a synthetic batch of 100 kb pairs at
~5 % divergence, score-only, default penalties; 1024 pairs at 8 GPUs = 128 pairs per GPU, held fixed per GPU as N
changes (weak scaling).  One "step" = one pass of the hot path over this rank's 128 pairs.

  value : sum over all ranks of n*s (n = max(tl,ql)) / device time, sequences already resident in HBM
  e2e   : the same through the C-ABI call a user makes (mwf_wfa_exact_batch semantics: create, stage host buffers
          through pinned memory, H2D, kernels, D2H of the results), host wall clock around the call
  roofline     : wavefront cells (r.n_iter) x 64 algorithmic bytes / kernel time, against the measured HBM copy peak
                 (the tile engine keeps the ring in shared memory: `traffic` = measured DRAM bytes per pass, from ncu)
  cpu_baseline : the unmodified reference (oracle/_ref) or the oracle port timed on this box's host cores
  single_pair  : BASELINE.json config 2 surrogate (one 150 kb pair, CIGAR, high-memory) on one GPU, for reference

`--impl reference` times the reference's own CPU implementation on all host threads (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "aligned cells/s (sum n*s)"
UNIT = "cells/s"
BYTES_PER_CELL_SCORE = 64  # 28 B read + 20 B written by wf_next, 16 B first probe of wf_extend
BYTES_PER_CELL_TB = 65


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=128)
    ap.add_argument("--len", type=int, default=100000)
    ap.add_argument("--div", type=float, default=0.05)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-single", action="store_true", help="skip the single 150 kb pair leg")
    ap.add_argument("--no-large", action="store_true", help="skip the 5 Mb pair legs (BASELINE configs 4 and 5, one pair each)")
    ap.add_argument("--cpu-threads", type=int, default=0)
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


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------------

def cpu_run(pairs, n_threads):
    """Align `pairs` score-only with the reference (oracle/_ref) or the oracle port on n_threads host threads.
    Returns (seconds, [(s, n_iter)], kind)."""
    from oracle import orc
    kind = "reference" if orc.reference() is not None else "port"
    fn = orc.reference_exact if kind == "reference" else orc.oracle_exact
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
    n_sample = min(nt, args.pairs_per_gpu)  # one pair per thread per step: a few seconds of wall time per step
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
    sample = "%d of the %d pairs/GPU (%d bp, p=%g), one per host thread" % (n_sample, args.pairs_per_gpu, args.len, args.div)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, world), "wavefront_cells_per_s": ni / dt,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nt, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "BASELINE config 3: synthetic batch of %d x %d bp pairs, ~%g%% divergence, score-only, "
                        "default penalties x=4,o1=4,e1=2,o2=15,e2=1; %d pairs per GPU"
                        % (args.pairs_per_gpu * world, args.len, args.div * 100, args.pairs_per_gpu),
            "pairs_per_gpu": args.pairs_per_gpu, "pairs_total": args.pairs_per_gpu * world, "seq_len": args.len,
            "divergence": args.div, "mode": "score-only", "parallelism": "pairs sharded across GPUs, no data-path collective",
            "l2": "no flush: the wavefront state a pass streams through (2 x 27 rows x ~50k diagonals x 4 B x 128 pairs ~ 1.4 GB per "
                  "time block, 5.5 GB allocated) is far larger than the 126 MB L2"}


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

    P = args.pairs_per_gpu
    pairs = synth.make_batch(P, args.len, args.div, rank * P)
    opt = mw.opt_init()
    stream = torch.cuda.Stream()
    sampler = ClockSampler(local_rank)

    # ---- value: device-resident inputs, CUDA events on the launching stream --------------------------------------
    b = mw.Batch(opt, pairs)
    b.set_stream(stream.cuda_stream)
    b.upload()
    for _ in range(args.warmup):
        b.run()
        b.wait()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    kernel_ms = []
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        b.run()
        launches += 1  # counted again from the engine below
    ev1.record(stream)
    b.wait()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    b_launches = int(b.launches)
    launches = int(sum_over_ranks(float(b_launches * args.steps)))  # kernels launched inside the timed region, all ranks
    res = b.fetch()
    fam_names = {mw.KERNEL_CTA: "cta", mw.KERNEL_GRID: "grid", mw.KERNEL_TILE: "tile"}
    kernel_used = fam_names.get(b.kernel_used, "?")
    # per-launch kernel duration, measured live by the engine's own CUDA events around the kernel of the last pass
    b.run()
    b.wait()
    kernel_ms.append(b.kernel_ms)
    b.close()

    ns_local = float(sum(max(len(t), len(q)) * r[0] for (t, q), r in zip(pairs, res)))
    ni_local = float(sum(r[2] for r in res))
    assert all(r[0] > 0 for r in res)
    ns_total, ni_total = sum_over_ranks(ns_local), sum_over_ranks(ni_local)
    ms_per_step = max_over_ranks(dev_ms / args.steps)
    value = ns_total / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the public C-ABI call, copies inside the timed region -------------------------
    host_bufs = mw.api.host_arrays(pairs)  # the caller's host buffers (plain C arrays of pointers and lengths), built once

    def e2e_step():
        with mw.Batch(opt, pairs, arrays=host_bufs) as bb:   # exactly what mwf_wfa_exact_batch() does, kept open to read its byte counters
            bb.upload()
            bb.run()
            rr = bb.fetch()
            if world > 1:  # the one collective of the path: mwf_rst_t records of every shard -> rank 0
                mdist.gather_to_root(list(range(rank * P, rank * P + P)), rr, P * world)
            return rr, bb.h2d_bytes, bb.d2h_bytes

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rr, h2d, d2h = e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    assert [(r[0], r[2]) for r in rr] == [(r[0], r[2]) for r in res]
    e2e = {"value": ns_total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "api": "mwf_b200_batch_create/upload/run/fetch/destroy (= mwf_wfa_exact_batch) with host buffers"}

    # ---- roofline of the dominant (only) kernel --------------------------------------------------------------------
    peak, peak_src = peaks()
    k_ms = max_over_ranks(kernel_ms[0])
    achieved = ni_local * BYTES_PER_CELL_SCORE / (kernel_ms[0] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("bench_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    per_pass = int(b_launches)
    roofline = {"bound": "hbm", "kernel": "wfa_%s_kernel" % kernel_used, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_cell": BYTES_PER_CELL_SCORE, "cells_per_pass": ni_local,
                "launches_per_pass": per_pass,
                "kernel_ms": kernel_ms[0], "kernel_ms_max_over_ranks": k_ms,
                "note": "one pass over the batch = launches_per_pass launches (score-0 init, then one plan + one tile kernel per block "
                        "of 32 scores); kernel_ms is the CUDA-event time over all of them on the launching stream, so achieved = "
                        "cells_per_pass x 64 B / kernel_ms understates the tile kernel alone. frac > 1 is expected here and is not a "
                        "skipped-work artefact: the 64 B/cell are what the reference's formulation streams per cell (SURVEY 8d), while "
                        "the tile engine keeps the 27 live ring rows in shared memory for 32 scores, so its measured DRAM traffic "
                        "(`traffic`, bytes per pass, ncu launch list in profiles/) is ~6 B/cell; every cell is computed (r.n_iter and "
                        "r.s equal the CPU reference's). What limits the kernel now is instruction issue (68 % of peak, 80 "
                        "instructions per cell) and the L1 data pipe (68 %): profiles/r1_tile_kernel.md"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": workload_config(args, world),
            "wavefront_cells_per_s": ni_total / (ms_per_step * 1e-3), "wall_ms_per_step": wall_ms / args.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline}

    # ---- single 150 kb pair (config 2 surrogate), rank 0 -----------------------------------------------------------
    if rank == 0 and not args.no_single:
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
        t0 = time.perf_counter()
        r2 = mw.wfa_exact(o2, t, q)
        e2e1 = time.perf_counter() - t0
        assert r2 == r1 and mw.cigar2score(o2, r1[3]) == (r1[0], len(t), len(q))
        n1 = max(len(t), len(q))
        line["single_pair"] = {"workload": "BASELINE config 2 surrogate: one synthetic 150 kb pair (p=0.038), CIGAR, high-memory",
                               "s": r1[0], "n_iter": r1[2], "n_cigar": r1[1], "kernel": fam, "kernel_ms": kms, "gpu_launches": launches1,
                               "value": n1 * r1[0] / (kms * 1e-3), "e2e_value": n1 * r1[0] / e2e1, "unit": UNIT,
                               "roofline_frac": r1[2] * BYTES_PER_CELL_TB / (kms * 1e-3) / 1e9 / peak,
                               "note": "one pair = one dependency chain of s scores; the tile engine cuts it into blocks of 64 scores "
                                       "x (width/384) tiles of 512 threads, at most ~140 tiles: latency-bound, not HBM-bound"}

    # ---- 5 Mb pairs (config 4 and config 5 surrogates, one pair each), rank 0 --------------------------------------------
    if rank == 0 and not args.no_large and not args.no_single:
        large = {}
        for name, p, kw, what in (("config4", 0.0097, {"flag": mw.F_CIGAR, "step": 5000},
                                   "BASELINE config 4 surrogate: one synthetic 5 Mb pair (p=0.0097, s ~ 231 k), low-memory mode -cp5000"),
                                  ("config5", 0.03, {"flag": mw.F_CIGAR},
                                   "BASELINE config 5 surrogate: one synthetic 5 Mb pair (p=0.03, s ~ 711 k, 5e11 cells), high-memory "
                                   "CIGAR; its 505 GB of traceback bytes do not fit HBM -- predicted from the shared 13-mer fraction of the pair before "
                                   "any alignment work -- so the engine goes straight to the segmented traceback (snapshots + recompute)")):
            t, q = synth.make_pair(5000000, p, 424242)
            oo = mw.opt_init(**kw)
            with mw.Batch(oo, [(t, q)]) as lb:
                lb.upload()
                t0 = time.perf_counter()
                lb.run()
                lb.wait()
                dt = time.perf_counter() - t0
                rl = lb.fetch()[0]
                ll = int(lb.launches)
            assert rl[0] > 0 and mw.cigar2score(oo, rl[3]) == (rl[0], len(t), len(q))
            large[name] = {"workload": what, "s": rl[0], "n_iter": rl[2], "n_cigar": rl[1], "seconds": dt, "gpu_launches": ll,
                           "value": max(len(t), len(q)) * rl[0] / dt, "unit": UNIT,
                           "check": "CIGAR re-scored with mwf_cigar2score: score == s and it consumes both sequences"}
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
                                "wavefront_cells_per_s": sum(r[1] for r in cres) / dt}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
