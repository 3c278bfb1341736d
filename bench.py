#!/usr/bin/env python
"""bench.py -- DP cells/s of the kalign alignment hot path on B200 (BASELINE.json metric).

A "step" is one pass of the DP stages of a multiple alignment -- the anchor-consistency batch
(N x K seq-seq Hirschberg alignments) plus the progressive alignment over the guide tree
(kb200_msa_align == anchor_consistency_build + create_msa_tree of the reference,
lib/src/aln_wrap.c:208-226) -- on one batch of synthetic sequences (kalign_b200/synth.py).

  value : DP cells / device-timed span of the step with the sequences already resident in HBM
          (CUDA events on the engine's stream; max over ranks for N>1)
  e2e   : DP cells / wall time of the public call kb200_kalign (host strings in, aligned host
          strings out: encode, H2D, distances, guide tree, DP stages, D2H, finalise)
  roofline : the sweep kernel (kb_sweep_kernel), algorithmic bytes (25 B per ss/sp
          cell, 128 B per pp cell, +4 B per cell that reads a consistency bonus) / its device time,
          against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline : the unmodified reference (oracle/_ref) on a bounded sample of the same workload
  --impl reference : the reference's own DP stages (anchor_consistency_build + create_msa_tree
          timed inside its unmodified kalign_run_seeded call sequence, all host threads)

Cells are counted once by the GPU engine (paths are bit-identical, so the box recursion and the
cell count are identical for both arms)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from kalign_b200 import synth  # noqa: E402

METRIC = "dp_cells_per_sec"
UNIT = "cells/s"

# workload name -> (synth config, kalign type, consistency anchors, label)
WORKLOADS = {
    "C2": ("C2", 8, 5, "1000 protein x ~400 aa, default mode (consistency K=5)"),
    "C3": ("C3", 2, 5, "10000 16S-like RNA x ~1500 nt, --type rna, default mode (consistency K=5)"),
    "C4": ("C4", 8, 0, "100000 protein x ~300 aa, --fast"),
    "C5": ("C5", 0, 5, "1000 SARS-CoV-2-like genomes x ~30 kb, --type dna, default mode (consistency K=5)"),
}


def effective_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (the GPU boxes
    expose 128 logical CPUs but run under a 16-CPU CFS quota; more threads than that only throttle)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(st):
    """algorithmic bytes of the cells the SWEEP kernel processed (the small-box kernel's cells are
    excluded: they are timed separately)"""
    ss = st["cells_ss"] - st["small_ss"]
    sp = st["cells_sp"] - st["small_sp"]
    pp = st["cells_pp"] - st["small_pp"]
    tot = st["cells_ss"] + st["cells_sp"] + st["cells_pp"]
    big = ss + sp + pp
    bonus = st["cells_bonus"] * (big / tot if tot > 0 else 0.0)
    return 25.0 * (ss + sp) + 128.0 * pp + 4.0 * bonus


def golden_sha(workload, n):
    """SHA-256 of the reference's alignment of this workload (tests/golden/full_<wl>.npz, written by
    tools/gen_golden_full.py from the unmodified reference), or None when no fixture fits"""
    p = os.path.join(ROOT, "tests", "golden", "full_%s.npz" % workload)
    if not os.path.exists(p):
        return None
    g = np.load(p, allow_pickle=False)
    if n is not None and int(g["n"]) != n:
        return None
    return str(g["msa_sha256"])


def issue_model():
    """thread-instructions per DP cell of the sweep kernel families and the DRAM traffic of its
    dominant launch, from the committed ncu captures of this command (tools/ncu_issue.py)"""
    p = os.path.join(ROOT, "profiles", "r02_sweep_issue.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {"inst_per_cell": {"none": 17.1, "sparse": 26.0}, "dram_bytes_per_launch": None,
            "source": "fallback: r01 ncu capture (17.1) / static SASS count (26)"}


def bpm_bench(ctx, seqs, type_):
    """the N x 32 anchor distance matrix of the workload through kb200_distances:
    pairs/s, word-steps/s (text symbols x 64-bit pattern words) and the fraction of the integer
    issue ceiling they use (53.8 executed thread-instructions per word-step, measured with ncu);
    HBM traffic is n + m + 4 bytes per pair -- negligible, reported as a fraction"""
    from kalign_b200 import _lib
    nuc = type_ in (0, 1, 2)
    tbl = {c: i for i, c in enumerate("ACGTU")} if nuc else {c: i % 13 for i, c in enumerate("ACDEFGHIKLMNPQRSTVWY")}
    order = sorted(range(len(seqs)), key=lambda i: -len(seqs[i]))
    lut = np.zeros(256, dtype=np.uint8)
    for c, v in tbl.items():
        lut[ord(c)] = min(v, 12) if not nuc else min(v, 3)
    codes = [lut[np.frombuffer(seqs[i].encode(), dtype=np.uint8)] for i in order]
    flat, offs, lens = _lib.pack(codes)
    n = len(codes)
    rows = np.arange(n, dtype=np.int32)
    cols = np.arange(0, n, max(1, n // 32), dtype=np.int32)[:32]
    ctx.distances(flat, offs, lens, rows, cols)           # warm-up
    s0 = ctx.stats()
    ctx.distances(flat, offs, lens, rows, cols)
    s1 = ctx.stats()
    t = s1["bpm_seconds"] - s0["bpm_seconds"]
    pairs = s1["bpm_pairs"] - s0["bpm_pairs"]
    L = lens.astype(np.int64)
    a, b = np.meshgrid(L, L[cols], indexing="ij")
    text, pat = np.maximum(a, b), np.minimum(np.minimum(a, b), 1024)
    words = (pat + 63) // 64
    wsteps = float(((text + 64 * words - pat) * words).sum())
    bytes_ = float((text + pat + 4).sum())
    hbm, _, _ = peaks()
    return {"pairs_per_sec": pairs / t, "word_steps_per_sec": wsteps / t, "seconds": t, "pairs": pairs,
            "alu_frac": wsteps * 53.8 / t / (148 * 4 * 32 * 1.965e9),
            "alu_note": "53.8 executed thread-instructions per (symbol, 64-bit word) step (ncu, profiles/r02_bpm_kernel_full_C3.csv: "
                        "1.2945e10 warp-instructions x 32 / 7.70e9 word-steps of the C3 N x 32 matrix; the loop body is 101 SASS "
                        "instructions, part of it predicated) against 148 SMs x 4 x 32 lanes x 1.965 GHz",
            "hbm_frac": bytes_ / t / 1e9 / hbm, "kernel": "kb_bpm_kernel"}


def apair_bench(ctx, rows, with_cpu):
    """the N x N identity distances of the realign loop (compute_aln_pairwise_dist,
    lib/src/aln_apair_dist.c:9) on the alignment the timed steps produced, through kb200_aln_pairwise_dist:
    column-pairs/s in the kernels (CUDA events) and for the whole host-pointer call (upload of the rows,
    download of the n x n floats into the caller's row allocations); the unmodified reference's serial loop
    on a bounded sample of the same rows beside it"""
    n = len(rows)
    if n > 20000:
        return {"skipped": "n = %d: the n x n float matrix (%.0f GB) is not benchmarked from python" % (n, 4e-9 * n * n)}
    alnlen = len(rows[0])
    ctx.aln_pairwise_dist(rows[:256])                 # warm-up (module load, pinned pool)
    s0 = ctx.stats()
    t0 = time.perf_counter()
    dm = ctx.aln_pairwise_dist(rows)
    t1 = time.perf_counter()
    s1 = ctx.stats()
    tk = s1["apair_seconds"] - s0["apair_seconds"]
    colpairs = s1["apair_col_pairs"] - s0["apair_col_pairs"]
    out = {"n": n, "alnlen": alnlen, "kernel_seconds": tk, "call_seconds": t1 - t0,
           "col_pairs_per_sec_in_kernel": colpairs / tk if tk > 0 else None,
           "col_pairs_per_sec_call": colpairs / (t1 - t0),
           "issue_frac": (colpairs * 2.108 / tk / (148 * 4 * 32 * 1.965e9)) if tk > 0 else None,
           "issue_note": "2.108 executed thread-instructions per column-pair (ncu, profiles/r02_apair_kernel_full.csv: 2.838e9 "
                         "warp-instructions x 32 / 4.307e10 column-pairs of a 4096 x 5136 alignment; smsp__issue_active 71 %) "
                         "against 148 SMs x 4 x 32 lanes x 1.965 GHz",
           "h2d_bytes": float(n) * ((alnlen + 3) // 4 * 4), "d2h_bytes": 4.0 * n * n, "kernel": "kb_apair_tile_kernel"}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import kbind
        ns = min(n, max(64, int((2.0e9 / max(1, alnlen)) ** 0.5)))      # ~1e9 byte compares: a few seconds
        sample = rows[:ns]
        c0 = time.perf_counter()
        want = kbind.ref_aln_pairwise_dist(sample) if kbind.have_ref() else kbind.oracle_aln_pairwise_dist(sample)
        c1 = time.perf_counter()
        out["cpu_reference"] = {"col_pairs_per_sec": 0.5 * ns * (ns - 1) * alnlen / (c1 - c0), "cores": 1,
                                "kind": "reference" if kbind.have_ref() else "port",
                                "sample": "first %d rows of the same alignment (the reference's loop is serial)" % ns,
                                "identical_to_gpu": bool(np.array_equal(want, dm[:ns, :ns]))}
    return out


def full_reference_record(workload):
    """the unmodified reference's run of the FULL workload, recorded once in the build container by
    tools/gen_golden_full.py (tests/golden/full_<wl>.npz: stage times and the hash the GPU result is
    compared with); not re-run here -- the full C3 takes ~18 minutes on the CPU"""
    p = os.path.join(ROOT, "tests", "golden", "full_%s.npz" % workload)
    if not os.path.exists(p):
        return None
    g = np.load(p, allow_pickle=False)
    t = [float(x) for x in g["times"]]
    return {"n": int(g["n"]), "threads": int(g["threads"]), "dist_tree_s": t[0], "anchor_s": t[1], "tree_aln_s": t[2],
            "total_s": t[3], "dp_stage_seconds": t[1] + t[2], "msa_sha256": str(g["msa_sha256"]),
            "where": "build container (8 cores), tools/gen_golden_full.py"}


def delta(a, b):
    return {k: b[k] - a[k] for k in b}


def ref_sample(wl, n_sample, n_threads):
    """the unmodified reference on a bounded sample of the workload; returns timings"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kbind
    cfg, type_, K, _ = WORKLOADS[wl]
    seqs = synth.config(cfg, n=n_sample)
    run = kbind.RefRun(seqs, n_threads=n_threads, type_=type_, consistency=K, weight=2.0)
    t = run.times()
    rows = run.aligned()
    run.close()
    return seqs, t, rows


def count_cells_gpu(seqs, type_, K):
    from kalign_b200 import _lib
    ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
    m = _lib.Msa(ctx, seqs, n_threads=effective_cpus(), type_=type_, consistency=K, weight=2.0)
    s0 = ctx.stats()
    m.align()
    s1 = ctx.stats()
    rows = m.result()
    m.close()
    ctx.close()
    return s1["dp_cells"] - s0["dp_cells"], rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("KB200_WORKLOAD", "C3"), choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=None, help="override the number of sequences")
    ap.add_argument("--ref-sample", type=int, default=None, help="sequences in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg, type_, K, label = WORKLOADS[args.workload]
    # the CPU quota is shared by all ranks of the node
    host_threads = max(1, effective_cpus() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))))
    default_sample = {"C2": 400, "C3": 160, "C4": 4000, "C5": 6}[args.workload]
    n_sample = args.ref_sample or default_sample

    if args.impl == "reference":
        if rank != 0:
            return 0
        # rank 0 alone runs the reference arm: it may use every host thread of the node
        host_threads = max(1, effective_cpus())
        # bounded sample of the same workload, every step re-runs it
        times = []
        seqs = None
        for it in range(args.warmup + args.steps):
            seqs, t, rows = ref_sample(args.workload, n_sample, host_threads)
            if it >= args.warmup:
                times.append(t["anchor"] + t["tree_aln"])
        cells, gpu_rows = count_cells_gpu(seqs, type_, K)
        identical = (gpu_rows == rows)
        T = float(np.mean(times)) if times else float("nan")
        v = cells / T
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s sample: first %d sequences of %s" % (args.workload, n_sample, label),
                           "timed": "anchor_consistency_build + create_msa_tree inside the reference's own call sequence"},
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_threads, "kind": "reference",
                                 "sample": "%d sequences of %s, %d OpenMP threads (the anchor phase is serial in the reference)" % (n_sample, args.workload, host_threads)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "cells_per_step": cells, "cells_counted_by": "gpu engine, untimed (bit-identical paths)",
                "msa_identical_to_gpu": identical}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    from kalign_b200 import _lib
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    seqs = synth.config(cfg, n=args.n)
    # N>1: STRONG scaling -- the same multiple alignment; the N x K anchor pairs and the tasks of every
    # guide-tree level are sharded across the ranks, position maps / coded paths / merged profiles
    # are all-gathered over NCCL (the library's own communicator); value = total cells / max time
    if world > 1:
        from kalign_b200 import parallel
        uid = parallel.exchange_unique_id(ctx.lib, rank, dist, device=torch.device("cuda", local_rank))
        parallel.attach(ctx, rank, world, uid)
    t0 = time.perf_counter()
    m = _lib.Msa(ctx, seqs, n_threads=host_threads, type_=type_, consistency=K, weight=2.0)
    t_create = time.perf_counter() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        m.align()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    s0 = ctx.stats()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        m.align()
    barrier()
    w1 = time.perf_counter()
    s1 = ctx.stats()
    clocks = sampler.stop()
    d = delta(s0, s1)
    t_dev = d["align_seconds"]
    cells = d["dp_cells"]
    if world > 1:
        tt = torch.tensor([t_dev, cells], dtype=torch.float64, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_dev = float(tmax[0])
        cells = float(tsum[1])
    value = cells / t_dev
    # ---- e2e through the public call, host buffers in / out
    e2e = None
    if rank == 0 or world > 1:
        e_times = []
        py_times = []
        se0 = ctx.stats()
        n_e2e = max(1, min(2, args.steps))
        for it in range(1 + n_e2e):
            if it == 1:
                se0 = ctx.stats()
            a = time.perf_counter()
            tc = []
            rows = ctx.kalign(seqs, n_threads=host_threads, type_=type_, consistency=K, weight=2.0, timing=tc)
            b = time.perf_counter()
            if it >= 1:
                e_times.append(tc[0])
                py_times.append(b - a)
        se1 = ctx.stats()
        de = delta(se0, se1)
        te = float(np.mean(e_times))
        e_cells = de["dp_cells"] / n_e2e
        if world > 1:
            tt = torch.tensor([te, e_cells], dtype=torch.float64, device="cuda")
            tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            te, e_cells = float(tmax[0]), float(tsum[1])
        e2e = {"value": e_cells / te, "unit": UNIT, "seconds_per_call": te,
               "h2d_bytes_per_step": de["h2d_bytes"] / n_e2e, "d2h_bytes_per_step": de["d2h_bytes"] / n_e2e,
               "includes": "encode, H2D, bpm distances, guide tree, DP stages, finalise, D2H, malloc'd result rows",
               "timed": "wall clock around the C-ABI call kb200_kalign(char** seq, int* len, ...) -> char*** aligned (the call "
                        "that replaces kalign(), lib/include/kalign/kalign.h:45; the reference arm times kalign's stages the same "
                        "way, in C); the python wrapper's str <-> bytes conversion around it is reported separately",
               "seconds_per_call_incl_python_marshalling": float(np.mean(py_times))}
        ok = all(r.replace("-", "") == s for r, s in zip(rows, seqs))
        e2e["residues_preserved"] = bool(ok)
        e2e["msa_sha256"] = synth.msa_sha256(rows)
    peak, peak_src, _ = peaks()
    abytes = algorithmic_bytes(d)
    ach = abytes / d["sweep_seconds"] / 1e9 if d["sweep_seconds"] > 0 else 0.0
    # ---- roofline of the dominant kernel (kb_sweep_kernel).  The DP state of a strip lives in
    # registers and only one row in 32*K touches memory, so the kernel is bound by ISSUE SLOTS, not by
    # HBM: ceiling = SMs x 4 schedulers x 32 lanes x SM clock thread-instructions/s; achieved = cells
    # the kernel processed x thread-instructions per cell (measured with ncu on this command,
    # profiles/r02_sweep_issue.json: smsp__inst_executed x 32 / cells of the launch, per kernel family)
    # / the kernel's device time measured live with CUDA events.  The SURVEY 8d algorithmic-bytes
    # figure against the measured HBM peak is kept beside it (hbm_model): it exceeds 1 because the
    # model counts state rows that never leave the register file.
    im = issue_model()
    in_kernel = d["dp_cells"] - d["small_ss"] - d["small_sp"] - d["small_pp"]
    tot_cells = d["cells_ss"] + d["cells_sp"] + d["cells_pp"]
    bonus_in_kernel = d["cells_bonus"] * (in_kernel / tot_cells if tot_cells > 0 else 0.0)
    tinst = (in_kernel - bonus_in_kernel) * im["inst_per_cell"]["none"] + bonus_in_kernel * im["inst_per_cell"]["sparse"]
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    n_sm = 148
    issue_peak = n_sm * 4 * 32 * sm_clock
    issue_ach = tinst / d["sweep_seconds"] if d["sweep_seconds"] > 0 else 0.0
    roofline = {"bound": "issue", "kernel": "kb_sweep_kernel", "achieved": issue_ach / 1e9, "peak": issue_peak / 1e9,
                "unit": "G thread-instructions/s", "frac": issue_ach / issue_peak,
                "traffic": im.get("dram_bytes_per_launch"),
                "traffic_note": im.get("traffic_note"),
                "peak_source": "%d SMs x 4 schedulers x 32 lanes x %.0f MHz (median SM clock sampled during the timed region)" % (n_sm, sm_clock / 1e6),
                "inst_per_cell": im["inst_per_cell"], "inst_per_cell_source": im["source"],
                "cells_per_sec_ceiling": issue_peak / im["inst_per_cell"]["none"],
                "hbm_model": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                              "peak_source": peak_src, "algorithmic_bytes_per_step": abytes / max(1, args.steps),
                              "note": "SURVEY 8d contract: 25 B per ss/sp cell, 128 B per pp cell, +4 B per bonus cell; "
                                      "measured DRAM traffic is ~1000x lower (traffic), so this fraction is not an efficiency"},
                "kernel_seconds_per_step": d["sweep_seconds"] / max(1, args.steps),
                "kernel_share_of_step": d["sweep_seconds"] / d["align_seconds"] if d["align_seconds"] > 0 else None,
                "cells_in_kernel_per_step": in_kernel / max(1, args.steps),
                "cells_per_sec_in_kernel": in_kernel / d["sweep_seconds"] if d["sweep_seconds"] > 0 else None,
                "small_box_kernel_seconds_per_step": d["small_seconds"] / max(1, args.steps)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / max(1, args.steps), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, label), "nseq": len(seqs),
                       "parallelism": "1 GPU" if world == 1 else "tasks of each tree level + anchor pairs sharded over %d GPUs, NCCL all-gather per level" % world,
                       "l2": "inputs larger than L2 (row buffers + work lists are GBs; no flush needed)",
                       "step": "kb200_msa_align = anchor batch + progressive alignment, sequences resident in HBM"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(d["n_launches"]),
            "roofline": roofline,
            "cells_per_step": cells / max(1, args.steps),
            "cells_by_kind": {"ss": d["cells_ss"], "sp": d["cells_sp"], "pp": d["cells_pp"], "with_bonus": d["cells_bonus"]},
            "wall_s_timed_region": w1 - w0, "create_seconds": t_create,
            "collectives_per_step": d["n_collectives"] / max(1, args.steps), "collective_bytes_per_step": d["collective_bytes"] / max(1, args.steps),
            "dp_round_seconds_per_step": d["dp_seconds"] / max(1, args.steps)}
    try:
        # single-GPU line only: with a communicator attached kb200_distances is a collective (row shard)
        line["bpm"] = bpm_bench(ctx, seqs, type_) if world == 1 else None
    except Exception as e:  # noqa: BLE001
        line["bpm"] = {"error": repr(e)}
    # ---- identity with the reference: hash of the alignment the timed steps produced (every rank
    #      holds the full result) against the golden hash of the unmodified reference's alignment
    step_rows = m.result()
    try:
        line["apair"] = apair_bench(ctx, step_rows, rank == 0 and not args.no_cpu_baseline) if world == 1 else None
    except Exception as e:  # noqa: BLE001
        line["apair"] = {"error": repr(e)}
    sha = synth.msa_sha256(step_rows)
    want = golden_sha(args.workload, len(seqs))
    same_all = True
    if world > 1:
        # every rank must hold the same alignment: compare the hashes' first 8 bytes
        hv = torch.tensor([int(sha[:15], 16)], dtype=torch.int64, device="cuda")
        hmax = hv.clone(); dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        hmin = hv.clone(); dist.all_reduce(hmin, op=dist.ReduceOp.MIN)
        same_all = bool(int(hmax[0]) == int(hmin[0]))
    line["msa_sha256"] = sha
    line["msa_sha256_reference"] = want
    line["msa_identical_to_reference"] = (None if want is None else bool(sha == want and same_all and
                                          (e2e is None or e2e.get("msa_sha256") == want)))
    line["msa_identical_on_all_ranks"] = same_all
    if want is not None and not line["msa_identical_to_reference"]:
        sys.stderr.write("bench.py: ALIGNMENT DIFFERS FROM THE REFERENCE (sha %s, want %s)\n" % (sha, want))
    # ---- CPU baseline beside it (rank 0, N=1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sseqs, t, ref_rows = ref_sample(args.workload, n_sample, host_threads)
            scells, gpu_rows = count_cells_gpu(sseqs, type_, K)
            T = t["anchor"] + t["tree_aln"]
            full = full_reference_record(args.workload)
            if full is not None and len(seqs) == full["n"]:
                full["value"] = (cells / max(1, args.steps)) / full["dp_stage_seconds"]
                full["unit"] = UNIT
            line["cpu_baseline"] = {"value": scells / T, "unit": UNIT, "cores": host_threads, "kind": "reference",
                                    "full_config_run": full,
                                    "sample": "first %d sequences of %s; anchor_consistency_build %.2fs (serial in the reference) + create_msa_tree %.2fs on %d threads" % (n_sample, args.workload, t["anchor"], t["tree_aln"], host_threads),
                                    "msa_identical_to_gpu": gpu_rows == ref_rows}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads, "kind": "reference",
                                    "sample": "failed: %r" % (e,)}
    m.close()
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
