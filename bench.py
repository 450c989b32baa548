#!/usr/bin/env python3
"""Benchmark of the CAVI hot path (BASELINE.json: CAVI windows/sec at 1/2/4/8 B200 vs the
host-CPU reference).

A *step* is one pass of the hot path over one batch of synthetic input: every window of one
synthetic genome per GPU (workload ``hiv5k_uqs`` = BASELINE configs[1]: HIV-1 9.7 kb, window
201, 5,000x coverage, 5 haplotypes, use_quality_scores, K = 100 -> 145 windows), each run to
the end of its CAVI loop.

  value   windows (x restarts) per second with the window tensors already resident in HBM:
          only the device time of the iteration loop (CUDA events) is counted
  e2e     the same through the host-buffer API: pinned host -> device copy of all inputs, the
          loop, device -> host copy of the results, every step
  roofline  the dominant kernel over a whole run, every launch bracketed by CUDA events on the
          launching stream (a separate, untimed pass over the same workload): algorithmic HBM
          bytes and executed / algorithmic FP64 flops per launch against the measured HBM copy
          bandwidth (MEASURED_PEAKS.json) and the FP64 tensor (DMMA) ceiling measured on this
          pool; the all-clusters-live first updates are reported separately (tensor bound)
  cpu_baseline  the NumPy oracle port (same einsum calls as the reference) on one host core,
          bounded sample, extrapolated with the iteration counts the GPU run took

``--impl reference`` times the reference's CPU algorithm (the oracle port; the reference is
pure Python/NumPy and cannot be shipped to the GPU box) with its own parallelism: a
multiprocessing Pool over windows with cpu_count()-1 single-threaded workers
(viloca/shotgun.py:527-538).

Multi-GPU: one process per GPU (torchrun), windows sharded by the host-side longest-first
scheduler, no data-path collective; weak scaling with fixed per-GPU work (one copy of the same
synthetic genome per GPU; the shards mix windows of the copies).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cavi_windows_per_sec"
UNIT = "windows/s"
ITER_TABLE = os.path.join(ROOT, "profiles", "bench_iterations.json")
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "dram_traffic.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hiv5k_uqs")
    ap.add_argument("--windows", type=int, default=None, help="windows per synthetic genome (default: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--record-iterations", action="store_true", help="write profiles/bench_iterations.json")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_items(spec, n_windows, world):
    """Global work list of a weak-scaling job: one copy of the synthetic genome per GPU, every
    window an independent item; longest-first partition over the ranks (no collective).  The
    copies are identical (seed 0), so per-GPU work does not depend on the number of GPUs."""
    from viloca_b200 import scheduler
    items = [(sample, w) for sample in range(world) for w in range(n_windows)]
    costs = [float(spec.coverage) * spec.window_length * spec.n_clusters * spec.n_starts] * len(items)
    return items, scheduler.lpt_partition(costs, world)


def operand_stats(windows):
    """Per window: unique reads N, padded reads, dense operand columns (real / padded to tiles
    of 16) and sparse entries, from the library's own layout planner (host only)."""
    import ctypes as C
    from viloca_b200 import _capi as capi
    out = []
    for w in windows:
        d = w.desc()
        ncols, nsp = C.c_int32(0), C.c_int64(0)
        col_lb = np.full(capi.lib.vl_window_plan_capacity(w.length), -1, dtype=np.int32)
        capi.check(capi.lib.vl_window_plan(C.byref(d), C.byref(ncols), C.byref(nsp), capi.i32_ptr(col_lb)))
        out.append({"N": w.n_reads, "Npad": -(-w.n_reads // 64) * 64, "L": w.length, "K": w.n_clusters,
                    "NCs": int(ncols.value), "NC": int((col_lb[:ncols.value] >= 0).sum()), "nnz": int(nsp.value)})
    return out


def launch_work(ops, stats):
    """Work of ONE launch of each contraction kernel over the restarts still running
    (stats = CaviBatch.problem_stats(): running, 8-cluster tiles, clusters in the layout).
    executed flops: what the DMMAs do (padded reads / columns / cluster tiles);
    algorithmic flops: 2 * N * columns * clusters-in-layout of the formulation the kernels use;
    dense-equivalent flops: SURVEY.md 8(d)'s 2*N*K*L*B of the reference's einsum (B = 5);
    bytes: compulsory HBM traffic of the launch (operand streamed once, responsibilities and
    haplotype tables read / written once)."""
    ex = alg = dense = bytes_h = bytes_z = 0.0
    assert len(ops) == len(stats)
    for o, (running, kt, nlay, _) in zip(ops, stats):
        if not running:
            continue
        ex += 2.0 * o["Npad"] * o["NCs"] * 8 * kt
        alg += 2.0 * o["N"] * o["NC"] * nlay
        dense += 2.0 * o["N"] * o["K"] * o["L"] * 5
        dm = 8.0 * o["N"] * o["NC"]
        zb = 8.0 * o["N"] * nlay
        tab = 8.0 * o["L"] * 5 * nlay
        gm = 8.0 * o["NC"] * nlay
        bytes_h += dm + zb + 8.0 * o["N"] + 2 * tab + gm + 12.0 * o["nnz"]
        bytes_z += dm + gm + zb + 12.0 * o["nnz"]
    return {"executed": ex, "algorithmic": alg, "dense_equivalent": dense, "bytes_haplo": bytes_h, "bytes_cluster": bytes_z}


def profiled_run(batch, ops):
    """One whole run with every launch timed (vl_batch_profile_step), in the chunk pattern of
    vl_batch_run.  Returns per-kernel totals over the run and over its all-clusters-live start."""
    batch.upload()
    tot = {k: {"ms": 0.0, "launches": 0, "executed": 0.0, "algorithmic": 0.0, "dense_equivalent": 0.0, "bytes": 0.0}
           for k in ("haplo", "cluster")}
    first = None
    n_chunks = 0
    while True:
        before = batch.problem_stats()
        if not before[:, 0].any():
            break
        n = 4 if n_chunks < 6 else 32
        prof = batch.profile_step(n)
        after = batch.problem_stats()
        wb, wa = launch_work(ops, before), launch_work(ops, after)
        for k in ("haplo", "cluster"):
            ms, cnt = prof[k]
            tot[k]["ms"] += ms
            tot[k]["launches"] += cnt
            for key in ("executed", "algorithmic", "dense_equivalent"):
                tot[k][key] += n * 0.5 * (wb[key] + wa[key])
            tot[k]["bytes"] += n * 0.5 * (wb["bytes_" + k] + wa["bytes_" + k])
        if n_chunks == 0:
            first = {k: dict(v) for k, v in tot.items()}
        n_chunks += 1
    return tot, first


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU legs (the only places that may execute oracle/)
# ------------------------------------------------------------------------------------------

def _oracle_window_iterations(win, n_iter):
    """Seconds for n_iter oracle iterations (update + ELBO, exactly the reference's per-iteration
    work incl. its redundant einsums) on one window, single thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _bridge as br
    w, ref, data = br.dense_inputs(win)
    init = br.oracle_init(win)
    t0 = time.perf_counter()
    traj = br.orc.run(win.mode, w, ref, data, init, convergence_threshold=win.conv_thres, max_iterations=n_iter)
    return time.perf_counter() - t0, traj.n_updates


def _pool_worker(arg):
    name, index, n_iter = arg
    os.environ["OMP_NUM_THREADS"] = "1"
    from viloca_b200 import synthetic
    win = synthetic.make_window(synthetic.WORKLOADS[name], index)
    dt, n = _oracle_window_iterations(win, n_iter)
    return dt, n, win.n_reads


def load_iteration_table(name):
    try:
        return json.load(open(ITER_TABLE))[name]
    except (OSError, KeyError, ValueError):
        return None


def cpu_baseline_single_core(name, windows, n_updates, budget_s=12.0):
    """One core, bounded: a few dozen oracle iterations on the first windows, extrapolated to the
    whole workload with the iteration counts the GPU run actually took (cost/iteration is
    linear in N_unique)."""
    spent, read_iters, sample = 0.0, 0.0, []
    n_iter = 24
    for i, win in enumerate(windows[:8]):
        dt, n = _oracle_window_iterations(win, n_iter)
        spent += dt
        read_iters += n * win.n_reads
        sample.append(f"w{i}:{n}it")
        if spent > budget_s:
            break
    sec_per_read_iter = spent / read_iters
    total = sum(u * w.n_reads for w, u in zip(windows, n_updates)) * sec_per_read_iter
    return {"value": len(windows) / total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle port, {'+'.join(sample)} of {name} ({spent:.1f}s), extrapolated with the GPU run's "
                      f"per-window iteration counts", "sec_per_read_iteration": sec_per_read_iter}


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import multiprocessing as mp
    from viloca_b200 import synthetic
    spec = synthetic.WORKLOADS[args.workload]
    procs = max((os.cpu_count() or 2) - 1, 1)          # shotgun.py:527
    n_iter = 10
    table = load_iteration_table(args.workload)
    mean_updates = float(np.mean(table["n_updates"])) if table else 300.0
    mean_reads = float(np.mean(table["n_reads"])) if table else None
    per_step, step_s = [], []
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=procs) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(args.workload, (step * procs + i) % spec.n_windows, n_iter) for i in range(procs)]
            t0 = time.perf_counter()
            out = pool.map(_pool_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                step_s.append(dt)
                # windows/s = (window-iterations done / iterations a window needs) / time, corrected
                # for the sampled windows' size relative to the workload mean
                scale = 1.0
                if mean_reads:
                    scale = float(np.mean([o[2] for o in out])) / mean_reads
                per_step.append(sum(o[1] for o in out) / mean_updates * scale / dt)
    value = float(np.mean(per_step))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(step_s)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "mode": spec.mode, "genome_length": spec.genome_length,
                   "window_length": spec.window_length, "windows_per_gpu": spec.n_windows, "coverage": spec.coverage,
                   "n_clusters": spec.n_clusters, "n_starts": spec.n_starts, "haplotypes": spec.n_haplotypes,
                   "parallelism": f"multiprocessing.Pool({procs}) over windows, one single-threaded process per window "
                                  "(viloca/shotgun.py:527-538)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"Pool({procs}) x {n_iter} iterations per window per step; converted with "
                                   f"mean iterations/window = {mean_updates:.1f} "
                                   f"({'profiles/bench_iterations.json' if table else 'assumed'})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from viloca_b200 import engine, synthetic

    rank, world, local_rank = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    spec = synthetic.WORKLOADS[args.workload]
    n_windows = args.windows or spec.n_windows
    items, parts = shard_items(spec, n_windows, world)
    mine = [items[i] for i in parts[rank]]
    # weak scaling with FIXED per-GPU work: every sample is the same synthetic genome (seed 0), so each
    # rank runs every window index exactly once whatever the number of GPUs
    windows = [synthetic.make_window(spec, w, master_seed=0).pin() for sample, w in mine]
    n_problems = sum(w.n_starts for w in windows)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch = engine.CaviBatch(windows, device=device)
    h2d = batch.input_bytes()
    d2h = 0
    launches = 0
    launches_run = 0
    outcomes = None
    for _ in range(args.warmup):
        batch.upload(); batch.run(); outcomes = batch.fetch(pinned=True)
    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_ms = 0.0
    ev0.record()
    for _ in range(args.steps):
        batch.upload()
        launches += 2 * len(windows) + 1                # operand preparation (2 kernels per window) + init
        n_run = batch.run()
        launches += n_run
        launches_run += n_run
        run_ms += batch.last_device_ms()
        outcomes = batch.fetch(pinned=True)            # results land in page-locked buffers owned by the batch
        launches += 2 * len(windows)                    # mean_haplo / mean_cluster export per window
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    for o, w in zip(outcomes, windows):
        d2h += w.n_clusters * w.length * 5 * 8 + w.n_reads * w.n_clusters * 8 + 2 * w.n_clusters * 8 + 128
        d2h += sum(len(r.history_elbo) for r in o.restarts) * 8

    # per-kernel roofline numbers: a separate pass with every launch bracketed by CUDA events
    ops = operand_stats(windows)
    ops_per_problem = [o for o, w in zip(ops, windows) for _ in range(w.n_starts)]     # problem = (window, restart)
    prof_tot, prof_first = profiled_run(batch, ops_per_problem)
    n_updates = [o.restarts[o.best_restart].n_updates for o in outcomes]
    all_updates = [r.n_updates for o in outcomes for r in o.restarts]
    read_iters = sum(r.n_updates * w.n_reads for o, w in zip(outcomes, windows) for r in o.restarts)
    raw_read_iters = sum(r.n_updates * getattr(w, "n_raw", w.n_reads) for o, w in zip(outcomes, windows) for r in o.restarts)

    t = torch.tensor([run_ms, e2e_ms, float(n_problems), float(read_iters), float(raw_read_iters), float(launches),
                      float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    run_ms_max, e2e_ms_max = float(tmax[0]), float(tmax[1])
    total_problems = float(tsum[2])

    if args.record_iterations and rank == 0:
        os.makedirs(os.path.dirname(ITER_TABLE), exist_ok=True)
        table = {}
        if os.path.exists(ITER_TABLE):
            table = json.load(open(ITER_TABLE))
        table[args.workload] = {"n_updates": all_updates, "n_reads": [w.n_reads for w in windows],
                                "n_raw": [getattr(w, "n_raw", w.n_reads) for w in windows]}
        json.dump(table, open(ITER_TABLE, "w"))

    if rank == 0:
        kern_ms = {k: v["ms"] for k, v in prof_tot.items()}
        try:
            fp64_peak = max(r["tflops"] for r in json.load(open(FP64_PEAK_FILE))["results"] if r["kind"] == "dmma884")
            fp64_src = "DMMA m8n8k4 register loop measured on this pool, profiles/r01_fp64_peak.json"
        except (OSError, ValueError, KeyError):
            fp64_peak, fp64_src = 37.0, "fallback: nominal B200 FP64 tensor peak"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs_sustained"])
            hbm_src = "MEASURED_PEAKS.json hbm_gbs_sustained"
        except (OSError, ValueError, KeyError):
            try:
                hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
                hbm_src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth; no sustained figure in the file)"
            except (OSError, ValueError, KeyError):
                hbm_peak, hbm_src = 6550.0, "fallback of B200_PROFILING.md"
        traffic = None
        try:
            traffic = json.load(open(TRAFFIC_FILE)).get(args.workload, {}).get("chunk")
        except (OSError, ValueError):
            pass

        def rates(t, ms=None):
            sec = max(t["ms"] if ms is None else ms, 1e-9) * 1e-3
            return {"ms": t["ms"] if ms is None else ms, "launches": t["launches"],
                    "gbytes_per_s": t["bytes"] / sec * 1e-9, "executed_tflops": t["executed"] / sec * 1e-12,
                    "algorithmic_tflops": t["algorithmic"] / sec * 1e-12,
                    "dense_equivalent_tflops": t["dense_equivalent"] / sec * 1e-12}

        # The timed runs execute both contractions of every update inside vl_chunk_kernel (one
        # launch per chunk of updates and kernel class): its work is the sum of the two phases,
        # its duration the device time of the timed loop (launch gaps between chunks included).
        both = {k: prof_tot["haplo"][k] + prof_tot["cluster"][k] for k in ("bytes", "executed", "algorithmic", "dense_equivalent")}
        both["ms"], both["launches"] = run_ms / args.steps, launches_run / args.steps
        fused = rates(both)
        frac_hbm, frac_fp = fused["gbytes_per_s"] / hbm_peak, fused["executed_tflops"] / fp64_peak
        bound = "hbm" if frac_hbm >= frac_fp else "tensor"
        per_phase = {k: rates(prof_tot[k]) for k in ("haplo", "cluster")}
        start = {k: rates(prof_first[k]) for k in ("haplo", "cluster")}
        roofline = {
            "bound": bound, "kernel": "vl_chunk_kernel (haplotype-update + responsibility-update tiles of a chunk of updates)",
            "achieved": fused["gbytes_per_s"] if bound == "hbm" else fused["executed_tflops"],
            "peak": hbm_peak if bound == "hbm" else fp64_peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
            "frac": max(frac_hbm, frac_fp), "traffic": traffic,
            "peak_source": hbm_src if bound == "hbm" else fp64_src,
            "timed_run": {**fused, "frac_hbm": frac_hbm, "frac_fp64_executed": frac_fp},
            "per_phase_separate_launches": {k: {**v, "frac_hbm": v["gbytes_per_s"] / hbm_peak,
                                                "frac_fp64_executed": v["executed_tflops"] / fp64_peak}
                                            for k, v in per_phase.items()},
            "all_clusters_live_start": {k: {**v, "frac_fp64_executed": v["executed_tflops"] / fp64_peak}
                                        for k, v in start.items()},
            "fp64_peak_tflops": fp64_peak, "fp64_peak_source": fp64_src, "hbm_peak_gbs": hbm_peak,
            "convention": "work per launch = sum over the restarts still running (separate whole-run pass over the same "
                          "workload, vl_batch_profile_step, every launch bracketed by CUDA events, state sampled every "
                          "4-32 updates); bytes = compulsory HBM bytes (dense operand N x columns x 8 streamed once per "
                          "contraction, responsibilities / haplotype tables once); executed flops = DMMA work incl. "
                          "padding; algorithmic flops = 2 N columns clusters-in-layout; dense-equivalent = SURVEY 8(d)'s "
                          "2 N K L B (B = 5) of the reference's einsum, which this formulation does not execute. timed_run "
                          "divides the whole-run work by the device time of the timed loop; per_phase_separate_launches "
                          "are the two phases launched separately (two launches per update, the profiling path); "
                          "all_clusters_live_start = first 4 updates, all K clusters in the layout (tensor bound)",
        }
        line = {
            "metric": METRIC, "value": total_problems * args.steps / (run_ms_max * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": run_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "mode": spec.mode, "genome_length": spec.genome_length,
                       "window_length": spec.window_length, "windows_per_gpu": len(windows), "coverage": spec.coverage,
                       "n_unique_mean": float(np.mean([w.n_reads for w in windows])), "n_clusters": spec.n_clusters,
                       "n_starts": spec.n_starts, "haplotypes": spec.n_haplotypes, "parallelism": f"windows x{world}", "per_gpu_work": f"the same {n_windows}-window genome on every GPU (seed 0)",
                       "dense_columns_mean": float(np.mean([o["NC"] for o in ops])),
                       "sparse_entries_mean": float(np.mean([o["nnz"] for o in ops])),
                       "l2": "inputs larger than L2 (working set %.0f MB per GPU, re-uploaded every step)" % (batch.workspace_bytes / 1e6)},
            "e2e": {"value": total_problems * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(float(tsum[6])), "d2h_bytes_per_step": int(float(tsum[7])),
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(float(tsum[5])),
            "read_iterations_per_sec": float(tsum[3]) * args.steps / (run_ms_max * 1e-3),
            "raw_read_iterations_per_sec": float(tsum[4]) * args.steps / (run_ms_max * 1e-3),
            "iterations": {"mean": float(np.mean(all_updates)), "max": int(max(all_updates)), "min": int(min(all_updates))},
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single_core(args.workload, windows, n_updates)
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
