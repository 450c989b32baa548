#!/usr/bin/env python3
"""Benchmark of the CAVI hot path (BASELINE.json: CAVI windows/sec at 1/2/4/8 B200 vs the
host-CPU reference).

A *step* is one pass of the hot path over one synthetic genome: every window of the workload,
each run to the end of its CAVI loop.  Default workload ``sars_amplicon`` = BASELINE configs[3],
the largest configuration that fits one GPU (SARS-CoV-2 amplicon mode, ~400 bp inserts, 20,000x
coverage, 10 lineages, use_quality_scores, K = 100 -> 98 windows); ``hiv5k_uqs`` (configs[1]) and a
bounded shard of ``sars_ont`` (configs[4]) are measured after it and reported under ``extra``.

  value   windows (x restarts) per second with the window tensors already resident in HBM:
          only the device time of the iteration loop (CUDA events) is counted
  e2e     the same through the host-buffer API: pinned host -> device copy of all inputs, the
          loop, device -> host copy of the results, every step
  roofline  the dominant kernel over a whole run, every launch bracketed by CUDA events on the
          launching stream (a separate, untimed pass over the same workload): algorithmic HBM
          bytes and executed / algorithmic FP64 flops per launch against the measured HBM copy
          bandwidth (MEASURED_PEAKS.json) and the FP64 tensor (DMMA) ceiling measured in the same
          run (vl_debug_fp64_peak); the all-clusters-live first updates are reported separately
  cpu_baseline  the reference's own update / ELBO functions (baseline/_ref, unmodified) on one host
          core, bounded sample, extrapolated with the iteration counts the GPU run took

``--impl reference`` times the reference's CPU implementation with its own parallelism: one
single-threaded process per window, cpu_count()-1 of them side by side (viloca/shotgun.py:527-538),
each running a bounded number of CAVI iterations of the unmodified reference modules
(baseline/ref_driver.py) on a window generated before the clock starts.

Multi-GPU: one process per GPU (torchrun); the windows of ONE genome are sharded over the ranks by
the host-side longest-first scheduler (cost = unique reads x window length x clusters x restarts x
predicted iterations), no data-path collective: strong scaling.
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
    ap.add_argument("--workload", default="sars_amplicon")
    ap.add_argument("--no-extra", action="store_true", help="skip the hiv5k_uqs / sars_ont lines reported under 'extra'")
    ap.add_argument("--windows", type=int, default=None, help="windows per synthetic genome (default: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--waves", type=float, default=0.0, help="run in HBM-bounded waves of at most this many GiB each (default: only when the shard does not fit the GPU)")
    ap.add_argument("--record-iterations", action="store_true", help="write profiles/bench_iterations.json")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def config_of(spec, n_windows):
    """The workload as both arms report it (identical keys and values in both lines)."""
    return {"workload": spec.name, "mode": spec.mode, "genome_length": spec.genome_length,
            "window_length": spec.window_length, "windows": n_windows, "coverage": spec.coverage,
            "n_clusters": spec.n_clusters, "n_starts": spec.n_starts, "haplotypes": spec.n_haplotypes,
            "l2": "inputs larger than L2 (every window's tensors are re-uploaded each step)"}


def predicted_costs(spec, n_windows):
    """Cost of every window for the longest-first scheduler: reads x length x clusters x restarts
    x predicted iterations.  The iteration predictor is the committed table of a previous run of
    the same synthetic genome (profiles/bench_iterations.json: profile-guided, the counts vary
    10..17178 between windows); without it every window counts the same."""
    table = load_iteration_table(spec.name)
    base = float(spec.coverage) * spec.window_length * spec.n_clusters * spec.n_starts
    if not table or len(table.get("n_updates", [])) != n_windows * spec.n_starts:
        return [base] * n_windows, "uniform (no iteration table)"
    it = np.asarray(table["n_updates"], dtype=np.float64).reshape(n_windows, spec.n_starts).sum(axis=1)
    nr = np.asarray(table["n_reads"], dtype=np.float64)
    ln = np.asarray(table.get("length", [spec.window_length] * n_windows), dtype=np.float64)
    return list(nr * ln * spec.n_clusters * it), "profile-guided (profiles/bench_iterations.json)"


def shard_windows(spec, n_windows, world):
    """Strong scaling: the windows of ONE genome, longest first, over the ranks (no collective)."""
    from viloca_b200 import scheduler
    costs, how = predicted_costs(spec, n_windows)
    return scheduler.lpt_partition(costs, world), how


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

def load_iteration_table(name):
    try:
        return json.load(open(ITER_TABLE))[name]
    except (OSError, KeyError, ValueError):
        return None


def genome_work(spec, n_windows, windows=None, n_updates=None):
    """sum over the windows of the genome of (iterations to the end of the loop) x unique reads x
    window length: the work the reference's dense einsums do is proportional to it.
    From the GPU run's own counts when given (trajectories are parity-checked identical), else
    from the committed table of a previous run."""
    if windows is not None:
        return float(sum(u * w.n_reads * w.length for w, u in zip(windows, n_updates))), "this run's iteration counts"
    table = load_iteration_table(spec.name)
    if not table:
        return None, "no iteration table"
    it = np.asarray(table["n_updates"], dtype=np.float64).reshape(n_windows, -1).sum(axis=1)
    nr = np.asarray(table["n_reads"], dtype=np.float64)
    ln = np.asarray(table.get("length", [spec.window_length] * n_windows), dtype=np.float64)
    return float((it * nr * ln).sum()), "profiles/bench_iterations.json"


def cpu_baseline_single_core(spec, windows, n_updates_per_window, n_genome_windows, budget_s=15.0):
    """One core, bounded: a few iterations of the reference's loop body on the first windows,
    extrapolated to the whole genome with the iteration counts the GPU run actually took."""
    from baseline import ref_driver
    spent, work, sample = 0.0, 0.0, []
    n_iter = 6 if windows[0].n_reads * windows[0].length > 400000 else 24
    for i, win in enumerate(windows[:8]):
        tag = ""
        if win.n_reads * win.length > 20_000_000:          # one iteration of the full window would take minutes
            n_iter = 2
            win = ref_driver.first_reads(win, 20_000_000 // win.length)
            tag = f"(first {win.n_reads} reads)"
        dt, n, _ = ref_driver.time_iterations(win, n_iter)
        spent += dt
        work += n * win.n_reads * win.length
        sample.append(f"w{i}{tag}:{n}it")
        if spent > budget_s:
            break
    sec_per_unit = spent / work
    total_work, src = genome_work(spec, n_genome_windows, windows, n_updates_per_window)
    scale = n_genome_windows / len(windows)          # (a rank's shard at N > 1; the leg only runs at N = 1)
    return {"value": n_genome_windows / (total_work * scale * sec_per_unit), "unit": UNIT, "cores": 1,
            "kind": ref_driver.kind(),
            "sample": f"{'+'.join(sample)} of {spec.name} ({spent:.1f}s of update_eqs.update + elbo_eqs.compute_elbo, "
                      f"baseline/_ref unmodified), extrapolated with {src}",
            "sec_per_read_base_iteration": sec_per_unit}


def _reference_worker(name, index, n_iter, n_rounds, barrier, queue):
    """One window's CAVI process of the reference's pool: generate the window (before the clock),
    then n_rounds times: wait for the start signal, run n_iter iterations, report."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from baseline import ref_driver
    from viloca_b200 import synthetic
    import dataclasses
    win = synthetic.make_window(dataclasses.replace(synthetic.WORKLOADS[name], n_starts=1), index)   # (restart 0 is what is timed)
    if win.n_reads * win.length > 20_000_000:              # one iteration of the full window would take minutes: bounded sample
        win = ref_driver.first_reads(win, 20_000_000 // win.length)
    for r in range(n_rounds):
        barrier.wait()
        dt, n, elbo = ref_driver.time_iterations(win, n_iter)
        queue.put((r, index, dt, n, win.n_reads, win.length, elbo))


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import multiprocessing as mp
    from baseline import ref_driver
    from viloca_b200 import synthetic
    spec = synthetic.WORKLOADS[args.workload]
    n_windows = args.windows or spec.n_windows
    procs = max((os.cpu_count() or 2) - 1, 1)          # shotgun.py:527
    heavy = spec.coverage * spec.window_length > 2_000_000
    n_iter = 2 if spec.coverage * spec.window_length > 20_000_000 else (4 if heavy else 10)
    rounds = args.warmup + args.steps
    ctx = mp.get_context("fork")
    barrier, queue = ctx.Barrier(procs + 1), ctx.Queue()
    stride = max(n_windows // procs, 1)
    workers = [ctx.Process(target=_reference_worker, args=(args.workload, (i * stride) % n_windows, n_iter, rounds, barrier, queue),
                           daemon=True) for i in range(procs)]
    for w in workers:
        w.start()
    rates, step_s, sampled = [], [], set()
    for r in range(rounds):
        barrier.wait()                                  # every worker holds its window: the clock starts here
        t0 = time.perf_counter()
        out = [queue.get() for _ in range(procs)]
        dt = time.perf_counter() - t0
        if r >= args.warmup:
            step_s.append(dt)
            rates.append(sum(o[3] * o[4] * o[5] for o in out) / dt)     # read-base-iterations per second of the pool
            sampled.update(o[1] for o in out)
    for w in workers:
        w.join(timeout=10)
    rate = float(np.mean(rates))
    total_work, src = genome_work(spec, n_windows)
    if total_work is None:
        print(json.dumps({"impl": "reference", "unavailable": "no iteration table (profiles/bench_iterations.json) for " + args.workload}))
        return
    value = n_windows / (total_work / rate)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(step_s)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(spec, n_windows),
        "parallelism": f"{procs} single-threaded processes side by side, one window each (viloca/shotgun.py:527-538)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": ref_driver.kind(),
                         "sample": f"per step {procs} windows x {n_iter} iterations of the loop body of run_cavi (cavi.py:120-169; "
                                   f"update_eqs.update + elbo_eqs.compute_elbo of baseline/_ref, unmodified), windows built before "
                                   f"the clock starts; windows/s = windows of the genome / (sum over its windows of iterations x "
                                   f"reads x length [{src}] / measured read-base-iterations per second of the pool), i.e. a "
                                   f"perfectly balanced pool (the longest window alone would take longer)",
                         "read_base_iterations_per_sec": rate, "windows_sampled": sorted(sampled)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------

def timed_steps(batch, windows, steps, warmup, barrier, device):
    """warmup untimed steps, then `steps` timed ones: (device ms of the loops, e2e ms, launches,
    launches of the loop, outcomes, clocks).  A step = upload from pinned host buffers, the
    loop of every window to its end, results back into pinned host buffers."""
    import torch
    outcomes = None
    for _ in range(warmup):
        batch.upload(); batch.run(); outcomes = batch.fetch(pinned=True)
    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run_ms, launches, launches_run = 0.0, 0, 0
    ev0.record()
    for _ in range(steps):
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
    return run_ms, ev0.elapsed_time(ev1), launches, launches_run, outcomes, sampler.stop()


def timed_steps_waves(windows, steps, warmup, barrier, device, budget=None):
    """The same for a shard that does not fit the GPU at once: HBM-bounded waves, the next wave
    uploaded on its own stream while this one runs (engine.run_in_waves).  run_ms is the sum of the
    waves' loop times (CUDA events on each wave's stream); e2e is the wall clock between device-wide
    synchronisations (several streams are in flight, one event pair cannot bracket them)."""
    from viloca_b200 import engine
    outcomes, st = None, {}
    for _ in range(warmup):
        outcomes = engine.run_in_waves(windows, device=device, budget_bytes=budget, stats=st)
    sampler = ClockSampler(device)
    sampler.start()
    barrier()
    run_ms, launches, launches_run = 0.0, 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        st = {}
        outcomes = engine.run_in_waves(windows, device=device, budget_bytes=budget, stats=st)
        run_ms += st["device_ms"]
        launches_run += st["launches"]
        launches += st["launches"] + 4 * len(windows) + st["waves"]
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    return run_ms, e2e_ms, launches, launches_run, outcomes, sampler.stop(), st


def d2h_bytes(outcomes, windows):
    n = 0
    for o, w in zip(outcomes, windows):
        n += w.n_clusters * w.length * 5 * 8 + w.n_reads * w.n_clusters * 8 + 2 * w.n_clusters * 8 + 128
        n += sum(len(r.history_elbo) for r in o.restarts) * 8
    return n


def extra_line(name, device, indices=None, steps=2, warmup=1, gen_procs=1, **kw):
    """A short measurement of another BASELINE config on this GPU (reported under `extra`)."""
    import torch
    from viloca_b200 import engine, synthetic
    spec = synthetic.WORKLOADS[name]
    t0 = time.perf_counter()
    windows = [w.pin() for w in synthetic.make_workload(name, indices=indices, procs=gen_procs, **kw)]
    gen_s = time.perf_counter() - t0
    n_problems = sum(w.n_starts for w in windows)
    with engine.CaviBatch(windows, device=device) as batch:
        run_ms, e2e_ms, _, launches_run, outcomes, _ = timed_steps(batch, windows, steps, warmup, torch.cuda.synchronize, device)
        updates = [r.n_updates for o in outcomes for r in o.restarts]
        out = {"workload": name, "windows": len(windows), "of_windows": spec.n_windows, "n_starts": spec.n_starts,
               "n_unique_mean": float(np.mean([w.n_reads for w in windows])), "length_mean": float(np.mean([w.length for w in windows])),
               "value": len(windows) * steps / (run_ms * 1e-3), "unit": UNIT, "ms_per_step": run_ms / steps,
               "window_restarts_per_sec": n_problems * steps / (run_ms * 1e-3),
               "e2e": {"value": len(windows) * steps / (e2e_ms * 1e-3), "ms_per_step": e2e_ms / steps,
                       "h2d_bytes_per_step": batch.input_bytes(), "d2h_bytes_per_step": d2h_bytes(outcomes, windows)},
               "steps": steps, "warmup": warmup, "loop_launches_per_step": launches_run / steps,
               "iterations": {"mean": float(np.mean(updates)), "max": int(max(updates)), "min": int(min(updates))},
               "workspace_mb": batch.workspace_bytes / 1e6, "host_generation_s": gen_s}
    return out


def run_b200_arm(args):
    # host-only work first (the generators fork a pool; nothing has touched CUDA yet)
    from viloca_b200 import synthetic
    rank, world, local_rank = dist_env()
    spec = synthetic.WORKLOADS[args.workload]
    n_windows = args.windows or spec.n_windows
    parts, lpt_how = shard_windows(spec, n_windows, world)
    mine = parts[rank]
    gen_procs = max(1, min(32, (os.cpu_count() or 2) // max(world, 1)))
    t_gen = time.perf_counter()
    windows = synthetic.make_workload(args.workload, indices=mine, procs=gen_procs)
    t_gen = time.perf_counter() - t_gen

    import torch
    import torch.distributed as dist
    from viloca_b200 import _capi as capi, engine
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    torch.cuda.set_device(device)
    windows = [w.pin() for w in windows]
    n_problems = sum(w.n_starts for w in windows)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # does the rank's shard fit the GPU at once?  Otherwise it runs in HBM-bounded waves
    sizes = [engine.window_workspace_bytes(w) for w in windows]
    free_hbm, _ = torch.cuda.mem_get_info(torch.device("cuda", device))
    wave_info = None
    ops = operand_stats(windows)
    if sum(sizes) + (1 << 30) <= free_hbm and not args.waves:
        batch = engine.CaviBatch(windows, device=device)
        workspace_bytes = batch.workspace_bytes
        h2d = batch.input_bytes()
        run_ms, e2e_ms, launches, launches_run, outcomes, clocks = timed_steps(batch, windows, args.steps, args.warmup, barrier, device)
        # per-kernel roofline numbers: a separate pass with every launch bracketed by CUDA events
        ops_per_problem = [o for o, w in zip(ops, windows) for _ in range(w.n_starts)]     # problem = (window, restart)
        prof_tot, prof_first = profiled_run(batch, ops_per_problem)
        prof_scale = 1.0
    else:
        budget = int(args.waves * (1 << 30)) if args.waves else None
        h2d = sum(w.input_bytes() for w in windows)
        workspace_bytes = sum(sizes)
        run_ms, e2e_ms, launches, launches_run, outcomes, clocks, wave_info = timed_steps_waves(
            windows, args.steps, args.warmup, barrier, device, budget)
        # roofline pass on the first wave only (its work scaled to the shard by the iteration counts)
        first = engine.plan_waves(sizes, wave_info["budget_bytes"])[0]
        batch = engine.CaviBatch([windows[i] for i in first], device=device)
        ops_first = [ops[i] for i in first]
        prof_tot, prof_first = profiled_run(batch, [o for o, i in zip(ops_first, first) for _ in range(windows[i].n_starts)])
        work = lambda idx: float(sum(r.n_updates * windows[i].n_reads * windows[i].length for i in idx for r in outcomes[i].restarts))
        prof_scale = work(range(len(windows))) / max(work(first), 1.0)
        for k in prof_tot:
            for key in ("bytes", "executed", "algorithmic", "dense_equivalent", "ms"):
                prof_tot[k][key] *= prof_scale
    d2h = d2h_bytes(outcomes, windows)
    n_updates_window = [sum(r.n_updates for r in o.restarts) for o in outcomes]
    all_updates = [r.n_updates for o in outcomes for r in o.restarts]
    read_iters = sum(r.n_updates * w.n_reads for o, w in zip(outcomes, windows) for r in o.restarts)
    raw_read_iters = sum(r.n_updates * getattr(w, "n_raw", w.n_reads) for o, w in zip(outcomes, windows) for r in o.restarts)

    # (a window counts once, whatever the number of its restarts: the reference's unit of work)
    t = torch.tensor([run_ms, e2e_ms, float(len(windows)), float(read_iters), float(raw_read_iters), float(launches),
                      float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    else:
        tmax = tsum = tmin = t
    run_ms_max, e2e_ms_max = float(tmax[0]), float(tmax[1])
    total_problems = float(tsum[2])

    if args.record_iterations and world == 1:
        os.makedirs(os.path.dirname(ITER_TABLE), exist_ok=True)
        table = {}
        if os.path.exists(ITER_TABLE):
            table = json.load(open(ITER_TABLE))
        table[args.workload] = {"n_updates": all_updates, "n_reads": [w.n_reads for w in windows],
                                "n_raw": [getattr(w, "n_raw", w.n_reads) for w in windows],
                                "length": [w.length for w in windows]}
        json.dump(table, open(ITER_TABLE, "w"))

    if rank == 0:
        # FP64 tensor ceiling of THIS GPU, measured now (DMMA m8n8k4 register loop on every SM)
        peak = C_double()
        fp64_src = "DMMA m8n8k4 register loop measured in this run (vl_debug_fp64_peak)"
        try:
            capi.check(capi.lib.vl_debug_fp64_peak(device, ctypes_byref(peak)))
            fp64_peak = float(peak.value)
        except Exception as exc:                        # noqa: BLE001
            fp64_peak, fp64_src = 37.0, f"fallback 37 TFLOP/s ({exc})"
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

        # The timed runs execute both contractions of every update inside vl_sched_kernel (one
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
            "bound": bound, "kernel": "vl_sched_kernel (haplotype-update + responsibility-update tiles of a chunk of updates)",
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
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(spec, n_windows),
            "parallelism": f"the {n_windows} windows of one genome over {world} GPU(s), longest first; cost = {lpt_how}",
            "workload_stats": {"windows_rank0": len(windows), "n_unique_mean": float(np.mean([w.n_reads for w in windows])),
                               "dense_columns_mean": float(np.mean([o["NC"] for o in ops])),
                               "sparse_entries_mean": float(np.mean([o["nnz"] for o in ops])),
                               "workspace_mb_rank0": workspace_bytes / 1e6, "waves": wave_info, "host_generation_s": t_gen,
                               "rank_ms_per_step": {"max": run_ms_max / args.steps, "min": float(tmin[0]) / args.steps}},
            "e2e": {"value": total_problems * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(float(tsum[6])), "d2h_bytes_per_step": int(float(tsum[7])),
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(float(tsum[5])),
            "window_restarts_per_sec": total_problems * spec.n_starts * args.steps / (run_ms_max * 1e-3),
            "read_iterations_per_sec": float(tsum[3]) * args.steps / (run_ms_max * 1e-3),
            "raw_read_iterations_per_sec": float(tsum[4]) * args.steps / (run_ms_max * 1e-3),
            "iterations": {"mean": float(np.mean(all_updates)), "max": int(max(all_updates)), "min": int(min(all_updates))},
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single_core(spec, windows, n_updates_window, n_windows)
    batch.close()
    batch = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = {}
        torch.cuda.empty_cache()
        for name, kw in EXTRA_WORKLOADS:
            if name == args.workload:
                continue
            try:
                extra[name] = extra_line(name, device, gen_procs=gen_procs, **kw)
            except Exception as exc:                    # noqa: BLE001  (an extra line never costs the headline)
                extra[name] = {"error": repr(exc)}
            torch.cuda.empty_cache()
        line["extra"] = extra
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# other BASELINE configs measured after the headline workload (N = 1 only): configs[1] whole, and a
# bounded shard of configs[4] (2 of its 120 windows x 20 restarts: one window's workspace is 10 GB)
EXTRA_WORKLOADS = (("hiv5k_uqs", {}), ("sars_ont", {"indices": [0, 1], "steps": 1}))


def C_double():
    import ctypes
    return ctypes.c_double(0.0)


def ctypes_byref(x):
    import ctypes
    return ctypes.byref(x)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
