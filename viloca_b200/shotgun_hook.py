"""Batched replacement for ``Pool.map(run_dpm_wrapper, runlist)`` (viloca/shotgun.py:524-538).

``runlist`` is what ``shotgun.win_to_run`` builds (shotgun.py:312-333): 10-tuples
``(filein, j, a, seed, inference_type, n_max_haplotypes, n_mfa_starts, unique_modus,
conv_thres, record_history)``.  All windows are loaded, sharded over the visible GPUs by the
longest-first scheduler and run as one engine batch per GPU (one host thread per GPU; the C ABI
releases the GIL).  Outputs are the per-window files the serial path writes into the CWD.

``run_dpm_batch(..., collect=True)`` also returns the finished state dicts, and ``snv_tables`` turns
them into the per-window SNV tables and co-occurring-mutation rows that ``shorah_snv.getSNV`` /
``get_cooccuring_muts_haplo_df`` (viloca/shorah_snv.py:190-256, 486-526) would otherwise re-parse
from the support files.
"""
from __future__ import annotations

import logging
import threading

import numpy as np
from typing import Optional, Sequence

from . import scheduler
from .local_haplotype_inference import _common

MODES = {"use_quality_scores": "uqs", "learn_error_params": "lep"}


MAX_CLUSTERS = 128     # VL_MAX_CLUSTERS of include/viloca_b200.h (the reference's CLI allows any K, cli.py:179-181)


def _check_clusters(k: int, what: str) -> None:
    if int(k) > MAX_CLUSTERS:
        raise ValueError(f"{what}: n_max_haplotypes = {k} exceeds the engine's limit of {MAX_CLUSTERS} clusters "
                         f"(VL_MAX_CLUSTERS); run this window with the reference, or with K <= {MAX_CLUSTERS}")


def jobs_from_runlist(runlist) -> list:
    jobs = []
    for run_setting in runlist:
        (filein, _j, a, seed, inference_type, n_max_haplotypes, n_mfa_starts, unique_modus,
         conv_thres, record_history) = run_setting
        if inference_type not in MODES:
            raise ValueError(f"inference_type {inference_type!r} is not served by the CUDA engine")
        _check_clusters(n_max_haplotypes, filein)
        mode = MODES[inference_type]
        stem = filein.split(".reads.")[0]                       # shotgun.py:151-152
        seed = int(seed[0]) if hasattr(seed, "__len__") else int(seed)   # shotgun.py:442-443 passes an array
        jobs.append(_common.WindowJob(
            mode=mode, freads_in=filein, fref_in=stem + ".ref.fas",
            fname_qualities=(stem + ".qualities.npy") if mode == "uqs" else None,
            output_dir="./", n_starts=int(n_mfa_starts), K=int(n_max_haplotypes), alpha0=float(a),
            unique_modus=unique_modus, record_history=record_history, seed=seed,
            convergence_threshold=conv_thres if mode == "uqs" else 1e-3))
    return jobs


def window_coordinates(job) -> tuple:
    """(chrom, beg, end) from the window file name ``w-<chrom>-<beg>-<end>.reads.fas`` (b2w.py:604-616)."""
    stem = job.freads_in.split("/")[-1].split(".reads.")[0]
    chrom, beg, end = stem[2:].rsplit("-", 2)
    return chrom, beg, end


def snv_tables(jobs, finished, threshold: float = 0.9) -> list:
    """Per window: {"window": (chrom, beg, end), "snvs": {SNP_id: SNV}, "cooccurring": [rows]} from the
    in-memory results of :func:`run_dpm_batch` (``collect=True``)."""
    from . import snv_windows as sw
    out = []
    for job, (state, _result) in zip(jobs, finished):
        chrom, beg, end = window_coordinates(job)
        haps = sw.haplotypes_from_state(state)
        out.append({"window": (chrom, beg, end),
                    "snvs": sw.window_snvs(haps, job.ref_seq, chrom, beg, end, threshold),
                    "cooccurring": sw.cooccurring_mutations(haps, job.ref_seq, chrom, beg, end)})
    return out


def jobs_from_arrays(windows, inference_type: str, n_max_haplotypes: int, n_mfa_starts: int, alpha0: float, seed: int = 27,
                     unique_modus: bool = True, conv_thres: float = 1e-3, write_files: bool = False,
                     output_dir: str = "./") -> list:
    """In-memory hand-over from the window builder: what b2w.py:604-616, 670-672 writes per window and
    shotgun.py:312-333 reads back, passed as arrays instead.  Each window is a dict (or object) with
    ``name`` (``w-<chrom>-<beg>-<end>``, the stem of the files the reference would write), ``read_ids`` and
    ``read_seqs`` (equal-length strings over ACGT-N, in b2w's order; or ``reads_text``, the FASTA text b2w would
    have written, as bytes), ``qualities`` ((n_raw, L) integer
    array; ignored for learn_error_params) and ``ref`` ((id, sequence) of the reference window).  The other
    arguments are the fields of the reference's run-setting tuple (shotgun.py:326).  ``write_files=False``
    keeps the results in memory (``run_dpm_batch(..., collect=True)`` + ``snv_tables``)."""
    if inference_type not in MODES:
        raise ValueError(f"inference_type {inference_type!r} is not served by the CUDA engine")
    mode = MODES[inference_type]
    _check_clusters(n_max_haplotypes, "jobs_from_arrays")
    jobs = []
    for w in windows:
        get = (lambda k: w[k]) if isinstance(w, dict) else (lambda k: getattr(w, k))
        name = get("name")
        q = None
        if mode == "uqs":
            q = np.ascontiguousarray(get("qualities"))
            if q.dtype not in (np.float64, np.int64):
                q = q.astype(np.int64)                                   # b2w saves int64 (b2w.py:508)
        try:
            text = get("reads_text")                                     # the FASTA text b2w would write, if the caller has it
        except (KeyError, AttributeError):
            text = None
        if text is None:
            text = _common.fasta_text(get("read_ids"), get("read_seqs"))
        jobs.append(_common.WindowJob(
            mode=mode, freads_in=name + ".reads.fas", fref_in=name + ".ref.fas",
            fname_qualities=(name + ".qualities.npy") if mode == "uqs" else None,
            output_dir=output_dir, n_starts=int(n_mfa_starts), K=int(n_max_haplotypes), alpha0=float(alpha0),
            unique_modus=unique_modus, record_history=False, seed=int(seed),
            convergence_threshold=conv_thres if mode == "uqs" else 1e-3,
            reads_text=text, qualities=q,
            ref_record=tuple(get("ref")), write_files=write_files))
    return jobs


def run_dpm_batch(runlist, devices: Optional[Sequence[int]] = None, collect: bool = False,
                  load_workers: Optional[int] = None):
    """``runlist``: the reference's run-setting tuples (files on disk), or WindowJobs from
    :func:`jobs_from_arrays` (windows in memory)."""
    import torch
    runlist = list(runlist)
    todo = runlist if runlist and isinstance(runlist[0], _common.WindowJob) else jobs_from_runlist(runlist)
    if not todo:
        return ([], []) if collect else None
    # windows that fit host memory together are loaded at once (a process pool, like the reference's); a
    # genome that does not (ONT: 1.6 GB of initial states per window) is loaded group by group inside run_jobs
    sizes = [_common.host_bytes_estimate(j) for j in todo]
    lazy = sum(sizes) > _common.host_budget_bytes()
    jobs = todo if lazy else _common.load_jobs(todo, workers=load_workers)
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if not devices:
        raise RuntimeError("no CUDA device: the CAVI engine has no CPU fallback")
    parts = scheduler.lpt_partition([float(sz) if lazy else j.window.cost() for j, sz in zip(jobs, sizes)], len(devices))
    errors = []
    finished = [None] * len(jobs)
    host_share = max(1, _common.host_budget_bytes() // len(devices))     # the device threads load side by side

    def work(dev, idx):
        try:
            if idx:
                for i, fin in zip(idx, _common.run_jobs([jobs[i] for i in idx], device=dev, host_budget=host_share)):
                    finished[i] = fin
        except Exception as exc:                                 # shotgun.py:137-143: log and re-raise
            logging.error(f"Error in process for device {dev}: {exc}")
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(dev, idx)) for dev, idx in zip(devices, parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return (jobs, finished) if collect else None
