"""Shared host logic of the two inference modes: window files -> compact tensors, the batched
engine call, and the finishing steps that turn a converged state into the two FASTA files the
rest of VILOCA parses (use_quality_scores/run_dpm_mfa.py:18-120,
learn_error_params/run_dpm_mfa.py:33-110).

Nothing here computes CAVI arithmetic on the CPU: the update equations, the ELBO and the
finishing contraction all run in the CUDA engine (viloca_b200.engine).
"""
from __future__ import annotations

import logging
import os
import pickle
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from .. import fasta, init_state, tensors
from .._consts import INIT_STRIDE

ALPHABET = "ACGT-"


# ---------------------------------------------------------------------------------------
# window files -> tensors
# ---------------------------------------------------------------------------------------

@dataclass
class PreparedWindow:
    """The reads of one window after the reference's preparation step, in compact form."""
    codes: np.ndarray                  # (N, L) uint8
    weights: np.ndarray                # (N,) multiplicities (int64 for unique, float ones otherwise)
    qualities: Optional[np.ndarray]    # (N, L) float64 (uqs) or None (lep)
    read_ids: list                     # per row: ids of the raw reads collapsed into it, first-seen order
    sequences: list                    # per row: the read string (keys of read_mapping)
    n_raw: int = 0


def read_window_fasta(path: str):
    ids, seqs = [], []
    for rec in fasta.parse(path):
        ids.append(rec.id)
        seqs.append(rec.seq)
    return ids, seqs


def prepare_reads(ids: Sequence[str], seqs: Sequence[str], qualities: Optional[np.ndarray],
                  alphabet: str, unique_modus: bool) -> PreparedWindow:
    """use_quality_scores/preparation.py:24-90 and learn_error_params/preparation.py:51-94.

    unique: identical read strings collapse into the row of their first occurrence; weight =
    multiplicity; quality = running mean (q*w + q_new)/(w + 1) in file order; afterwards
    Q == 0 -> 2.  non-unique: one row per read, weight 1, Q == 0 -> 2.
    The O(N^2) Hamming scan of the lep loader gives the same rows in the same order.
    """
    n_raw = len(seqs)
    if n_raw == 0:
        raise ValueError("window without reads")
    length = len(seqs[0])
    if unique_modus:
        row_of = {}
        rows, read_ids, weights = [], [], []
        qual_rows = []
        for i, s in enumerate(seqs):
            r = row_of.get(s)
            if r is None:
                row_of[s] = len(rows)
                rows.append(s)
                read_ids.append([ids[i]])
                weights.append(1)
                if qualities is not None:
                    qual_rows.append(qualities[i])
            else:
                if qualities is not None:
                    qual_rows[r] = (qual_rows[r] * weights[r] + qualities[i]) / (weights[r] + 1)
                weights[r] += 1
                read_ids[r].append(ids[i])
        w = np.asarray(weights, dtype=np.int64)
    else:
        rows = list(seqs)
        read_ids = [[i] for i in ids]
        w = np.ones(n_raw)
        qual_rows = [qualities[i] for i in range(n_raw)] if qualities is not None else []
    codes = np.empty((len(rows), length), dtype=np.uint8)
    for r, s in enumerate(rows):
        if len(s) != length:
            raise ValueError("reads of one window must have equal length")
        codes[r] = tensors.encode(s, alphabet)
    q = None
    if qualities is not None:
        q = np.empty((len(rows), np.asarray(qualities).shape[1]), dtype=np.float64)
        for r, row in enumerate(qual_rows):
            q[r] = row
        q = np.where(q == 0, 2, q)
    return PreparedWindow(codes=codes, weights=w, qualities=q, read_ids=read_ids, sequences=rows, n_raw=n_raw)


def prepare_reads_file(path: str, qualities: Optional[np.ndarray], alphabet: str, unique_modus: bool) -> PreparedWindow:
    """The window's read file through the native reader (``vl_reads_parse`` / ``vl_reads_fill``): the same
    PreparedWindow as ``prepare_reads(*read_window_fasta(path), ...)`` bit for bit, without the per-record
    Python loops.  Files the native reader does not take (multi-line sequences, qualities that are not
    a dense (n_raw, length) int64 / float64 array, non-ASCII alphabets) go through the generic path."""
    with open(path, "rb") as fh:
        text = fh.read()
    return prepare_reads_text(text, qualities, alphabet, unique_modus, path)


def fasta_text(ids: Sequence[str], seqs: Sequence[str]) -> bytes:
    """The window's reads as b2w writes them (b2w.py:613-616: one ``>id`` line and one sequence line per
    read), in memory."""
    return "".join(f">{i}\n{s}\n" for i, s in zip(ids, seqs)).encode("ascii")


def prepare_reads_text(text: bytes, qualities: Optional[np.ndarray], alphabet: str, unique_modus: bool,
                       path: Optional[str] = None) -> PreparedWindow:
    """prepare_reads_file on the FASTA text itself (a file's bytes, or a window handed over in memory)."""
    import ctypes as C
    from .. import _capi as capi
    q_kind, q = 0, None
    if qualities is not None:
        q = np.asarray(qualities)
        if q.dtype == np.int64 and q.ndim == 2:
            q_kind = 1
        elif q.dtype == np.float64 and q.ndim == 2:
            q_kind = 2
        else:
            q_kind = -1
    ok = q_kind >= 0 and len(alphabet) < 250 and all(ord(c) < 128 for c in alphabet)
    handle, n_raw, n_rows, length = C.c_void_p(), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    if ok and capi.lib.vl_reads_parse(text, len(text), 1 if unique_modus else 0, C.byref(handle), C.byref(n_raw),
                                      C.byref(n_rows), C.byref(length)) != 0:
        ok = False
        if "equal length" in capi.last_error() or "without reads" in capi.last_error():
            raise ValueError(capi.last_error())
    if ok and q is not None and q.shape != (n_raw.value, length.value):
        capi.lib.vl_reads_free(handle)
        ok = False
    if not ok:
        if path is None:
            import io
            recs = list(fasta.parse(io.StringIO(text.decode("ascii", "replace"))))
            return prepare_reads([r.id for r in recs], [r.seq for r in recs], qualities, alphabet, unique_modus)
        return prepare_reads(*read_window_fasta(path), qualities, alphabet, unique_modus)
    try:
        nr, nu, ln = n_raw.value, n_rows.value, length.value
        codes = np.empty((nu, ln), dtype=np.uint8)
        weights = np.empty(nu, dtype=np.float64)
        mean_q = np.empty((nu, ln), dtype=np.float64) if q is not None else None
        row_of = np.empty(nr, dtype=np.int32)
        id_spans = np.empty((nr, 2), dtype=np.int64)
        seq_spans = np.empty((nu, 2), dtype=np.int64)
        qc = np.ascontiguousarray(q) if q is not None else None
        capi.check(capi.lib.vl_reads_fill(
            handle, alphabet.encode("ascii"), qc.ctypes.data_as(C.c_void_p) if qc is not None else None, max(q_kind, 0),
            capi.u8_ptr(codes), capi.f64_ptr(weights), capi.f64_ptr(mean_q), capi.i32_ptr(row_of),
            id_spans.ctypes.data_as(C.POINTER(C.c_int64)), seq_spans.ctypes.data_as(C.POINTER(C.c_int64))))
    finally:
        capi.lib.vl_reads_free(handle)
    ids = [text[a:a + n].decode("ascii", "replace") for a, n in id_spans.tolist()]
    read_ids = [[] for _ in range(nu)]
    for i, r in enumerate(row_of.tolist()):
        read_ids[r].append(ids[i])
    rows = [text[a:a + n].decode("ascii", "replace") for a, n in seq_spans.tolist()]
    w = weights.astype(np.int64) if unique_modus else weights
    return PreparedWindow(codes=codes, weights=w, qualities=mean_q, read_ids=read_ids, sequences=rows, n_raw=nr)


def load_reference(path: str, alphabet: str, upper: bool):
    """First record of the reference FASTA -> (codes, id).  uqs upper-cases the sequence
    (use_quality_scores/preparation.py:163-165); lep does not, so lower-case bases are outside
    its alphabet (learn_error_params/preparation.py:97-108)."""
    for rec in fasta.parse(path):
        seq = rec.seq.upper() if upper else rec.seq
        return tensors.encode(seq, alphabet), rec.id, rec.seq
    raise ValueError(f"{path}: no sequence")


# ---------------------------------------------------------------------------------------
# running a batch of windows on the engine
# ---------------------------------------------------------------------------------------

@dataclass
class WindowJob:
    """One call of run_dpm_mfa.main, as data."""
    mode: str
    freads_in: str
    fref_in: str
    fname_qualities: Optional[str]
    output_dir: str
    n_starts: int
    K: int
    alpha0: float
    alphabet: str = ALPHABET
    unique_modus: bool = True
    convergence_threshold: float = 1e-3
    record_history: bool = False
    seed: int = 27
    max_iter: int = 50000
    # in-memory hand-over from the window builder (instead of the three files of b2w.py:613-616, 670-672):
    # reads as FASTA text or as (ids, sequences), the (n_raw, L) quality array, the reference window
    reads_text: Optional[bytes] = None
    qualities: Optional[np.ndarray] = None
    ref_record: Optional[tuple] = None          # (id, sequence)
    write_files: bool = True                    # False: results stay in memory (shotgun_hook.snv_tables)
    # filled by load()
    prepared: Optional[PreparedWindow] = None
    ref_codes: Optional[np.ndarray] = None
    ref_seq: str = ""
    window: object = None

    @property
    def output_name(self) -> str:
        window_id = self.freads_in.split("/")[-1][:-4]      # run_dpm_mfa.py:33
        return self.output_dir + window_id + "-"

    def load(self):
        from ..window import Window
        if self.alphabet != ALPHABET:
            raise ValueError("the engine is built for the alphabet 'ACGT-' (B = 5)")
        quals = None
        if self.reads_text is not None:
            quals = self.qualities if self.mode == "uqs" else None
            self.prepared = prepare_reads_text(self.reads_text, quals, self.alphabet, self.unique_modus)
            _rid, rseq = self.ref_record
            seq = rseq.upper() if self.mode == "uqs" else rseq          # as load_reference
            self.ref_codes, self.ref_seq = tensors.encode(seq, self.alphabet), rseq
        else:
            if self.mode == "uqs":
                with open(self.fname_qualities, "rb") as fh:
                    quals = np.load(fh, allow_pickle=True)
            self.prepared = prepare_reads_file(self.freads_in, quals, self.alphabet, self.unique_modus)
            self.ref_codes, _, self.ref_seq = load_reference(self.fref_in, self.alphabet, upper=(self.mode == "uqs"))
        p = self.prepared
        n = p.codes.shape[0]
        zs, scs = init_state.draw_restarts(self.mode, self.K, self.alpha0, n, self.n_starts, self.seed)
        kw = {}
        if self.mode == "uqs":
            kw["log_hit"], kw["log_off"] = tensors.log_error_terms(p.qualities)
        self.window = Window(mode=self.mode, codes=p.codes, weights=np.asarray(p.weights, dtype=np.float64),
                             ref_codes=self.ref_codes, init_cluster=zs, init_scalars=scs, alpha0=self.alpha0,
                             conv_thres=self.convergence_threshold, max_iter=self.max_iter,
                             name=self.output_name, **kw)
        return self


def _load_one(job: WindowJob) -> WindowJob:
    return job.load()


def _load_cost_seconds(job: WindowJob) -> float:
    """Rough host time of WindowJob.load: parsing / dedup ~ 12 ns per byte of the read file, the seeded
    Dirichlet draws ~ 120 ns per (read, cluster, restart)."""
    try:
        size = len(job.reads_text) if job.reads_text is not None else os.path.getsize(job.freads_in)
    except OSError:
        return 0.0
    return 12e-9 * size + 1.2e-7 * (size / 200.0) * job.K * job.n_starts


def host_bytes_estimate(job: WindowJob) -> int:
    """Host memory a loaded window takes, from the size of its read file alone (nothing is parsed):
    codes + qualities + the two log tables of the unique reads (<= raw reads) and the seeded initial
    responsibilities (reads x K x restarts doubles) -- the term that makes a 100 000-read, 20-restart
    window 1.6 GB."""
    if job.window is not None:
        w = job.window
        return int(w.input_bytes()) + (0 if job.prepared is None or job.prepared.qualities is None else job.prepared.qualities.nbytes)
    try:
        if job.reads_text is not None:
            size, head = len(job.reads_text), job.reads_text[:65536]
        else:
            size = os.path.getsize(job.freads_in)
            with open(job.freads_in, "rb") as fh:
                head = fh.read(65536)
    except OSError:
        return 0
    lines = head.split(b"\n")
    rec = (len(lines[0]) + len(lines[1]) + 2) if len(lines) > 1 else 1
    length = len(lines[1]) if len(lines) > 1 else 1
    n_raw = max(1, size // max(rec, 1))
    return int(n_raw * (length * (1 + 8 + 16) + 8.0 * job.K * job.n_starts))


def host_budget_bytes() -> int:
    """Host memory the loaded windows of one run_jobs call may take together: 40 % of what is available."""
    try:
        import psutil
        return int(0.4 * psutil.virtual_memory().available)
    except Exception:                      # noqa: BLE001
        return 64 << 30


def release_job(job: WindowJob) -> None:
    """Drop the large host arrays of a finished window (inputs, initial states); what the caller and
    ``snv_tables`` need afterwards -- names, the reference window, the finished state -- stays."""
    job.window = None
    job.qualities = None
    job.reads_text = None
    if job.prepared is not None:
        job.prepared.qualities = None


def load_jobs(jobs: Sequence[WindowJob], workers: Optional[int] = None) -> list:
    """Read and prepare the windows of ``jobs`` (FASTA / qualities -> compact tensors, seeded initial
    states).  This is host work of the same order as the GPU run itself, so it is spread over the host
    cores the way the reference spreads its windows (shotgun.py:527-538: ``Pool(max_proc)``):
    ``workers`` = None picks by estimated serial time -- a pool of cpu_count() - 1 spawned processes
    above a few seconds (large windows: N x K Dirichlet draws per restart), otherwise a handful of
    threads (the native reader, NumPy's loops and file reads release the GIL; 20 -> 7 ms per
    5 000-read window on 8 cores); 0 or 1: in this process; n > 1: n processes.  Workers never touch
    CUDA (spawned, not forked: the parent may hold CUDA and threads); results are identical to
    loading serially."""
    jobs = list(jobs)
    todo = [j for j in jobs if j.window is None]
    cores = os.cpu_count() or 2
    if workers is None:
        if sum(_load_cost_seconds(j) for j in todo) > 8.0:
            workers = max(1, cores - 1)
        else:
            if len(todo) >= 4 and cores > 2:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(min(8, cores - 1)) as ex:
                    list(ex.map(_load_one, todo))
                return jobs
            workers = 1
    workers = min(workers, len(todo))
    if workers <= 1:
        for j in todo:
            j.load()
        return jobs
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(workers) as pool:
        loaded = pool.map(_load_one, todo, chunksize=max(1, len(todo) // (4 * workers)))
    it = iter(loaded)
    return [j if j.window is not None else next(it) for j in jobs]


def _result_dicts(job: WindowJob, outcome, restart: int, state: dict):
    """(state_curr_dict, dict_result) in the shape run_cavi returns (cavi.py:171-210)."""
    r = outcome.restarts[restart]
    st = dict(state)
    st["elbo"] = r.elbo
    res = {"exit_message": r.exit_message, "n_iterations": r.n_iterations, "converged": r.converged,
           "elbo": r.elbo, "history_elbo": list(r.history_elbo), "run_id": restart,
           "n_reads": job.window.n_reads, "n_cluster": job.K, "alpha0": job.alpha0, "alphabet": job.alphabet}
    res.update(st)
    return st, res


def group_haplotypes(mean_haplo: np.ndarray, alphabet: str = ALPHABET):
    """get_unique_haplotypes (analyze_results.py:102-107, 144-157): per-position argmax (first
    maximum wins), identical strings grouped, groups sorted lexicographically."""
    idx = np.argmax(mean_haplo, axis=2)
    letters = np.frombuffer(alphabet.encode("ascii"), dtype=np.uint8)[idx]       # (K, L) bytes
    groups = {}
    for k in range(idx.shape[0]):
        groups.setdefault(letters[k].tobytes().decode("ascii"), []).append(k)
    return sorted(groups.items())


def summarize(job: WindowJob, batch, window_index: int, restart: int, state: dict) -> dict:
    """summarize_results (analyze_results.py:56-99): merged clusters, one more haplotype
    update on them (on the device), posterior product, tie-aware read assignment."""
    groups = group_haplotypes(state["mean_haplo"], job.alphabet)
    group_of = np.empty(job.K, dtype=np.int32)
    for j, (_, members) in enumerate(groups):
        group_of[members] = j
    merged_z, merged_h = batch.merge_haplo(window_index, restart, group_of, len(groups))
    p = job.prepared
    top = merged_z.max(axis=1, keepdims=True)
    member = merged_z == top                       # (N, G): every cluster attaining the row maximum
    mult = np.asarray(p.weights)
    summary = {}
    for j, (seq, _) in enumerate(groups):
        base_idx = tensors.encode(seq, job.alphabet).astype(np.int64)
        vals = merged_h[j, np.arange(len(seq)), base_idx]
        posterior = np.cumprod(vals)[-1] if len(vals) else 1     # left-to-right product
        rows = np.nonzero(member[:, j])[0]
        assigned = [rid for n in rows for rid in p.read_ids[n]]
        # uqs lists every raw read in both keys (analyze_results.py:121-133); lep's unique list holds
        # only the representative id of each row (learn_error_params/analyze_results.py:236-243)
        assigned_unique = [p.read_ids[n][0] for n in rows] if job.mode == "lep" else list(assigned)
        weight = mult[rows].sum()
        weight = int(weight) if job.unique_modus else int(round(float(weight)))
        summary.update({
            "haplotype" + str(j): seq,
            "approximatePosterior" + str(j): posterior,
            "assignedUniqueReads" + str(j): assigned_unique,
            "assignedReads" + str(j): list(assigned),
            "weight" + str(j): weight,
        })
    return summary


def haplotypes_to_fasta(state: dict, path: str) -> None:
    """analyze_results.py:26-53.  Haplotypes without supporting reads are skipped; if none has
    support the file is not created (the reference only writes inside the loop)."""
    n_hap = len([k for k in state if k.startswith("haplotype")])
    records, wrote = [], False
    for k in range(n_hap):
        ave_reads = state["weight" + str(k)]
        if ave_reads == 0:
            continue
        head = "hap_" + str(k) + " | posterior=" + str(state["approximatePosterior" + str(k)]) + \
               " ave_reads=" + str(ave_reads)
        records.append(("haplotype" + str(k), head, state["haplotype" + str(k)]))
        wrote = True
    if wrote:
        fasta.write(records, path)


def correct_reads(state: dict, path: str) -> None:
    """analyze_results.py:9-23."""
    n_hap = len([k for k in state if k.startswith("haplotype")])
    if n_hap == 0:
        return
    # every read of a haplotype gets the same sequence lines: format them once per haplotype
    parts = []
    for k in range(n_hap):
        body = fasta.format_body(state["haplotype" + str(k)])
        tail = " |posterior=1\n" + body                 # fasta.format_title(id, "|posterior=1") for any other id
        parts.extend(">" + read_id + tail if read_id != "|posterior=1" else ">|posterior=1\n" + body
                     for read_id in state["assignedReads" + str(k)])
    with open(path, "w") as fh:
        fh.write("".join(parts))


def _record_history_run(job: WindowJob, batch, window_index: int):
    """record_history=True: the reference keeps alpha / mean_log_pi / mean_log_gamma /
    mean_cluster after every even iteration (cavi.py:107-112, 146-150).  Here the batch is
    stepped one update at a time and the state copied back each time (slow path, only on
    request)."""
    hist = {r: {"history_alpha": [], "history_mean_log_pi": [], "history_mean_log_gamma": [],
                "history_mean_cluster": []} for r in range(job.n_starts)}
    w = job.window
    for r in range(job.n_starts):
        sc = w.init_scalars[r]
        hist[r]["history_alpha"].append(w.alpha0 * np.ones(w.n_clusters))
        from scipy.special import digamma          # the initial entry of cavi.py:108: psi(alpha0) - psi(K alpha0)
        hist[r]["history_mean_log_pi"].append(digamma(w.alpha0 * np.ones(w.n_clusters)) - digamma(w.n_clusters * w.alpha0))
        hist[r]["history_mean_log_gamma"].append((sc[2], sc[3]))
        hist[r]["history_mean_cluster"].append(w.init_cluster[r].copy())
    active = set(range(job.n_starts))
    it = 0
    while active and it < job.max_iter:
        batch.step(1)
        for r in sorted(active):
            sel = [r if i == window_index else "none" for i in range(len(batch.windows))]
            out = batch.fetch(state=sel, want_haplo=False, want_history=False)[window_index]
            if it % 2 == 0:
                hist[r]["history_mean_log_pi"].append(out.state["mean_log_pi"].copy())
                hist[r]["history_mean_log_gamma"].append(out.state["mean_log_gamma"])
                hist[r]["history_mean_cluster"].append(out.state["mean_cluster"].copy())
            if out.restarts[r].status != 0:
                active.discard(r)
        it += 1
    return hist


def _finish_job(job: WindowJob, batch, i: int, out, histories=None):
    """Best restart of one finished window -> state / result dicts, summary, the two output files."""
    best = out.best_restart
    logging.info("reference " + job.fref_in)
    logging.info("reads " + job.freads_in)
    logging.info("Maximal ELBO " + str(out.restarts[best].elbo) + "in run " + str(best))
    logging.info("CAVI termination " + str(out.restarts[best].exit_message))
    state, result = _result_dicts(job, out, best, out.state)
    if job.record_history:
        order = sorted(range(job.n_starts), key=lambda r: out.restarts[r].elbo, reverse=True)
        per = []
        for r in order:
            o_r = batch.fetch(state=[r if x == i else "none" for x in range(len(batch.windows))])[i]
            s_r, d_r = _result_dicts(job, o_r, r, o_r.state)
            d_r.update((histories or {}).get(i, {}).get(r, {}))
            per.append((s_r, d_r))
        with open(job.output_name + "all_results.pkl", "wb") as fh:
            pickle.dump(per, fh)
        logging.info("Results dicts of all runs written to " + job.output_name + "all_results.pkl")
    state.update(summarize(job, batch, i, best, state))
    if job.write_files:
        haplotypes_to_fasta(state, job.output_name + "support.fas")
        correct_reads(state, job.output_name + "cor.fas")
    return state, result


def run_jobs(jobs: Sequence[WindowJob], device: int = 0, budget_bytes: Optional[int] = None,
             host_budget: Optional[int] = None) -> list:
    """Load, run and finish a list of windows on one GPU; writes ``<output_name>support.fas``
    and ``<output_name>cor.fas`` per window (run_dpm_mfa.py:119-120).  Returns the best
    restart's state dict (with the summary merged in) per job.  The windows go through the GPU in
    waves bounded by the free HBM (``engine.run_in_waves``: the next wave is uploaded while this
    one runs), the way the reference bounds its pool by ``max_proc`` (shotgun.py:527-538)."""
    from ..engine import CaviBatch, plan_waves, run_in_waves
    jobs = list(jobs)
    for job in jobs:
        if job.write_files:
            os.makedirs(job.output_dir, exist_ok=True)
    # Host memory is bounded too: when the windows would not fit together (an ONT genome: 120 windows x
    # 1.6 GB of initial states), they are loaded, run and released group by group
    sizes = [host_bytes_estimate(j) for j in jobs]
    budget = host_budget_bytes() if host_budget is None else int(host_budget)
    if len(jobs) > 1 and sum(sizes) > budget:
        finished = [None] * len(jobs)
        for group in plan_waves(sizes, budget):
            part = load_jobs([jobs[i] for i in group])
            for i, job, fin in zip(group, part, run_jobs(part, device=device, budget_bytes=budget_bytes, host_budget=1 << 62)):
                finished[i] = fin
                jobs[i].ref_seq, jobs[i].prepared = job.ref_seq, job.prepared      # (a process pool returns copies)
                release_job(job)
                release_job(jobs[i])
        return finished
    jobs = load_jobs(jobs)
    finished = [None] * len(jobs)
    plain = [i for i, j in enumerate(jobs) if not j.record_history]
    if plain:
        def consume(batch, idx, outcomes):
            for local, (k, out) in enumerate(zip(idx, outcomes)):
                finished[plain[k]] = _finish_job(jobs[plain[k]], batch, local, out)
        run_in_waves([jobs[i].window for i in plain], device=device, budget_bytes=budget_bytes, consume=consume,
                     fetch_options={"state": "best"})
    slow = [i for i, j in enumerate(jobs) if j.record_history]
    if slow:
        # record_history: single-step so that per-iteration states can be copied out
        with CaviBatch([jobs[i].window for i in slow], device=device) as batch:
            batch.upload()
            histories = {local: _record_history_run(jobs[i], batch, local) for local, i in enumerate(slow)}
            batch.run()
            outcomes = batch.fetch(state="best")
            for local, (i, out) in enumerate(zip(slow, outcomes)):
                finished[i] = _finish_job(jobs[i], batch, local, out, histories)
    return finished
