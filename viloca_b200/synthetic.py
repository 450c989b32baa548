"""Seeded synthetic windows of the shapes BASELINE.json names.

Window tensors are generated directly (not through BAM): H haplotypes at a small divergence
from a random reference, reads sampled from them with Phred-driven substitution errors,
optional deletions ('-') and N-padding, then collapsed with the reference's dedup rule
(first-seen order, running-mean qualities, Q == 0 -> 2; use_quality_scores/preparation.py:48-90).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import init_state, tensors
from .window import Window

GAP = 4          # code of '-'
CODE_N = tensors.CODE_N


@dataclass
class WorkloadSpec:
    name: str
    mode: str                 # "uqs" | "lep"
    genome_length: int
    window_length: int        # L (or the centre of the per-window length range)
    window_step: int
    n_windows: int
    coverage: int             # N_raw per window
    n_haplotypes: int
    q_mean: float
    q_sd: float
    q_lo: int = 2
    q_hi: int = 41
    called_length: int | None = None   # called bases per row, rest 'N' (None = full window)
    deletion_rate: float = 0.0
    length_jitter: int = 0             # amplicon mode: L_w ~ U{L - j .. L + j}
    n_clusters: int = 100
    n_starts: int = 1
    alpha0: float = 1e-4
    conv_thres: float = 1e-3
    seed_init: int = 27


# BASELINE.json configs[1..4]; configs[0] is the reference's real fixture.
WORKLOADS = {
    "hiv5k_uqs": WorkloadSpec("hiv5k_uqs", "uqs", 9719, 201, 67, 145, 5000, 5, 36.0, 4.0, called_length=150),
    "hiv5k_lep": WorkloadSpec("hiv5k_lep", "lep", 9719, 201, 67, 145, 5000, 5, 36.0, 4.0, called_length=150),
    "sars_amplicon": WorkloadSpec("sars_amplicon", "uqs", 29903, 400, 300, 98, 20000, 10, 36.0, 4.0,
                                  length_jitter=20),
    "sars_ont": WorkloadSpec("sars_ont", "uqs", 29903, 1000, 250, 120, 100000, 10, 18.0, 5.0, q_hi=40,
                             deletion_rate=0.05, n_starts=20),
}


def make_haplotypes(rng, length: int, n_haplotypes: int, divergence=(0.01, 0.03)):
    """Random reference and H haplotypes at 1-3 % substitutions from it."""
    ref = rng.integers(0, 4, size=length).astype(np.uint8)
    haps = np.tile(ref, (n_haplotypes, 1))
    for h in range(n_haplotypes):
        rate = rng.uniform(*divergence)
        pos = np.nonzero(rng.random(length) < rate)[0]
        haps[h, pos] = (ref[pos] + rng.integers(1, 4, size=pos.size)) % 4
    return ref, haps


def sample_reads(rng, haps, freqs, n_raw: int, spec: WorkloadSpec):
    """Raw (N_raw, L) base codes and integer Phred qualities."""
    n_h, length = haps.shape
    src = rng.choice(n_h, size=n_raw, p=freqs)
    codes = haps[src].copy()
    quals = np.clip(np.rint(rng.normal(spec.q_mean, spec.q_sd, size=(n_raw, length))), spec.q_lo, spec.q_hi).astype(np.int64)
    err = rng.random((n_raw, length)) < 10.0 ** (-quals / 10.0)
    shift = rng.integers(1, 4, size=(n_raw, length)).astype(np.uint8)
    codes = np.where(err, (codes + shift) % 4, codes).astype(np.uint8)
    if spec.deletion_rate > 0:
        codes[rng.random((n_raw, length)) < spec.deletion_rate] = GAP
    if spec.called_length is not None and spec.called_length < length:
        start = rng.integers(0, length - spec.called_length + 1, size=n_raw)
        col = np.arange(length)[None, :]
        outside = (col < start[:, None]) | (col >= (start + spec.called_length)[:, None])
        codes[outside] = CODE_N
    return codes, quals


def dedup_reads(codes_raw: np.ndarray, quals_raw: np.ndarray):
    """The reference's unique_modus collapse: first-seen order, multiplicity weights,
    running-mean qualities (q*w + q_new)/(w + 1), then Q == 0 -> 2."""
    uniq, first, inverse = np.unique(codes_raw, axis=0, return_index=True, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    order = np.argsort(first, kind="stable")          # unique rows in first-seen order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    row_of = rank[inverse]                            # raw read -> unique row
    n_u = order.size
    weights = np.zeros(n_u, dtype=np.int64)
    quals = np.zeros((n_u, codes_raw.shape[1]), dtype=np.float64)
    for i in range(codes_raw.shape[0]):
        r = row_of[i]
        if weights[r] == 0:
            quals[r] = quals_raw[i]
            weights[r] = 1
        else:
            quals[r] = (quals[r] * weights[r] + quals_raw[i]) / (weights[r] + 1)
            weights[r] += 1
    quals = np.where(quals == 0, 2, quals)
    return codes_raw[first[order]], weights, quals, row_of


def make_window(spec: WorkloadSpec, index: int, master_seed: int = 0, n_reads: int | None = None,
                length: int | None = None, unique: bool = True, max_iter: int = 50000) -> Window:
    """Window ``index`` of a workload: data from default_rng([master_seed, index]), initial
    states from default_rng(seed_init + s)."""
    rng = np.random.default_rng([master_seed, index])
    L = length if length is not None else spec.window_length
    if length is None and spec.length_jitter:
        L = int(rng.integers(spec.window_length - spec.length_jitter, spec.window_length + spec.length_jitter + 1))
    n_raw = n_reads if n_reads is not None else spec.coverage
    ref, haps = make_haplotypes(rng, L, spec.n_haplotypes)
    freqs = 0.75 ** np.arange(spec.n_haplotypes)
    freqs /= freqs.sum()
    codes_raw, quals_raw = sample_reads(rng, haps, freqs, n_raw, spec)
    if unique:
        codes, weights, quals, _ = dedup_reads(codes_raw, quals_raw)
    else:
        codes, weights = codes_raw, np.ones(n_raw)
        quals = np.where(quals_raw == 0, 2, quals_raw).astype(np.float64)
    n = codes.shape[0]
    zs, scs = init_state.draw_restarts(spec.mode, spec.n_clusters, spec.alpha0, n, spec.n_starts, spec.seed_init)
    kw = {}
    if spec.mode == "uqs":
        kw["log_hit"], kw["log_off"] = tensors.log_error_terms(quals)
    win = Window(mode=spec.mode, codes=codes, weights=weights.astype(np.float64), ref_codes=ref,
                 init_cluster=zs, init_scalars=scs, alpha0=spec.alpha0, conv_thres=spec.conv_thres,
                 max_iter=max_iter, name=f"{spec.name}-w{index}", **kw)
    win.n_raw = int(n_raw)
    win.qualities = quals
    win.true_haplotypes = haps
    win.true_frequencies = freqs
    return win


def _make_one(arg):
    name, index, master_seed, kw = arg
    return make_window(WORKLOADS[name], index, master_seed, **kw)


def make_workload(name: str, n_windows: int | None = None, master_seed: int = 0, indices=None, procs: int = 1, **kw) -> list:
    """The first ``n_windows`` windows of a workload (default: all), or the windows ``indices``.
    ``procs`` > 1 generates them in a process pool (host-only work; call it before the process
    touches CUDA -- the pool forks)."""
    spec = WORKLOADS[name]
    if indices is None:
        indices = range(spec.n_windows if n_windows is None else n_windows)
    indices = list(indices)
    if procs > 1 and len(indices) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, len(indices))) as pool:
            return pool.map(_make_one, [(name, i, master_seed, kw) for i in indices], chunksize=1)
    return [make_window(spec, i, master_seed, **kw) for i in indices]
