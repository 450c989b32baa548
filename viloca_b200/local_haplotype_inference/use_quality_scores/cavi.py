"""``run_cavi`` / ``multistart_cavi`` with the reference's signatures
(use_quality_scores/cavi.py:26-84), executed by the CUDA engine."""
from __future__ import annotations

import numpy as np

from ... import tensors
from ..._consts import INIT_STRIDE
from . import initialization


def _window_from_dense(n_cluster, alpha0, reference_binary, reads_seq_binary, reads_weights,
                       reads_log_error_proba, convergence_threshold, init_states, max_iter):
    from ...engine import Window
    reads = np.asarray(reads_seq_binary)
    E = np.asarray(reads_log_error_proba, dtype=np.float64)
    codes = tensors.codes_from_onehot(reads)
    n, l = codes.shape
    rows, cols = np.meshgrid(np.arange(n), np.arange(l), indexing="ij")
    called = np.where(codes < reads.shape[2], codes, 0).astype(np.int64)
    log_hit = E[rows, cols, called]
    log_off = E[rows, cols, (called + 1) % reads.shape[2]]
    zs = np.stack([st["mean_cluster"] for st in init_states])
    scs = np.zeros((len(init_states), INIT_STRIDE))
    for r, st in enumerate(init_states):
        scs[r, :4] = (st["gamma_a"], st["gamma_b"], st["mean_log_gamma"][0], st["mean_log_gamma"][1])
        scs[r, 4:6] = 1.0
    return Window(mode="uqs", codes=codes, weights=np.asarray(reads_weights, dtype=np.float64),
                  ref_codes=tensors.codes_from_onehot(np.asarray(reference_binary)),
                  init_cluster=zs, init_scalars=scs, alpha0=alpha0, log_hit=log_hit, log_off=log_off,
                  conv_thres=convergence_threshold, max_iter=max_iter)


def _results(window, outcome_of_restart, init_states, alpha0, alphabet, n_cluster):
    out = []
    for r, o in enumerate(outcome_of_restart):
        rr = o.restarts[r]
        state = dict(o.state)
        state["elbo"] = rr.elbo
        res = {"exit_message": rr.exit_message, "n_iterations": rr.n_iterations, "converged": rr.converged,
               "elbo": rr.elbo, "history_elbo": list(rr.history_elbo), "run_id": r,
               "n_reads": window.n_reads, "n_cluster": n_cluster, "alpha0": alpha0, "alphabet": alphabet}
        res.update(state)
        out.append((state, res))
    return out


def _run(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary, reads_weights,
         reads_log_error_proba, convergence_threshold, rngs, max_iter=50000, device=0):
    from ...engine import CaviBatch
    reads = np.asarray(reads_seq_binary)
    inits = [initialization.draw_init_state(n_cluster, alpha0, alphabet, reads.shape[1], reads.shape[0],
                                            reference_binary, rng) for rng in rngs]
    win = _window_from_dense(n_cluster, alpha0, reference_binary, reads, reads_weights,
                             reads_log_error_proba, convergence_threshold, inits, max_iter)
    with CaviBatch([win], device=device) as batch:
        batch.upload()
        batch.run()
        per_restart = [batch.fetch(state=r)[0] for r in range(len(rngs))]
    return _results(win, per_restart, inits, alpha0, alphabet, n_cluster)


def run_cavi(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary, reads_weights,
             reads_log_error_proba, start_id, output_dir, convergence_threshold, record_history,
             rng, seed_number):
    """One CAVI run -> (state_curr_dict, dict_result), keys as in cavi.py:171-210."""
    state, res = _run(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary, reads_weights,
                      reads_log_error_proba, convergence_threshold, [rng])[0]
    res["run_id"] = start_id
    return state, res


def multistart_cavi(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary, reads_weights,
                    reads_log_error_proba, n_starts, output_dir, convergence_threshold,
                    record_history, seed):
    """R independent restarts, restart s seeded with default_rng(seed + s) (cavi.py:41-43).
    The reference collects results in a module-global list that is never cleared and cannot
    run inside a Pool worker; this returns exactly the R results, in
    restart order."""
    rngs = [np.random.default_rng(seed=seed + s) for s in range(n_starts)]
    return _run(n_cluster, alpha0, alphabet, reference_binary, reads_seq_binary, reads_weights,
                reads_log_error_proba, convergence_threshold, rngs)
