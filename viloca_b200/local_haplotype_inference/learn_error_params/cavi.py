"""``run_cavi`` / ``multistart_cavi`` with the reference's signatures
(learn_error_params/cavi.py:27-85), executed by the CUDA engine.

The reference's driver for this mode raises as shipped (its callers pass one argument fewer than the
functions take, and the loop appends to lists that are never defined).
These functions take the arguments the reference's *definitions* name, run the loop as written there
(ELBO compared every second update, decrease tolerance 1e-8, convergence threshold fixed at 1e-3,
loop counter set back while not converged) and return what the reference's ``run_cavi`` returns:
the 3-tuple ``(state_init_dict, state_curr_dict, dict_result)`` with the summary of
``analyze_results.summarize_results`` merged into the last two (cavi.py:241-255).
"""
from __future__ import annotations

import numpy as np

from ... import tensors
from ..._consts import INIT_STRIDE
from .. import _common
from . import initialization


def _prepared_from_reads_list(reads_list, reads_seq_binary, reads_weights):
    """The rows of ``reads_list`` (objects with ``id``, ``identical_reads``, ``seq_string``) as a
    PreparedWindow."""
    codes = tensors.codes_from_onehot(np.asarray(reads_seq_binary))
    ids = [[r.id] + list(getattr(r, "identical_reads", [])) for r in reads_list]
    seqs = [str(getattr(r, "seq_string", "")) for r in reads_list]
    return _common.PreparedWindow(codes=codes, weights=np.asarray(reads_weights), qualities=None, read_ids=ids,
                                  sequences=seqs, n_raw=sum(len(i) for i in ids))


def _run(K, alpha0, alphabet, reference_binary, reference_seq, reads_list, reads_seq_binary, reads_weights,
         output_dir, rngs, start_ids, max_iter=50000, device=0):
    from ...engine import CaviBatch
    from ...window import Window
    inits = [initialization.draw_init_state(K, alpha0, alphabet, reads_list, reference_binary, rng) for rng in rngs]
    prepared = _prepared_from_reads_list(reads_list, reads_seq_binary, reads_weights)
    zs = np.stack([st["mean_cluster"] for st in inits])
    scs = np.zeros((len(inits), INIT_STRIDE))
    for r, st in enumerate(inits):
        scs[r] = (st["gamma_a"], st["gamma_b"], st["mean_log_gamma"][0], st["mean_log_gamma"][1],
                  st["theta_c"], st["theta_d"], st["mean_log_theta"][0], st["mean_log_theta"][1])
    win = Window(mode="lep", codes=prepared.codes, weights=np.asarray(reads_weights, dtype=np.float64),
                 ref_codes=tensors.codes_from_onehot(np.asarray(reference_binary)), init_cluster=zs,
                 init_scalars=scs, alpha0=alpha0, max_iter=max_iter)
    job = _common.WindowJob(mode="lep", freads_in="", fref_in="", fname_qualities=None, output_dir=output_dir,
                            n_starts=len(rngs), K=K, alpha0=alpha0, alphabet=alphabet)
    job.prepared, job.window, job.ref_seq = prepared, win, str(reference_seq)
    results = []
    with CaviBatch([win], device=device) as batch:
        batch.upload()
        batch.run()
        for r, start_id in enumerate(start_ids):
            out = batch.fetch(state=r)[0]
            state, res = _common._result_dicts(job, out, r, out.state)
            summary = _common.summarize(job, batch, 0, r, state)
            state.update(summary)
            dict_result = {"run_id": start_id, "N": len(reads_list), "K": K, "alpha0": alpha0, "alphabet": alphabet,
                           "theta0": inits[r]["mean_log_theta"][0], "gamma0": inits[r]["mean_log_gamma"][0],
                           "mean_h0": inits[r]["mean_haplo"], "mean_z0": inits[r]["mean_cluster"],
                           "mean_log_pi": inits[r]["mean_log_pi"],
                           "exit_message": res["exit_message"], "exitflag": _exitflag(res["exit_message"]),
                           "n_iterations": res["n_iterations"], "converged": res["converged"], "elbo": res["elbo"],
                           "history_elbo": res["history_elbo"]}
            dict_result.update(summary)
            results.append((inits[r], state, dict_result))
    return results


def _exitflag(message: str):
    """cavi.py:137-192: 0 on convergence, -1 when the ELBO decreased, "" otherwise."""
    return 0 if message == "ELBO converged." else (-1 if message == "Error: ELBO is decreasing." else "")


def run_cavi(K, alpha0, alphabet, reference_binary, reference_seq, reads_list, reads_seq_binary, reads_weights,
             start_id, output_dir, record_history, rng, seed_number):
    """One CAVI run -> (state_init_dict, state_curr_dict, dict_result) (cavi.py:71-255)."""
    return _run(K, alpha0, alphabet, reference_binary, reference_seq, reads_list, reads_seq_binary, reads_weights,
                output_dir, [rng], [start_id])[0]


def multistart_cavi(K, alpha0, alphabet, reference_binary, reference_seq, reads_list, reads_seq_binary,
                    reads_weights, n_starts, output_dir, record_history, seed):
    """``n_starts`` independent restarts, restart s seeded with default_rng(seed + s) (cavi.py:42-44), all in
    one engine batch; results in restart order (the reference collects them in completion order in a
    module-global list that is never cleared)."""
    rngs = [np.random.default_rng(seed=seed + s) for s in range(n_starts)]
    return _run(K, alpha0, alphabet, reference_binary, reference_seq, reads_list, reads_seq_binary, reads_weights,
                output_dir, rngs, list(range(n_starts)))
