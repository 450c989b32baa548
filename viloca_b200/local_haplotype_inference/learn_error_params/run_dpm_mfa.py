"""Entry point with the reference's signature (learn_error_params/run_dpm_mfa.py:33).

The reference's driver for this mode raises as shipped (TypeError, then NameError; SURVEY.md
§8(a) L-C1).  This entry point runs the loop as written there, minus the appends to the
undefined history lists: ELBO compared every second update, decrease tolerance 1e-8,
convergence threshold fixed at 1e-3."""
from __future__ import annotations

from .. import _common


def main(freads_in, fref_in, output_dir, n_starts, K, alpha0, alphabet="ACGT-", unique_modus=True,
         record_history=False, seed=27):
    job = _common.WindowJob(mode="lep", freads_in=freads_in, fref_in=fref_in, fname_qualities=None,
                            output_dir=output_dir, n_starts=int(n_starts), K=int(K), alpha0=float(alpha0),
                            alphabet=alphabet, unique_modus=unique_modus, record_history=record_history,
                            seed=seed)
    _common.run_jobs([job])


def main_batch(jobs, device=0):
    wj = [_common.WindowJob(mode="lep", fname_qualities=None, **{"n_starts": 1, **j}) for j in jobs]
    return _common.run_jobs(wj, device=device)
