"""Entry point with the reference's signature (use_quality_scores/run_dpm_mfa.py:18-31)."""
from __future__ import annotations

from .. import _common


def main(freads_in, fref_in, fname_qualities, output_dir, n_starts, K, alpha0, alphabet="ACGT-",
         unique_modus=True, convergence_threshold=1e-03, record_history=False, seed=27):
    """Reads ``freads_in`` / ``fref_in`` / ``fname_qualities``, runs ``n_starts`` CAVI restarts on
    the GPU, keeps the maximal-ELBO run and writes ``<window>-support.fas`` and
    ``<window>-cor.fas`` into ``output_dir``.  Returns None like the reference."""
    job = _common.WindowJob(mode="uqs", freads_in=freads_in, fref_in=fref_in, fname_qualities=fname_qualities,
                            output_dir=output_dir, n_starts=int(n_starts), K=int(K), alpha0=float(alpha0),
                            alphabet=alphabet, unique_modus=unique_modus,
                            convergence_threshold=convergence_threshold, record_history=record_history,
                            seed=seed)
    _common.run_jobs([job])


def main_batch(jobs, device=0):
    """Batched form: ``jobs`` is a list of keyword dicts of :func:`main`; all windows (and all
    their restarts) run in one engine batch on ``device``."""
    wj = [_common.WindowJob(mode="uqs", **{"n_starts": 1, **j}) for j in jobs]
    return _common.run_jobs(wj, device=device)
