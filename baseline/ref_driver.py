"""Bounded samples of the reference's CPU CAVI for the two CPU legs of bench.py.

``time_iterations`` runs ``n_iter`` iterations of the loop body of ``cavi.run_cavi``
(use_quality_scores/cavi.py:120-169: digamma-sum refresh for iter <= 1, ``update_eqs.update``,
``elbo_eqs.compute_elbo``) on the dense tensors of one window and returns the seconds the loop took;
input construction is outside the clock.  The functions called are the UNMODIFIED reference
modules from ``baseline/_ref`` (kind "reference") when that copy exists, otherwise the NumPy oracle
port (kind "port").  The reference's ``run_cavi`` itself cannot be bounded (it has no iteration
cap and the large windows need 10^3-10^4 iterations of 0.5-50 s each), so the loop is driven here,
call for call, for a fixed number of iterations.

Nothing here imports the product package's CUDA binding.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import install_ref  # noqa: E402

N_BASES = 5


def kind() -> str:
    return "reference" if install_ref.available() else "port"


def dense_inputs(win):
    """(reads one-hot, weights, reference one-hot, E) as the reference's preparation step builds them
    (use_quality_scores/preparation.py:48-157): int8 one-hots and the dense FP64 E for uqs, float
    one-hots for lep; rows of zeros at 'N'."""
    dt = np.int8 if win.mode == "uqs" else np.float64

    def onehot(codes):
        out = np.zeros(codes.shape + (N_BASES,), dtype=dt)
        ok = codes < N_BASES
        out[np.nonzero(ok) + (codes[ok].astype(np.int64),)] = 1
        return out

    reads = onehot(win.codes)
    ref = onehot(win.ref_codes)
    E = None
    if win.mode == "uqs":
        E = np.where(reads > 0, win.log_hit[:, :, None], win.log_off[:, :, None])
        E[win.codes >= N_BASES] = 0
    return reads, np.asarray(win.weights), ref, E


def init_state(win, restart: int = 0) -> dict:
    """state_init_dict (initialization.py:6-43 + cavi.py:95-105) from the window's stored seeded draw."""
    from scipy.special import betaln, digamma
    from scipy.stats._multivariate import _lnB as lnB
    sc = win.init_scalars[restart]
    alpha = win.alpha0 * np.ones(win.n_clusters)
    st = {"alpha": alpha, "mean_log_pi": digamma(alpha) - digamma(alpha.sum(axis=0)),
          "gamma_a": sc[0], "gamma_b": sc[1], "mean_log_gamma": (sc[2], sc[3]),
          "mean_cluster": win.init_cluster[restart].copy(), "mean_haplo": None}
    st["lnB_alpha0"] = lnB(alpha)
    st["betaln_a0_b0"] = betaln(sc[0], sc[1])
    if win.mode == "lep":
        st.update({"theta_c": sc[4], "theta_d": sc[5], "mean_log_theta": (sc[6], sc[7])})
        st["betaln_c0_d0"] = betaln(sc[4], sc[5])
    return st


def _reference_loop(win, n_iter: int, restart: int):
    from scipy.special import digamma
    update_eqs, elbo_eqs, _ = install_ref.import_reference(win.mode)
    reads, weights, ref, E = dense_inputs(win)
    st0 = init_state(win, restart)
    cur = st0
    it, history = 0, []
    t0 = time.perf_counter()
    for _ in range(n_iter):
        if it <= 1:          # cavi.py:122-128
            cur["digamma_alpha_sum"] = digamma(cur["alpha"].sum(axis=0))
            cur["digamma_a_b_sum"] = digamma(cur["gamma_a"] + cur["gamma_b"])
            if win.mode == "lep":
                cur["digamma_c_d_sum"] = digamma(cur["theta_c"] + cur["theta_d"])
        if win.mode == "uqs":
            cur = update_eqs.update(reads, weights, ref, E, st0, cur)
            elbo = elbo_eqs.compute_elbo(weights, ref, E, st0, cur)
        else:
            cur = update_eqs.update(reads, weights, None, ref, st0, cur)
            elbo = elbo_eqs.compute_elbo(weights, reads, ref, st0, cur)
        cur["elbo"] = elbo
        # the loop counter as written: uqs counts up (cavi.py:169); lep appends the ELBO on even
        # counts only and sets the counter back to 0 while not converged (lep cavi.py:167-199), so
        # its digamma sums are refreshed every second update
        if win.mode == "uqs" or it % 2 == 0:
            history.append(elbo)
        if win.mode == "lep" and it > 1:
            it = it + 1 if abs(elbo - history[-2]) < 1e-3 else 0
        it += 1
    return time.perf_counter() - t0, n_iter, float(elbo)


def _port_loop(win, n_iter: int, restart: int):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _bridge as br
    w, ref, data = br.dense_inputs(win)
    init = br.oracle_init(win, restart)
    t0 = time.perf_counter()
    traj = br.orc.run(win.mode, w, ref, data, init, convergence_threshold=0.0, max_iterations=n_iter)
    return time.perf_counter() - t0, traj.n_updates, float(traj.history_elbo[-1])


def first_reads(win, n: int):
    """The window restricted to its first ``n`` unique reads (a bounded sample of a window whose single
    iteration would take minutes on one core; the cost of an iteration is linear in the reads)."""
    import dataclasses
    if n >= win.n_reads:
        return win
    kw = {"codes": win.codes[:n], "weights": win.weights[:n], "init_cluster": win.init_cluster[:, :n]}
    if win.mode == "uqs":
        kw.update({"log_hit": win.log_hit[:n], "log_off": win.log_off[:n]})
    return dataclasses.replace(win, **kw)


def time_iterations(win, n_iter: int, restart: int = 0):
    """(seconds, iterations done, last ELBO) of ``n_iter`` CAVI iterations on one window, one thread."""
    if install_ref.available():
        return _reference_loop(win, n_iter, restart)
    return _port_loop(win, n_iter, restart)
