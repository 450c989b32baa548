"""Glue between the product's compact windows and the dense tensors the oracle consumes."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cavi_oracle as orc  # noqa: E402


def dense_inputs(win):
    """(weights, ref_onehot, data) with data = E for uqs, float one-hot reads for lep."""
    # one-hot dtypes as in the reference pipeline (int8 for uqs, float64 for lep)
    dt = np.int8 if win.mode == "uqs" else np.float64
    reads = orc.onehot_from_codes(win.codes, dtype=dt)
    ref = orc.onehot_from_codes(win.ref_codes, dtype=dt)
    if win.mode == "uqs":
        E = np.where(reads > 0, win.log_hit[:, :, None], win.log_off[:, :, None])
        E[win.codes == 255] = 0
        return win.weights, ref, E
    return win.weights, ref, reads


def oracle_init(win, restart=0):
    """state_init_dict of the reference for the window's stored initial state."""
    sc = win.init_scalars[restart]
    K = win.n_clusters
    alpha = win.alpha0 * np.ones(K)
    st = {
        "alpha": alpha,
        "mean_log_pi": orc.digamma(alpha) - orc.digamma(alpha.sum(axis=0)),
        "gamma_a": sc[0], "gamma_b": sc[1], "mean_log_gamma": (sc[2], sc[3]),
        "mean_cluster": win.init_cluster[restart].copy(),
        "mean_haplo": None,
    }
    st["lnB_alpha0"] = orc.log_multivariate_beta(alpha)
    st["betaln_a0_b0"] = orc.betaln(sc[0], sc[1])
    if win.mode == "lep":
        st.update({"theta_c": sc[4], "theta_d": sc[5], "mean_log_theta": (sc[6], sc[7])})
        st["betaln_c0_d0"] = orc.betaln(sc[4], sc[5])
    return st


def oracle_run(win, restart=0, record=False, max_iterations=None):
    w, ref, data = dense_inputs(win)
    init = oracle_init(win, restart)
    return orc.run(win.mode, w, ref, data, init, convergence_threshold=win.conv_thres,
                   max_iterations=max_iterations, record=record)


def scalars_of(state, mode):
    """Pack an oracle state into the VL_SCALAR_STRIDE block vl_batch_set_state reads."""
    sc = np.zeros(16)
    sc[3], sc[4] = state["mean_log_gamma"]
    sc[9] = state["digamma_alpha_sum"]
    sc[10] = state["digamma_a_b_sum"]
    if mode == "lep":
        sc[7], sc[8] = state["mean_log_theta"]
        sc[11] = state["digamma_c_d_sum"]
    return sc


def rel_err(a, b, floor=1e-290):
    """max relative error over entries whose reference magnitude is above ``floor``; where the
    reference is below it (exact zeros, deep subnormals) the value must be below ``1e10 * floor`` too,
    otherwise the error is infinite (a non-zero where the reference has a zero is never ignored)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mask = np.abs(b) > floor
    if (np.abs(a[~mask]) > 1e10 * floor).any() or np.isnan(a).any() != np.isnan(b).any():
        return float("inf")
    if not mask.any():
        return 0.0
    return float(np.max(np.abs(a[mask] - b[mask]) / np.abs(b[mask])))
