"""Full-shape goldens from the UNMODIFIED reference (TEST INFRASTRUCTURE; minutes of CPU time).

Runs only where /root/reference is mounted.  The windows are the product's synthetic windows of
the BASELINE.json shapes (regenerated from their seeds by the tests, so only small summaries of
the reference's results are committed):

  hiv5k_w96       configs[1]: one full hiv5k_uqs window (N_raw = 5000, L = 201, K = 100) through the
                  reference's own ``cavi.run_cavi`` to the end of its loop -- n_iterations, exit
                  message, the whole ELBO history, final alpha / mean_cluster / mean_haplo
  amplicon_full   configs[3] shape: sars_amplicon window 5 (N_raw = 20000 -> N_unique ~ 2500, L ~ 400,
                  K = 100), 12 iterations of ``update_eqs.update`` + ``elbo_eqs.compute_elbo`` driven as
                  ``run_cavi`` drives them (cavi.py:120-169)
  ont_full        configs[4] shape: N = 20000 unique ONT-like reads, L = 1000, K = 100, 2 restarts
                  (seeds 27, 28), 4 iterations each (widest kernel class, 8-way split haplotype tiles)

For the two large shapes the committed summary holds the ELBO per iteration, alpha, mean_log_pi,
gamma a/b, the number of exact zeros of mean_cluster, and mean_cluster / mean_haplo at a fixed
sample of rows / positions.

    python oracle/make_golden_fullshape.py
"""
from __future__ import annotations

import contextlib
import dataclasses
import io
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VILOCA_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden", "fullshape")

sys.path.insert(0, os.path.join(HERE, "_bioshim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from numpy.random import default_rng  # noqa: E402
from scipy.special import digamma  # noqa: E402

ROW_SAMPLE = 96       # rows of mean_cluster / positions of mean_haplo kept in the summaries


def full_shape_windows():
    """name -> (window, restarts' seeds, iterations); shared with tests/test_gpu_parity.py."""
    from viloca_b200 import synthetic
    amp = synthetic.make_window(synthetic.WORKLOADS["sars_amplicon"], 5, max_iter=12)
    ont_spec = dataclasses.replace(synthetic.WORKLOADS["sars_ont"], n_starts=2)
    ont = synthetic.make_window(ont_spec, 1, n_reads=20000, max_iter=4)
    return {"amplicon_full": (amp, 12), "ont_full": (ont, 4)}


def dense(win):
    def onehot(codes):
        out = np.zeros(codes.shape + (5,), dtype=np.int8)
        ok = codes < 5
        out[np.nonzero(ok) + (codes[ok].astype(np.int64),)] = 1
        return out
    reads = onehot(win.codes)
    from viloca.local_haplotype_inference.use_quality_scores import preparation
    E = preparation.compute_reads_log_error_proba(win.qualities, reads, 5)
    return reads, onehot(win.ref_codes), E


def sample_index(n, k=ROW_SAMPLE):
    return np.unique(np.linspace(0, n - 1, min(k, n)).astype(np.int64))


def drive(win, restart, n_iter, reads, ref, E):
    """n_iter iterations exactly as cavi.run_cavi sequences them (cavi.py:95-152)."""
    from viloca.local_haplotype_inference.use_quality_scores import elbo_eqs, initialization, update_eqs
    seed = 27 + restart                                   # cavi.py:43
    st0 = initialization.draw_init_state(win.n_clusters, win.alpha0, "ACGT-", win.length, win.n_reads, ref, default_rng(seed))
    assert np.array_equal(st0["mean_cluster"], win.init_cluster[restart]), "the window's stored draw is the reference's"
    from scipy.special import betaln
    from scipy.stats._multivariate import _lnB as lnB
    st0.update({"lnB_alpha0": lnB(st0["alpha"]), "betaln_a0_b0": betaln(st0["gamma_a"], st0["gamma_b"])})   # cavi.py:95-105
    cur, hist = st0, []
    for it in range(n_iter):
        if it <= 1:
            cur["digamma_alpha_sum"] = digamma(cur["alpha"].sum(axis=0))
            cur["digamma_a_b_sum"] = digamma(cur["gamma_a"] + cur["gamma_b"])
        cur = update_eqs.update(reads, win.weights, ref, E, st0, cur)
        cur["elbo"] = elbo_eqs.compute_elbo(win.weights, ref, E, st0, cur)
        hist.append(cur["elbo"])
    return cur, np.array(hist)


def summary(prefix, state, hist, out):
    z, h = state["mean_cluster"], state["mean_haplo"]
    out[prefix + "history"] = hist
    out[prefix + "alpha"] = state["alpha"]
    out[prefix + "mean_log_pi"] = state["mean_log_pi"]
    out[prefix + "gamma_ab"] = np.array([state["gamma_a"], state["gamma_b"]])
    out[prefix + "z_zeros"] = int((z == 0).sum())
    out[prefix + "z_rows"] = z[sample_index(z.shape[0])]
    out[prefix + "h_positions"] = h[:, sample_index(h.shape[1])]
    out[prefix + "z_colsum"] = z.sum(axis=0)


def golden_large():
    for name, (win, n_iter) in full_shape_windows().items():
        t0 = time.time()
        reads, ref, E = dense(win)
        out = {"n_reads": win.n_reads, "length": win.length, "n_iter": n_iter, "n_starts": win.n_starts}
        for r in range(win.n_starts):
            state, hist = drive(win, r, n_iter, reads, ref, E)
            summary(f"r{r}_", state, hist, out)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(f"{name}: N {win.n_reads} L {win.length} R {win.n_starts} x {n_iter} iterations in {time.time() - t0:.0f}s")


def golden_hiv_window(index=96):
    from viloca.local_haplotype_inference.use_quality_scores import cavi
    from viloca_b200 import synthetic
    win = synthetic.make_window(synthetic.WORKLOADS["hiv5k_uqs"], index)
    reads, ref, E = dense(win)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        state, res = cavi.run_cavi(win.n_clusters, win.alpha0, "ACGT-", ref, reads, win.weights, E, 0, "./", win.conv_thres,
                                   False, default_rng(seed=27), 27)
    out = {"index": index, "n_reads": win.n_reads, "n_iterations": res["n_iterations"], "exit_message": res["exit_message"],
           "converged": res["converged"], "history": np.array(res["history_elbo"]), "alpha": state["alpha"],
           "mean_cluster": state["mean_cluster"], "mean_haplo": state["mean_haplo"]}
    np.savez_compressed(os.path.join(GOLD, f"hiv5k_w{index}.npz"), **out)
    print(f"hiv5k window {index}: {res['n_iterations']} iterations ({res['exit_message']}) in {time.time() - t0:.0f}s")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: goldens can only be regenerated where the reference is mounted")
    os.makedirs(GOLD, exist_ok=True)
    os.chdir(tempfile.mkdtemp())
    which = sys.argv[1:] or ["hiv", "large"]
    if "hiv" in which:
        golden_hiv_window()
    if "large" in which:
        golden_large()
