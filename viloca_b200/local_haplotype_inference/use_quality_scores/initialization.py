"""Seeded initial state with the reference's RNG call order
(use_quality_scores/initialization.py:6-72)."""
from __future__ import annotations

import numpy as np
from scipy.special import digamma

from ... import init_state


def draw_init_state(n_clusters, alpha0, alphabet, genome_length, n_reads, reference_binary, rng):
    """Same keys as the reference's state_init_dict.  ``mean_haplo`` (initialization.py:57-72)
    is never consumed by the updates (update_eqs.py:20 recomputes it first) and is provided
    only for callers that inspect the dict."""
    mean_z, sc = init_state.draw("uqs", n_clusters, alpha0, n_reads, rng)
    alpha = alpha0 * np.ones(n_clusters)
    ref = np.asarray(reference_binary, dtype=np.float64)
    b = len(alphabet)
    hit = np.exp(sc[2])
    miss = (1.0 - np.exp(sc[3])) / (b - 1)
    mean_h = np.broadcast_to(np.power(hit, ref) * np.power(miss, 1 - ref), (n_clusters,) + ref.shape).copy()
    return {
        "alpha": alpha,
        "mean_log_pi": digamma(alpha) - digamma(alpha.sum(axis=0)),
        "gamma_a": sc[0], "gamma_b": sc[1],
        "mean_log_gamma": (sc[2], sc[3]),
        "mean_haplo": mean_h,
        "mean_cluster": mean_z,
    }
