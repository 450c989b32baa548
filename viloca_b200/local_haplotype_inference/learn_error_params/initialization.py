"""Seeded initial state (learn_error_params/initialization.py:7-83)."""
from __future__ import annotations

import numpy as np
from scipy.special import digamma

from ... import init_state


def draw_init_state(n_clusters, alpha0, alphabet, reads_list, reference_binary, rng):
    genome_length = reads_list[0].seq_binary.shape[0]
    n_reads = len(reads_list)
    mean_z, sc = init_state.draw("lep", n_clusters, alpha0, n_reads, rng)
    alpha = alpha0 * np.ones(n_clusters)
    ref = np.asarray(reference_binary, dtype=np.float64)
    hit = np.exp(sc[2])
    miss = (1.0 - np.exp(sc[3])) / (len(alphabet) - 1)
    mean_h = np.broadcast_to(np.power(hit, ref) * np.power(miss, 1 - ref), (n_clusters, genome_length, len(alphabet))).copy()
    return {
        "alpha": alpha,
        "mean_log_pi": digamma(alpha) - digamma(alpha.sum(axis=0)),
        "theta_c": sc[4], "theta_d": sc[5], "mean_log_theta": (sc[6], sc[7]),
        "gamma_a": sc[0], "gamma_b": sc[1], "mean_log_gamma": (sc[2], sc[3]),
        "mean_haplo": mean_h,
        "mean_cluster": mean_z,
    }
