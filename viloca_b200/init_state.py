"""Seeded initial states, drawn on the host with the reference's exact RNG call order.

``numpy.random.Generator`` streams carry no cross-version guarantee, so the initial state is
always produced here by the installed NumPy and *passed in* to the kernels (BASELINE.json:
"seeded initialisation passed in").
"""
from __future__ import annotations

import numpy as np

from . import _consts as K_

MATCHES = 10       # use_quality_scores/initialization.py:17
MISMATCH = 2       # use_quality_scores/initialization.py:18


def draw(mode: str, n_clusters: int, alpha0: float, n_reads: int, rng: np.random.Generator):
    """Returns (mean_cluster (N, K), scalars (8,)) for one restart.

    uqs order (use_quality_scores/initialization.py:20-31): uniform(0.5, 1, 2) -> beta(a, b)
    -> dirichlet.  lep order (learn_error_params/initialization.py:23-41): uniform(0.5, 1, 4)
    -> beta(c, d) -> beta(a, b) -> dirichlet.
    """
    sc = np.zeros(K_.INIT_STRIDE)
    if mode == "uqs":
        k = rng.uniform(low=0.5, high=1.0, size=2)
        a, b = MATCHES * k[0], MISMATCH * k[1]
        gamma0 = rng.beta(a, b)
        sc[4:8] = (1.0, 1.0, 0.0, 0.0)
    elif mode == "lep":
        k = rng.uniform(low=0.5, high=1.0, size=4)
        a, b = MATCHES * k[0], MISMATCH * k[1]
        c, d = MATCHES * k[2], MISMATCH * k[3]
        theta0 = rng.beta(c, d)
        gamma0 = rng.beta(a, b)
        sc[4:8] = (c, d, np.log(theta0), np.log(1 - theta0))
    else:
        raise ValueError(mode)
    sc[0:4] = (a, b, np.log(gamma0), np.log(1 - gamma0))
    mean_z = rng.dirichlet(np.ones(n_clusters) * alpha0, size=n_reads)
    if np.any(np.isnan(mean_z)):
        # initialization.py:50-52 retries with alpha0*10 but omits `rng` (TypeError upstream)
        raise FloatingPointError("NaN in the Dirichlet initialisation of mean_cluster")
    return mean_z, sc


def draw_restarts(mode: str, n_clusters: int, alpha0: float, n_reads: int, n_starts: int, seed: int):
    """Restart s uses default_rng(seed + s) (use_quality_scores/cavi.py:43); a single start
    uses default_rng(seed) (run_dpm_mfa.py:74) — the same thing for s = 0."""
    zs = np.empty((n_starts, n_reads, n_clusters))
    scs = np.empty((n_starts, K_.INIT_STRIDE))

    def one(s):
        zs[s], scs[s] = draw(mode, n_clusters, alpha0, n_reads, np.random.default_rng(seed=seed + s))

    # every restart has its own generator, so the draws are independent of each other: large windows
    # (ONT: 100 000 reads x 100 clusters x 20 restarts, 0.9 s per restart) draw them side by side on
    # host threads (Generator.dirichlet releases the GIL); same bits as the serial loop
    if n_starts > 1 and float(n_reads) * n_clusters * n_starts > 5e6:
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(n_starts, max(1, (os.cpu_count() or 2) - 1))) as ex:
            list(ex.map(one, range(n_starts)))
    else:
        for s in range(n_starts):
            one(s)
    return zs, scs
