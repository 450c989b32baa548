"""CPU oracle for VILOCA's local-haplotype CAVI (TEST INFRASTRUCTURE — not product code).

A NumPy restatement of the reference algorithm for the one hot path this repository
accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module; the product package ``viloca_b200`` never does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (possible in the build container only), drives it on seeded inputs and
on the reference's own fixture ``tests/data_4``, and stores its outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this restatement against those
vectors (bit-for-bit for the update equations, since the same NumPy primitives are called in
the same order) and against the survey's known-answer run (54 iterations, final ELBO
-1943.9785808934962 on ``tests/data_4``).  The reference ships no test of its own for these
functions, so the reference's *outputs run here* are the pin.

Third-party arithmetic the reference relies on and this oracle calls the same way:
``numpy.einsum`` (no ``optimize``), ``numpy.random.Generator`` (uniform/beta/dirichlet),
``scipy.special.{digamma, betaln, gammaln}``.

All ``file:line`` citations are relative to ``/root/reference/viloca/local_haplotype_inference``;
``uqs`` = ``use_quality_scores``, ``lep`` = ``learn_error_params``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
from scipy.special import betaln, digamma, gammaln

ALPHABET = "ACGT-"
N_BASES = len(ALPHABET)
PRIOR_MATCH = 10.0      # uqs/initialization.py:17, lep/initialization.py:20
PRIOR_MISMATCH = 2.0    # uqs/initialization.py:18, lep/initialization.py:21
MIN_ITERATIONS = 10     # uqs/cavi.py:119, lep/cavi.py:141

MSG_CONVERGED = "ELBO converged."
MSG_DECREASING = "Error: ELBO is decreasing."
MSG_NAN = "Error: ELBO is nan."


def log_multivariate_beta(x):
    """ln B(x) = sum lgamma(x_i) - lgamma(sum x_i); scipy.stats._multivariate._lnB as
    used by uqs/cavi.py:4,100 and uqs/elbo_eqs.py:3,59."""
    x = np.asarray(x, dtype=np.float64)
    return np.sum(gammaln(x)) - gammaln(np.sum(x))


# --------------------------------------------------------------------------------------
# tensors built before the loop
# --------------------------------------------------------------------------------------

def onehot_from_codes(codes, n_bases=N_BASES, dtype=np.float64):
    """codes (…,) uint8 with 255 = not-in-alphabet -> one-hot (…, B) with an all-zero row
    for 255 (uqs/preparation.py:58-63, lep/preparation.py:15-22)."""
    codes = np.asarray(codes)
    out = np.zeros(codes.shape + (n_bases,), dtype=dtype)
    ok = codes < n_bases
    idx = np.nonzero(ok)
    out[idx + (codes[ok].astype(np.int64),)] = 1
    return out


def reads_log_error_tensor(qualities, reads_onehot):
    """Q-T1, uqs/preparation.py:123-157.  theta = 1-10^(-Q/10); entry = log theta at the
    called base, log((1-theta)/(B-1)) elsewhere, all-zero row where the read has 'N'."""
    n_bases = reads_onehot.shape[2]
    theta = 1 - 10 ** (-qualities / 10)
    off = (1 - theta) / (n_bases - 1)
    log_hit = np.tile(np.log(theta)[:, :, np.newaxis], (1, 1, n_bases))
    log_off = np.tile(np.log(off)[:, :, np.newaxis], (1, 1, n_bases))
    tensor = np.einsum("NLB,NLB->NLB", log_hit, reads_onehot)
    tensor += np.einsum("NLB,NLB->NLB", log_off, 1 - reads_onehot)
    called = np.tile((reads_onehot.sum(axis=2) > 0)[:, :, np.newaxis], (1, 1, n_bases))
    tensor[~called] = 0
    return tensor


# --------------------------------------------------------------------------------------
# initial state (host RNG; the order of the draws is part of the contract)
# --------------------------------------------------------------------------------------

def _initial_haplotype_table(n_clusters, length, n_bases, mean_log_gamma, ref_onehot):
    """Q-I3, uqs/initialization.py:57-72 (never consumed by the updates)."""
    hit = np.exp(mean_log_gamma[0]) * np.ones((n_clusters, length, n_bases))
    miss = (1.0 - np.exp(mean_log_gamma[1])) / (n_bases - 1) * np.ones((n_clusters, length, n_bases))
    return np.multiply(np.power(hit, ref_onehot), np.power(miss, 1 - ref_onehot))


def _draw_responsibilities(n_clusters, n_reads, alpha0, rng):
    """Q-I2, uqs/initialization.py:46-54.  The reference's NaN retry omits ``rng`` and would
    raise TypeError; the oracle raises ValueError in that (never observed) case."""
    z = rng.dirichlet(np.ones(n_clusters) * alpha0, size=n_reads)
    if np.any(np.isnan(z)):
        raise ValueError("NaN in Dirichlet initialisation of mean_cluster")
    return z


def draw_init_state(mode, n_clusters, alpha0, length, n_reads, ref_onehot, rng):
    """Q-I1 uqs/initialization.py:6-43; L-I1 lep/initialization.py:7-55.

    uqs draw order: uniform(size=2) -> beta(a,b) -> dirichlet.
    lep draw order: uniform(size=4) -> beta(c,d) (theta0) -> beta(a,b) (gamma0) -> dirichlet.
    """
    alpha = alpha0 * np.ones(n_clusters)
    mean_log_pi = digamma(alpha) - digamma(alpha.sum(axis=0))
    state = {"alpha": alpha, "mean_log_pi": mean_log_pi}
    if mode == "uqs":
        k = rng.uniform(low=0.5, high=1.0, size=2)
        a, b = PRIOR_MATCH * k[0], PRIOR_MISMATCH * k[1]
        gamma0 = rng.beta(a, b)
    elif mode == "lep":
        k = rng.uniform(low=0.5, high=1.0, size=4)
        a, b = PRIOR_MATCH * k[0], PRIOR_MISMATCH * k[1]
        c, d = PRIOR_MATCH * k[2], PRIOR_MISMATCH * k[3]
        theta0 = rng.beta(c, d)
        gamma0 = rng.beta(a, b)
        state.update({"theta_c": c, "theta_d": d,
                      "mean_log_theta": (np.log(theta0), np.log(1 - theta0))})
    else:
        raise ValueError(mode)
    mean_log_gamma = (np.log(gamma0), np.log(1 - gamma0))
    state.update({
        "gamma_a": a, "gamma_b": b, "mean_log_gamma": mean_log_gamma,
        "mean_haplo": _initial_haplotype_table(n_clusters, length, ref_onehot.shape[1],
                                               mean_log_gamma, ref_onehot),
        "mean_cluster": _draw_responsibilities(n_clusters, n_reads, alpha0, rng),
    })
    # uqs/cavi.py:98-105, lep/cavi.py:113-123
    state["lnB_alpha0"] = log_multivariate_beta(state["alpha"])
    state["betaln_a0_b0"] = betaln(a, b)
    if mode == "lep":
        state["betaln_c0_d0"] = betaln(c, d)
    return state


# --------------------------------------------------------------------------------------
# shared small pieces
# --------------------------------------------------------------------------------------

def _softmax_last_axis(logits):
    """max-subtracted softmax over the last axis; uqs/update_eqs.py:64-69 and :94-99."""
    top = np.max(logits, axis=-1)[..., np.newaxis]
    e = np.exp(logits - top)
    return e / e.sum(axis=-1)[..., np.newaxis]


def _reference_term(ref_onehot, mean_log_gamma):
    """uqs/update_eqs.py:82, lep/update_eqs.py:125-127."""
    n_bases = ref_onehot.shape[1]
    return ref_onehot * mean_log_gamma[0] + (1 - ref_onehot) * (mean_log_gamma[1] - np.log(n_bases - 1))


def _entropy_sum(p):
    """sum p*log p with 0*log 0 := 0.  The reference calls ``np.log(p, where=p>0)`` without
    ``out=`` (uqs/elbo_eqs.py:50-51, 75-76), leaving masked entries uninitialised before they
    are multiplied by 0; the oracle defines them as 0."""
    logp = np.zeros_like(p)
    np.log(p, where=p > 0, out=logp)
    subs = "NK,NK->" if p.ndim == 2 else "KLB,KLB->"
    return np.einsum(subs, logp, p)


def _pi_term(alpha0_vec, lnB_alpha0, alpha, mean_log_pi):
    """Q-E3 uqs/elbo_eqs.py:56-62."""
    p = (-1) * lnB_alpha0 + np.multiply(alpha0_vec - 1, mean_log_pi).sum(axis=0)
    q = (-1) * log_multivariate_beta(alpha) + np.multiply(alpha - 1, mean_log_pi).sum(axis=0)
    return p - q


def _cluster_term(z, mean_log_pi, weights):
    """Q-E2 uqs/elbo_eqs.py:45-53 (entropy is not weighted by the read multiplicity)."""
    wz = np.einsum("N,NK->NK", weights, z)
    return np.matmul(wz, mean_log_pi).sum(axis=0) - _entropy_sum(z)


def _haplo_term(ref_onehot, h, mean_log_gamma):
    """Q-E4 uqs/elbo_eqs.py:65-78."""
    n_bases = h.shape[2]
    b1 = mean_log_gamma[0]
    b2 = mean_log_gamma[1] - np.log(n_bases - 1)
    p = b1 * np.einsum("KLB,LB->", h, ref_onehot)
    p += b2 * np.einsum("KLB,LB->", h, (1 - ref_onehot))
    return p - _entropy_sum(h)


def _beta_term(a0, b0, betaln_0, mean_log, a_up, b_up):
    """Q-E5 uqs/elbo_eqs.py:81-88 (also used for theta in lep, lep/elbo_eqs.py:39)."""
    p = (a0 - 1) * mean_log[0] + (b0 - 1) * mean_log[1] - betaln_0
    q = (a_up - 1) * mean_log[0] + (b_up - 1) * mean_log[1] - betaln(a_up, b_up)
    return p - q


def _dirichlet_counts(alpha0_vec, z, weights):
    """Q-U3 uqs/update_eqs.py:114-121."""
    return alpha0_vec.copy() + np.einsum("N,NK->NK", weights, z.copy()).sum(axis=0)


def _mutation_counts(ref_onehot, h, a0, b0):
    """Q-U5 uqs/update_eqs.py:103-112."""
    a = np.float64(a0) + np.einsum("KLB,LB->", h, ref_onehot)
    b = np.float64(b0) + np.einsum("KLB,LB->", h, (1 - ref_onehot))
    return a, b


# --------------------------------------------------------------------------------------
# use_quality_scores
# --------------------------------------------------------------------------------------

def uqs_update_mean_haplo(weights, ref_onehot, E, z, mean_log_gamma):
    """Q-U1 uqs/update_eqs.py:73-101: contraction #2 + softmax over bases."""
    wz = np.einsum("N,NK->NK", weights, z)
    logits = np.einsum("NLB,NK->KLB", E, wz)
    logits[:] += _reference_term(ref_onehot, mean_log_gamma)
    return _softmax_last_axis(logits)


def uqs_update_mean_cluster(mean_log_pi, h, E):
    """Q-U2 uqs/update_eqs.py:55-71: contraction #1 + softmax over clusters."""
    logits = np.einsum("NLB,KLB->NK", E, h)
    logits[:] += mean_log_pi
    return _softmax_last_axis(logits)


def uqs_step(weights, ref_onehot, E, init, cur):
    """Q-U0 uqs/update_eqs.py:5-38.  a0, b0 and alpha0 always come from ``init``; the two
    digamma sums come from ``cur`` (stale by design, see ``run``)."""
    h = uqs_update_mean_haplo(weights, ref_onehot, E, cur["mean_cluster"], cur["mean_log_gamma"])
    z = uqs_update_mean_cluster(cur["mean_log_pi"], h, E)
    alpha = _dirichlet_counts(init["alpha"], z, weights)
    mean_log_pi = digamma(alpha) - cur["digamma_alpha_sum"]
    a, b = _mutation_counts(ref_onehot, h, init["gamma_a"], init["gamma_b"])
    mean_log_gamma = (digamma(a) - cur["digamma_a_b_sum"], digamma(b) - cur["digamma_a_b_sum"])
    return {"alpha": alpha, "mean_log_pi": mean_log_pi, "gamma_a": a, "gamma_b": b,
            "digamma_alpha_sum": cur["digamma_alpha_sum"],
            "digamma_a_b_sum": cur["digamma_a_b_sum"],
            "mean_log_gamma": mean_log_gamma, "mean_haplo": h, "mean_cluster": z}


def uqs_elbo_terms(weights, ref_onehot, E, init, cur):
    """Q-E0..E5 uqs/elbo_eqs.py:6-88; returns the five terms in the reference's order."""
    z, h = cur["mean_cluster"], cur["mean_haplo"]
    x = np.einsum("NLB,KLB->NK", E, h)                                # Q-E1 :38
    data = np.einsum("NK,NK->", np.einsum("N,NK->NK", weights, z), x)
    return (
        data,
        _pi_term(init["alpha"], init["lnB_alpha0"], cur["alpha"], cur["mean_log_pi"]),
        _cluster_term(z, cur["mean_log_pi"], weights),
        _haplo_term(ref_onehot, h, cur["mean_log_gamma"]),
        _beta_term(init["gamma_a"], init["gamma_b"], init["betaln_a0_b0"],
                   cur["mean_log_gamma"], cur["gamma_a"], cur["gamma_b"]),
    )


# --------------------------------------------------------------------------------------
# learn_error_params
# --------------------------------------------------------------------------------------

def _masked_complement(reads_onehot):
    """nonN * (1 - r); lep/update_eqs.py:131-135, lep/elbo_eqs.py:52-55."""
    called = reads_onehot.sum(axis=2) > 0
    return np.einsum("NL,NLB->NLB", called, np.add(np.ones(reads_onehot.shape), (-1) * reads_onehot))


def lep_update_mean_haplo(reads_onehot, weights, ref_onehot, z, mean_log_theta, mean_log_gamma):
    """L-U1 lep/update_eqs.py:115-159."""
    n_bases = ref_onehot.shape[1]
    b1 = mean_log_theta[0]
    b2 = mean_log_theta[1] - np.log(n_bases - 1)
    inv = _masked_complement(reads_onehot)
    wz = np.einsum("N,NK->NK", weights, z)
    logits = b1 * np.einsum("NLB,NK->KLB", reads_onehot, wz)
    logits += b2 * np.einsum("NLB,NK->KLB", inv, wz)
    logits[:] += _reference_term(ref_onehot, mean_log_gamma)
    return _softmax_last_axis(logits)


def lep_update_mean_cluster(reads_onehot, mean_log_pi, h, mean_log_theta):
    """L-U2 lep/update_eqs.py:90-112 — note the *unmasked* (1 - r) here."""
    n_bases = h.shape[2]
    logits = np.einsum("NLB,KLB->NK", reads_onehot, mean_log_theta[0] * h)
    logits += np.einsum("NLB,KLB->NK", (1 - reads_onehot), (mean_log_theta[1] - np.log(n_bases - 1)) * h)
    logits[:] += mean_log_pi
    return _softmax_last_axis(logits)


def lep_error_counts(reads_onehot, weights, z, h, c0, d0):
    """L-U3 lep/update_eqs.py:174-188."""
    wz = np.einsum("N,NK->NK", weights, z)
    c = np.float64(c0) + np.einsum("NK,NK->", np.einsum("NLB,KLB->NK", reads_onehot, h), wz)
    d = np.float64(d0) + np.einsum("NK,NK->", np.einsum("NLB,KLB->NK", reads_onehot, (1 - h)), wz)
    return c, d


def lep_step(reads_onehot, weights, ref_onehot, init, cur):
    """L-U0 lep/update_eqs.py:5-70."""
    h = lep_update_mean_haplo(reads_onehot, weights, ref_onehot, cur["mean_cluster"],
                              cur["mean_log_theta"], cur["mean_log_gamma"])
    z = lep_update_mean_cluster(reads_onehot, cur["mean_log_pi"], h, cur["mean_log_theta"])
    alpha = _dirichlet_counts(init["alpha"], z, weights)
    mean_log_pi = digamma(alpha) - cur["digamma_alpha_sum"]
    a, b = _mutation_counts(ref_onehot, h, init["gamma_a"], init["gamma_b"])
    mean_log_gamma = (digamma(a) - cur["digamma_a_b_sum"], digamma(b) - cur["digamma_a_b_sum"])
    c, d = lep_error_counts(reads_onehot, weights, z, h, init["theta_c"], init["theta_d"])
    mean_log_theta = (digamma(c) - cur["digamma_c_d_sum"], digamma(d) - cur["digamma_c_d_sum"])
    return {"alpha": alpha, "mean_log_pi": mean_log_pi, "theta_c": c, "theta_d": d,
            "mean_log_theta": mean_log_theta, "gamma_a": a, "gamma_b": b,
            "digamma_alpha_sum": cur["digamma_alpha_sum"],
            "digamma_c_d_sum": cur["digamma_c_d_sum"],
            "digamma_a_b_sum": cur["digamma_a_b_sum"],
            "mean_log_gamma": mean_log_gamma, "mean_haplo": h, "mean_cluster": z}


def lep_elbo_terms(reads_onehot, weights, ref_onehot, init, cur):
    """L-E1/E2 lep/elbo_eqs.py:6-63; six terms in the reference's order."""
    z, h = cur["mean_cluster"], cur["mean_haplo"]
    n_bases = h.shape[2]
    b1 = cur["mean_log_theta"][0]
    b2 = cur["mean_log_theta"][1] - np.log(n_bases - 1)
    inv = _masked_complement(reads_onehot)
    x = np.einsum("NLB,KLB->NK", b1 * reads_onehot, h)
    x += np.einsum("NLB,KLB->NK", b2 * inv, h)
    data = np.einsum("NK,NK->", np.einsum("N,NK->NK", weights, z), x)
    return (
        data,
        _pi_term(init["alpha"], init["lnB_alpha0"], cur["alpha"], cur["mean_log_pi"]),
        _cluster_term(z, cur["mean_log_pi"], weights),
        _haplo_term(ref_onehot, h, cur["mean_log_gamma"]),
        _beta_term(init["gamma_a"], init["gamma_b"], init["betaln_a0_b0"],
                   cur["mean_log_gamma"], cur["gamma_a"], cur["gamma_b"]),
        _beta_term(init["theta_c"], init["theta_d"], init["betaln_c0_d0"],
                   cur["mean_log_theta"], cur["theta_c"], cur["theta_d"]),
    )


# --------------------------------------------------------------------------------------
# drivers
# --------------------------------------------------------------------------------------

@dataclass
class Trajectory:
    """What one CAVI run leaves behind (keys of ``dict_result``, uqs/cavi.py:190-206)."""
    state: dict
    elbo: float
    history_elbo: list
    n_iterations: int          # the loop variable `iter` at exit (dict_result["n_iterations"])
    n_updates: int             # update() calls actually performed
    converged: bool
    exit_message: str
    states: list = field(default_factory=list)   # per-iteration states if recorded
    terms: list = field(default_factory=list)    # per-iteration ELBO terms if recorded


def _refresh_digamma_sums(mode, cur):
    """uqs/cavi.py:122-128, lep/cavi.py:143-154 — evaluated on the state *before* the update,
    and only while iter <= 1, which makes the sums stale at iteration 0."""
    cur["digamma_alpha_sum"] = digamma(cur["alpha"].sum(axis=0))
    cur["digamma_a_b_sum"] = digamma(cur["gamma_a"] + cur["gamma_b"])
    if mode == "lep":
        cur["digamma_c_d_sum"] = digamma(cur["theta_c"] + cur["theta_d"])


def run(mode, weights, ref_onehot, data, init, convergence_threshold=1e-3,
        max_iterations=None, record=False):
    """The CAVI loop.

    ``data`` is E (N,L,B) for uqs and the float one-hot reads for lep.

    uqs: uqs/cavi.py:114-169 (Q-C1).  ELBO appended every iteration; checks start at iter 2;
    "decreasing" needs a drop > 1e-5; convergence is sticky and the loop then only waits
    for iter >= 10.

    lep: lep/cavi.py:135-199 (L-C1) *as written*, minus the appends to undefined history lists
    (:169-172) that make the shipped driver raise NameError: ELBO appended on even iter only;
    when iter > 1: decreasing threshold 1e-8, convergence threshold fixed 1e-3 (the
    ``convergence_threshold`` argument is ignored, shotgun.py:213), converged -> iter += 1
    extra, otherwise iter = 0; then iter += 1.
    """
    step = uqs_step if mode == "uqs" else lep_step
    terms_fn = uqs_elbo_terms if mode == "uqs" else lep_elbo_terms
    cur = dict(init)
    history, states, terms_log = [], [], []
    it, total, converged, message, elbo = 0, 0, False, "", 0
    while (converged is False) or (it < MIN_ITERATIONS):
        if it <= 1:
            _refresh_digamma_sums(mode, cur)
        cur = step(data, weights, ref_onehot, init, cur) if mode == "lep" else \
            step(weights, ref_onehot, data, init, cur)
        terms = terms_fn(data, weights, ref_onehot, init, cur) if mode == "lep" else \
            terms_fn(weights, ref_onehot, data, init, cur)
        elbo = terms[0]
        for t in terms[1:]:
            elbo = elbo + t
        total += 1
        if record:
            states.append(cur)
            terms_log.append(terms)
        if mode == "uqs":
            history.append(elbo)
            if it > 1:
                if np.isnan(elbo):
                    message = MSG_NAN
                    break
                elif (history[-2] > elbo) and np.abs(elbo - history[-2]) > 1e-05:
                    message = MSG_DECREASING
                    break
                elif np.abs(elbo - history[-2]) < convergence_threshold:
                    converged = True
                    message = MSG_CONVERGED
        else:
            if it % 2 == 0:
                history.append(elbo)
            if it > 1:
                if np.isnan(elbo):
                    message = MSG_NAN
                    break
                elif (history[-2] > elbo) and np.abs(elbo - history[-2]) > 1e-08:
                    message = MSG_DECREASING
                    break
                elif np.abs(elbo - history[-2]) < 1e-03:
                    converged = True
                    it += 1
                    message = MSG_CONVERGED
                else:
                    it = 0
        cur["elbo"] = elbo
        it += 1
        if max_iterations is not None and total >= max_iterations:
            if not message:
                message = "Maximum number of iterations reached."
            break
    cur["elbo"] = elbo
    return Trajectory(state=cur, elbo=elbo, history_elbo=history,
                      n_iterations=it, n_updates=total, converged=converged, exit_message=message, states=states, terms=terms_log)


# --------------------------------------------------------------------------------------
# finishing (behind the same entry point; defines the on-disk outputs)
# --------------------------------------------------------------------------------------

def haplotype_strings(h, alphabet=ALPHABET):
    """Q-F1 uqs/analyze_results.py:102-107: per position argmax over bases, first max wins."""
    idx = np.argmax(h, axis=2)
    return ["".join(alphabet[i] for i in row) for row in idx]


def unique_haplotypes(h, alphabet=ALPHABET):
    """Q-F1 uqs/analyze_results.py:144-157: group identical strings, sort lexicographically."""
    groups = {}
    for k, s in enumerate(haplotype_strings(h, alphabet)):
        groups.setdefault(s, []).append(k)
    return sorted(groups.items())


def merge_clusters(z, groups):
    """Q-F2 uqs/analyze_results.py:160-167."""
    merged = np.zeros((z.shape[0], len(groups)))
    for n in range(z.shape[0]):
        for j, (_, members) in enumerate(groups):
            for k in members:
                merged[n][j] += z[n][k]
    return merged


def haplotype_posteriors(h_merged, groups, alphabet=ALPHABET):
    """Q-F4 uqs/analyze_results.py:110-118: plain running product over positions."""
    out = []
    for j, (seq, _) in enumerate(groups):
        p = 1
        for pos, base in enumerate(seq):
            p = p * h_merged[j][pos][alphabet.index(base)]
        out.append(p)
    return out


def assignments(z_merged, multiplicity):
    """Q-F5 uqs/analyze_results.py:121-141: a read belongs to every merged cluster that attains
    its row maximum; cluster weight = sum of multiplicities."""
    top = z_merged.max(axis=1, keepdims=True)
    member = z_merged == top
    weight = (member * np.asarray(multiplicity)[:, None]).sum(axis=0)
    return member, weight
