"""Generate tests/golden/* from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists; the resulting fixtures are
committed so that the GPU box (which has no reference) can check the oracle and the CUDA path
against them.  The reference is imported, never copied; Biopython (absent here) is replaced
by the shim in oracle/_bioshim for the FASTA wrappers only.

    python oracle/make_golden.py            # rewrites tests/golden/

Fixtures:
  uqs_steps.npz / lep_steps.npz   per-iteration states + ELBO terms of the reference's
                                  update_eqs.update / elbo_eqs.compute_elbo, driven exactly as
                                  cavi.run_cavi drives them (cavi.py:120-169)
  uqs_run.npz                     full cavi.run_cavi trajectories (history_elbo, n_iterations,
                                  exit_message, final state) incl. a multi-restart case
  init.npz                        initialization.draw_init_state outputs for fixed seeds
  prep.npz (+ prep_window/)       preparation.load_and_process_reads /
                                  compute_reads_log_error_proba / load_reference_seq outputs
  data4/                          the reference's own fixture window tests/data_4 run through
                                  run_dpm_mfa.main (support.fas, cor.fas, ELBO history)
  snv_windows.json                window-level SNV tables and co-occurring-mutation rows of the
                                  reference's shorah_snv.parseWindow / get_cooccuring_muts_haplo_df
                                  on the expected support files above, plus random reference /
                                  haplotype pairs (deletions, inserted 'X' columns) through
                                  _compare_ref_to_read and __mutations_in_haplotype
  lep_finish/                     a learn_error_params window run to the end of its loop (the reference's
                                  update_eqs.update / elbo_eqs.compute_elbo under the as-written control
                                  flow of learn_error_params/cavi.py:135-199; the shipped driver raises) and
                                  finished by the UNMODIFIED analyze_results.summarize_results /
                                  haplotypes_to_fasta / correct_reads (learn_error_params/analyze_results.py)
  bam_windows/                    windows with real bases and qualities cut from the reference's
                                  BAM fixtures tests/data_2,3,6,7 (oracle/bam_windows.py) and the
                                  files run_dpm_mfa.main writes for them
"""
from __future__ import annotations

import collections
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VILOCA_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, os.path.join(HERE, "_bioshim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from numpy.random import default_rng  # noqa: E402
from scipy.special import betaln, digamma  # noqa: E402
from scipy.stats._multivariate import _lnB as lnB  # noqa: E402


def _onehot(codes, dtype):
    out = np.zeros(codes.shape + (5,), dtype=dtype)
    ok = codes < 5
    out[np.nonzero(ok) + (codes[ok].astype(np.int64),)] = 1
    return out


def small_window(seed, n_raw, length, n_hap=3, called=None):
    """Seeded toy window (own generator; the inputs are stored in the fixture)."""
    rng = default_rng(seed)
    ref = rng.integers(0, 4, size=length).astype(np.uint8)
    haps = np.tile(ref, (n_hap, 1))
    for h in range(1, n_hap):
        pos = rng.choice(length, size=max(2, length // 12), replace=False)
        haps[h, pos] = (ref[pos] + rng.integers(1, 4, size=pos.size)) % 4
    src = rng.choice(n_hap, size=n_raw, p=np.array([0.6, 0.3, 0.1])[:n_hap] / np.sum([0.6, 0.3, 0.1][:n_hap]))
    codes = haps[src].copy()
    quals = rng.integers(0, 42, size=(n_raw, length)).astype(np.int64)   # includes Q = 0
    err = rng.random((n_raw, length)) < 0.02
    codes[err] = (codes[err] + 1) % 4
    codes[rng.random((n_raw, length)) < 0.02] = 4            # '-'
    if called is not None:
        start = rng.integers(0, length - called + 1, size=n_raw)
        col = np.arange(length)[None, :]
        codes[(col < start[:, None]) | (col >= (start + called)[:, None])] = 255
    ref_n = ref.copy()
    ref_n[length // 2] = 255                                   # one 'N' in the reference
    return codes, quals, ref_n


def write_window_files(dirname, stem, codes, quals, ref_codes):
    table = np.array(list("ACGT-") + ["N"])
    to_str = lambda c: "".join(table[np.where(c < 5, c, 5)])
    with open(os.path.join(dirname, stem + ".reads.fas"), "w") as fh:
        for i, row in enumerate(codes):
            fh.write(f">read{i} {100 + i}\n{to_str(row)}\n")
    np.save(os.path.join(dirname, stem + ".qualities.npy"), quals, allow_pickle=True)
    with open(os.path.join(dirname, stem + ".ref.fas"), "w") as fh:
        fh.write(f">REF {1}\n{to_str(ref_codes)}\n")


def drive_steps(mode, n_iter, K, alpha0, seed, reads_bin, weights, ref_bin, E):
    """The exact per-iteration sequence of cavi.run_cavi (cavi.py:95-169), recording every
    state and every ELBO term."""
    if mode == "uqs":
        from viloca.local_haplotype_inference.use_quality_scores import elbo_eqs, initialization, update_eqs
        st0 = initialization.draw_init_state(K, alpha0, "ACGT-", reads_bin.shape[1], reads_bin.shape[0], ref_bin,
                                             default_rng(seed))
    else:
        from viloca.local_haplotype_inference.learn_error_params import elbo_eqs, initialization, update_eqs

        class _R:   # draw_init_state only looks at len(reads_list) and reads_list[0].seq_binary.shape[0]
            def __init__(self, b):
                self.seq_binary = b
        reads_list = [_R(r) for r in reads_bin]
        st0 = initialization.draw_init_state(K, alpha0, "ACGT-", reads_list, ref_bin, default_rng(seed))
    st0.update({"lnB_alpha0": lnB(st0["alpha"]), "betaln_a0_b0": betaln(st0["gamma_a"], st0["gamma_b"])})
    if mode == "lep":
        st0["betaln_c0_d0"] = betaln(st0["theta_c"], st0["theta_d"])
    rec = collections.defaultdict(list)
    cur = st0
    for it in range(n_iter):
        if it <= 1:
            cur["digamma_alpha_sum"] = digamma(cur["alpha"].sum(axis=0))
            cur["digamma_a_b_sum"] = digamma(cur["gamma_a"] + cur["gamma_b"])
            if mode == "lep":
                cur["digamma_c_d_sum"] = digamma(cur["theta_c"] + cur["theta_d"])
        rec["in_digamma_alpha_sum"].append(cur["digamma_alpha_sum"])
        rec["in_digamma_a_b_sum"].append(cur["digamma_a_b_sum"])
        if mode == "lep":
            rec["in_digamma_c_d_sum"].append(cur["digamma_c_d_sum"])
        if mode == "uqs":
            cur = update_eqs.update(reads_bin, weights, ref_bin, E, st0, cur)
            terms = (
                elbo_eqs.elbo_data(weights, cur["mean_cluster"], cur["mean_haplo"], E),
                elbo_eqs.elbo_pi(st0["alpha"], st0["lnB_alpha0"], cur["alpha"], cur["mean_log_pi"]),
                elbo_eqs.elbo_cluster(cur["mean_cluster"], cur["mean_log_pi"], weights),
                elbo_eqs.elbo_haplo(ref_bin, cur["mean_haplo"], cur["mean_log_gamma"]),
                elbo_eqs.elbo_gamma(st0["gamma_a"], st0["gamma_b"], st0["betaln_a0_b0"], cur["mean_log_gamma"],
                                    cur["gamma_a"], cur["gamma_b"]),
            )
            elbo = elbo_eqs.compute_elbo(weights, ref_bin, E, st0, cur)
        else:
            cur = update_eqs.update(reads_bin, weights, None, ref_bin, st0, cur)
            terms = (
                elbo_eqs.elbo_data(weights, reads_bin, cur["mean_cluster"], cur["mean_haplo"], cur["mean_log_theta"]),
                elbo_eqs.elbo_pi(st0["alpha"], st0["lnB_alpha0"], cur["alpha"], cur["mean_log_pi"]),
                elbo_eqs.elbo_cluster(cur["mean_cluster"], cur["mean_log_pi"], weights),
                elbo_eqs.elbo_haplo(ref_bin, cur["mean_haplo"], cur["mean_log_gamma"]),
                elbo_eqs.elbo_gamma(st0["gamma_a"], st0["gamma_b"], st0["betaln_a0_b0"], cur["mean_log_gamma"],
                                    cur["gamma_a"], cur["gamma_b"]),
                elbo_eqs.elbo_gamma(st0["theta_c"], st0["theta_d"], st0["betaln_c0_d0"], cur["mean_log_theta"],
                                    cur["theta_c"], cur["theta_d"]),
            )
            elbo = elbo_eqs.compute_elbo(weights, reads_bin, ref_bin, st0, cur)
        for key in ("alpha", "mean_log_pi", "gamma_a", "gamma_b", "mean_haplo", "mean_cluster"):
            rec[key].append(np.array(cur[key]))
        rec["mean_log_gamma"].append(np.array(cur["mean_log_gamma"]))
        if mode == "lep":
            rec["theta_c"].append(cur["theta_c"]); rec["theta_d"].append(cur["theta_d"])
            rec["mean_log_theta"].append(np.array(cur["mean_log_theta"]))
        rec["terms"].append(np.array(terms))
        rec["elbo"].append(elbo)
    init = {"init_mean_cluster": st0["mean_cluster"], "init_mean_haplo": st0["mean_haplo"],
            "init_gamma_a": st0["gamma_a"], "init_gamma_b": st0["gamma_b"],
            "init_mean_log_gamma": np.array(st0["mean_log_gamma"]), "init_mean_log_pi": st0["mean_log_pi"],
            "init_lnB_alpha0": st0["lnB_alpha0"], "init_betaln_a0_b0": st0["betaln_a0_b0"]}
    if mode == "lep":
        init.update({"init_theta_c": st0["theta_c"], "init_theta_d": st0["theta_d"],
                     "init_mean_log_theta": np.array(st0["mean_log_theta"]),
                     "init_betaln_c0_d0": st0["betaln_c0_d0"]})
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(init)
    return out


def golden_steps():
    from viloca.local_haplotype_inference.use_quality_scores import preparation as prep_uqs
    for mode in ("uqs", "lep"):
        K, alpha0, seed, n_iter = 12, 1e-4, 5, 8
        codes, quals, ref = small_window(11, 48, 37, called=30)
        # non-unique rows are fine here; give them varied integer weights instead
        weights = default_rng(3).integers(1, 6, size=codes.shape[0])
        quals = np.where(quals == 0, 2, quals).astype(np.float64)
        ref_bin = _onehot(ref, np.int8 if mode == "uqs" else np.float64)
        if mode == "uqs":
            reads_bin = _onehot(codes, np.int8)
            E = prep_uqs.compute_reads_log_error_proba(quals, reads_bin, 5)
        else:
            reads_bin = _onehot(codes, np.float64)
            E = None
        out = drive_steps(mode, n_iter, K, alpha0, seed, reads_bin, weights, ref_bin, E)
        out.update({"codes": codes, "qualities": quals, "weights": weights, "ref_codes": ref,
                    "K": K, "alpha0": alpha0, "seed": seed})
        if mode == "uqs":
            out["E"] = E
        np.savez_compressed(os.path.join(GOLD, f"{mode}_steps.npz"), **out)
        print(f"{mode}_steps: elbo[0..2] = {out['elbo'][:3]}")


def golden_runs():
    """Full reference trajectories through cavi.run_cavi (uqs; the lep driver is broken as
    shipped)."""
    from viloca.local_haplotype_inference.use_quality_scores import cavi, preparation as prep_uqs
    cases = {}
    for name, (wseed, n_raw, length, K, seeds, thr) in {
        "a": (21, 60, 45, 10, [27], 1e-3),
        "b": (22, 90, 64, 16, [27, 28, 29], 1e-3),      # restarts: seed + s, cavi.py:43
        "c": (23, 40, 33, 8, [42], 1e-6),
    }.items():
        codes, quals, ref = small_window(wseed, n_raw, length, called=int(length * 0.8))
        weights = default_rng(wseed).integers(1, 4, size=n_raw)
        quals = np.where(quals == 0, 2, quals).astype(np.float64)
        reads_bin = _onehot(codes, np.int8)
        ref_bin = _onehot(ref, np.int8)
        E = prep_uqs.compute_reads_log_error_proba(quals, reads_bin, 5)
        for s in seeds:
            with contextlib.redirect_stdout(io.StringIO()):
                state, res = cavi.run_cavi(K, 1e-4, "ACGT-", ref_bin, reads_bin, weights, E, s - seeds[0], "./",
                                           thr, False, default_rng(seed=s), s)
            key = f"{name}_s{s}"
            cases[key + "_history"] = np.array(res["history_elbo"])
            cases[key + "_n_iterations"] = res["n_iterations"]
            cases[key + "_converged"] = res["converged"]
            cases[key + "_exit_message"] = res["exit_message"]
            cases[key + "_elbo"] = res["elbo"]
            cases[key + "_mean_haplo"] = state["mean_haplo"]
            cases[key + "_mean_cluster"] = state["mean_cluster"]
            cases[key + "_alpha"] = state["alpha"]
            print(f"run {key}: {res['n_iterations']} iterations, {res['exit_message']}, elbo {res['elbo']}")
        cases[name + "_codes"] = codes; cases[name + "_qualities"] = quals
        cases[name + "_weights"] = weights; cases[name + "_ref_codes"] = ref
        cases[name + "_K"] = K; cases[name + "_seeds"] = np.array(seeds); cases[name + "_thr"] = thr
    np.savez_compressed(os.path.join(GOLD, "uqs_run.npz"), **cases)


def golden_init():
    from viloca.local_haplotype_inference.use_quality_scores import initialization as init_uqs
    from viloca.local_haplotype_inference.learn_error_params import initialization as init_lep
    out = {}
    ref_bin = _onehot(np.array([0, 1, 2, 3, 4, 255, 0], dtype=np.uint8), np.int8)
    for seed in (0, 27, 42):
        st = init_uqs.draw_init_state(9, 1e-4, "ACGT-", 7, 13, ref_bin, default_rng(seed))
        out[f"uqs_{seed}_mean_cluster"] = st["mean_cluster"]
        out[f"uqs_{seed}_scalars"] = np.array([st["gamma_a"], st["gamma_b"], *st["mean_log_gamma"]])
        out[f"uqs_{seed}_mean_log_pi"] = st["mean_log_pi"]
        out[f"uqs_{seed}_mean_haplo"] = st["mean_haplo"]

        class _R:
            seq_binary = np.zeros((7, 5))
        st = init_lep.draw_init_state(9, 1e-4, "ACGT-", [_R()] * 13, ref_bin.astype(float), default_rng(seed))
        out[f"lep_{seed}_mean_cluster"] = st["mean_cluster"]
        out[f"lep_{seed}_scalars"] = np.array([st["gamma_a"], st["gamma_b"], *st["mean_log_gamma"],
                                               st["theta_c"], st["theta_d"], *st["mean_log_theta"]])
    np.savez_compressed(os.path.join(GOLD, "init.npz"), **out)


def golden_prep():
    from viloca.local_haplotype_inference.use_quality_scores import preparation as prep_uqs
    from viloca.local_haplotype_inference.learn_error_params import preparation as prep_lep
    d = os.path.join(GOLD, "prep_window")
    os.makedirs(d, exist_ok=True)
    codes, quals, ref = small_window(31, 40, 24, n_hap=2, called=20)
    codes[5] = codes[0]; codes[9] = codes[0]; codes[17] = codes[3]      # force duplicates
    quals[0, :4] = 0
    stem = "w-REF-1-24"
    write_window_files(d, stem, codes, quals, ref)
    # lower-case reference bases must be upper-cased by load_reference_seq (preparation.py:165)
    with open(os.path.join(d, stem + ".ref.fas")) as fh:
        head, seq = fh.read().split("\n")[:2]
    with open(os.path.join(d, stem + ".ref.fas"), "w") as fh:
        fh.write(head + "\n" + seq[:6].lower() + seq[6:] + "\n")
    out = {}
    for unique in (True, False):
        rb, w, q, mapping = prep_uqs.load_and_process_reads(os.path.join(d, stem + ".reads.fas"),
                                                            os.path.join(d, stem + ".qualities.npy"), "ACGT-", unique)
        tag = "unique" if unique else "nonunique"
        out[f"uqs_{tag}_reads"] = np.asarray(rb)
        out[f"uqs_{tag}_weights"] = np.asarray(w)
        out[f"uqs_{tag}_qualities"] = np.asarray(q)
        out[f"uqs_{tag}_E"] = prep_uqs.compute_reads_log_error_proba(np.asarray(q, dtype=float), np.asarray(rb), 5)
        if unique:
            out["uqs_unique_ids"] = np.array(["|".join(i for i, _ in v) for v in mapping.values()])
        else:
            out["uqs_nonunique_ids"] = np.array([m[0] for m in mapping])
    ref_bin, ref_id = prep_uqs.load_reference_seq(os.path.join(d, stem + ".ref.fas"), "ACGT-")
    out["uqs_ref"] = ref_bin
    out["uqs_ref_id"] = ref_id
    for unique in (True, False):
        rl = prep_lep.load_fasta2reads_list(os.path.join(d, stem + ".reads.fas"), "ACGT-", unique)
        rb, w = prep_lep.reads_list_to_array(rl)
        tag = "unique" if unique else "nonunique"
        out[f"lep_{tag}_reads"] = rb
        out[f"lep_{tag}_weights"] = w
        out[f"lep_{tag}_ids"] = np.array([r.id for r in rl])
        out[f"lep_{tag}_identical"] = np.array(["|".join(r.identical_reads) for r in rl])
    rseq, rid = prep_lep.load_reference_seq(os.path.join(d, stem + ".ref.fas"))
    out["lep_ref"] = prep_lep.reference2binary(rseq, "ACGT-")
    np.savez_compressed(os.path.join(GOLD, "prep.npz"), **out)
    print("prep: unique rows", out["uqs_unique_reads"].shape[0], "of", codes.shape[0])


def golden_data4():
    """The reference's fixture window tests/data_4 through run_dpm_mfa.main.  The window has no
    reference/qualities files; SURVEY.md §8(c) defines them: reference = per-column most common
    non-N character (ties: first seen), qualities = 30 everywhere, seed 42."""
    from viloca.local_haplotype_inference.use_quality_scores import run_dpm_mfa, cavi, preparation
    d = os.path.join(GOLD, "data4")
    os.makedirs(d, exist_ok=True)
    stem = "w-HXB2-2335-2535"
    src = os.path.join(REF, "tests", "data_4", stem + ".reads.fas")
    shutil.copy(src, os.path.join(d, stem + ".reads.fas"))     # fixture DATA (not source code)
    os.chmod(os.path.join(d, stem + ".reads.fas"), 0o644)
    seqs = [line.strip() for line in open(src) if not line.startswith(">")]
    L = len(seqs[0])
    ref = []
    for j in range(L):
        cnt = collections.OrderedDict()
        for s in seqs:
            if s[j] != "N":
                cnt[s[j]] = cnt.get(s[j], 0) + 1
        top = max(cnt.values())
        ref.append([c for c, v in cnt.items() if v == top][0])
    with open(os.path.join(d, stem + ".ref.fas"), "w") as fh:
        fh.write(">HXB2 2335\n" + "".join(ref) + "\n")
    np.save(os.path.join(d, stem + ".qualities.npy"), np.full((len(seqs), L), 30, dtype=np.int64), allow_pickle=True)
    meta = {}
    for mode_seed, n_starts in ((42, 1),):
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    run_dpm_mfa.main(os.path.join(d, stem + ".reads.fas"), os.path.join(d, stem + ".ref.fas"),
                                     os.path.join(d, stem + ".qualities.npy"), tmp + "/", n_starts, 100, 0.0001,
                                     seed=mode_seed)
                for suffix in ("support.fas", "cor.fas"):
                    shutil.copy(os.path.join(tmp, f"{stem}.reads-{suffix}"),
                                os.path.join(d, f"expected_seed{mode_seed}.reads-{suffix}"))
            finally:
                os.chdir(cwd)
        ref_bin, _ = preparation.load_reference_seq(os.path.join(d, stem + ".ref.fas"), "ACGT-")
        rb, w, q, mapping = preparation.load_and_process_reads(os.path.join(d, stem + ".reads.fas"),
                                                               os.path.join(d, stem + ".qualities.npy"), "ACGT-", True)
        E = preparation.compute_reads_log_error_proba(q, rb, 5)
        with contextlib.redirect_stdout(io.StringIO()):
            state, res = cavi.run_cavi(100, 0.0001, "ACGT-", ref_bin, rb, w, E, 0, "./", 1e-3, False,
                                       default_rng(seed=mode_seed), mode_seed)
        meta[f"seed{mode_seed}"] = {"n_iterations": int(res["n_iterations"]), "exit_message": res["exit_message"],
                                    "elbo": float(res["elbo"]), "history_elbo": [float(x) for x in res["history_elbo"]],
                                    "n_unique": int(rb.shape[0]), "sum_weights": int(np.sum(w))}
        print("data4:", meta[f"seed{mode_seed}"]["n_iterations"], "iterations, elbo", meta[f"seed{mode_seed}"]["elbo"])
    import numpy, scipy
    meta["versions"] = {"numpy": numpy.__version__, "scipy": scipy.__version__}
    with open(os.path.join(d, "expected.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


def golden_lep_finish():
    """learn_error_params end to end.  The shipped lep driver raises (TypeError / NameError, SURVEY.md
    section 8(a) L-C1), so the loop is driven here exactly as learn_error_params/cavi.py:135-199 is
    written (minus the appends to undefined lists), with the reference's own update_eqs.update and
    elbo_eqs.compute_elbo; the finishing steps and both writers are the unmodified reference functions."""
    from viloca.local_haplotype_inference.learn_error_params import analyze_results, elbo_eqs, initialization
    from viloca.local_haplotype_inference.learn_error_params import preparation, update_eqs
    d = os.path.join(GOLD, "lep_finish")
    os.makedirs(d, exist_ok=True)
    meta = {}
    for name, (wseed, n_raw, length, called, K, seed, unique) in {
        "a": (41, 70, 40, 34, 12, 27, True),
        "b": (43, 50, 32, None, 8, 5, False),
    }.items():
        codes, quals, ref = small_window(wseed, n_raw, length, called=called)
        codes[7] = codes[2]; codes[11] = codes[2]; codes[30] = codes[4]; codes[31] = codes[4]      # duplicates
        stem = f"w-LEP{name}-1-{length}"
        write_window_files(d, stem, codes, quals, ref)
        os.remove(os.path.join(d, stem + ".qualities.npy"))          # the mode reads no qualities
        f_reads, f_ref = os.path.join(d, stem + ".reads.fas"), os.path.join(d, stem + ".ref.fas")
        reference_seq, _ = preparation.load_reference_seq(f_ref)
        reference_binary = preparation.reference2binary(reference_seq, "ACGT-")
        reads_list = preparation.load_fasta2reads_list(f_reads, "ACGT-", unique)
        reads_seq_binary, reads_weights = preparation.reads_list_to_array(reads_list)
        st0 = initialization.draw_init_state(K, 1e-4, "ACGT-", reads_list, reference_binary, default_rng(seed=seed))
        st0.update({"lnB_alpha0": lnB(st0["alpha"]), "betaln_a0_b0": betaln(st0["gamma_a"], st0["gamma_b"]),
                    "betaln_c0_d0": betaln(st0["theta_c"], st0["theta_d"])})
        it, n_updates, converged, message, history, cur = 0, 0, False, "", [], st0
        while (converged is False) or (it < 10):
            if it <= 1:
                cur.update({"digamma_alpha_sum": digamma(cur["alpha"].sum(axis=0)),
                            "digamma_a_b_sum": digamma(cur["gamma_a"] + cur["gamma_b"]),
                            "digamma_c_d_sum": digamma(cur["theta_c"] + cur["theta_d"])})
            cur = update_eqs.update(reads_seq_binary, reads_weights, reads_list, reference_binary, st0, cur)
            elbo = elbo_eqs.compute_elbo(reads_weights, reads_seq_binary, reference_binary, st0, cur)
            n_updates += 1
            if it % 2 == 0:
                history.append(elbo)
            if it > 1:
                if np.isnan(elbo):
                    message = "Error: ELBO is nan."
                    break
                elif (history[-2] > elbo) and np.abs(elbo - history[-2]) > 1e-08:
                    message = "Error: ELBO is decreasing."
                    break
                elif np.abs(elbo - history[-2]) < 1e-03:
                    converged = True
                    it += 1
                    message = "ELBO converged."
                else:
                    it = 0
            cur.update({"elbo": elbo})
            it += 1
        cur.update({"elbo": elbo})
        summary = analyze_results.summarize_results(cur, "ACGT-", reads_seq_binary, reads_weights, reads_list,
                                                    reference_binary, reference_seq)
        cur.update(summary)
        analyze_results.haplotypes_to_fasta(cur, os.path.join(d, f"expected_{name}.reads-support.fas"))
        analyze_results.correct_reads(cur, os.path.join(d, f"expected_{name}.reads-cor.fas"))
        n_hap = len([k for k in summary if k.startswith("haplotype")])
        meta[name] = {
            "stem": stem, "K": K, "seed": seed, "unique": unique, "n_rows": len(reads_list),
            "n_updates": n_updates, "n_iterations": int(it), "exit_message": message, "elbo": float(elbo),
            "history_elbo": [float(x) for x in history],
            "haplotypes": [summary[f"haplotype{k}"] for k in range(n_hap)],
            "posteriors": [float(summary[f"approximatePosterior{k}"]) for k in range(n_hap)],
            "weights": [int(summary[f"weight{k}"]) for k in range(n_hap)],
            "assigned_unique": [list(summary[f"assignedUniqueReads{k}"]) for k in range(n_hap)],
            "assigned_all": [list(summary[f"assignedReads{k}"]) for k in range(n_hap)],
        }
        np.savez_compressed(os.path.join(d, f"state_{name}.npz"), mean_cluster=cur["mean_cluster"],
                            mean_haplo=cur["mean_haplo"], mean_log_theta=np.array(cur["mean_log_theta"]),
                            mean_log_gamma=np.array(cur["mean_log_gamma"]), alpha=cur["alpha"])
        print(f"lep_finish {name}: {n_updates} updates, {message} elbo {elbo}, {n_hap} haplotypes, weights {meta[name]['weights']}")
    import numpy, scipy
    meta["versions"] = {"numpy": numpy.__version__, "scipy": scipy.__version__}
    with open(os.path.join(d, "expected.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


BAM_WINDOWS = [
    # name, BAM, reference FASTA, start (0-based), length, max reads, n_starts, K, seed
    ("hiv_3120", "data_2/REF_aln.bam", "data_2/cohort_consensus.fasta", 3120, 201, 300, 1, 100, 42),
    ("hiv_2500_3starts", "data_2/REF_aln.bam", "data_2/cohort_consensus.fasta", 2500, 201, 300, 3, 100, 7),
    ("sars_24520", "data_3/REF_aln.bam", "data_3/NC_045512.2.fasta", 24520, 201, 300, 1, 100, 42),
    ("amplicon_90", "data_6/reads.shotgun.bam", "data_6/reference.fasta", 0, 90, None, 1, 100, 42),
    ("longreads_2900", "data_7/reads.shotgun.bam", "data_7/reference.fasta", 2900, 1000, None, 1, 100, 42),
]


def golden_bam_windows():
    """Windows of real reads (real Phred qualities, deletions, N padding) from the reference's
    BAM fixtures through the unmodified run_dpm_mfa.main (use_quality_scores)."""
    import time
    import bam_windows as bw
    from viloca.local_haplotype_inference.use_quality_scores import run_dpm_mfa, cavi, preparation
    top = os.path.join(GOLD, "bam_windows")
    os.makedirs(top, exist_ok=True)
    meta = {}
    for name, bam, fasta, start, length, max_reads, n_starts, K, seed in BAM_WINDOWS:
        refs, alns = bw.read_bam(os.path.join(REF, "tests", bam))
        ref_name, ref_seq = bw.read_fasta(os.path.join(REF, "tests", fasta))[0]
        assert refs[0][0] == ref_name
        ids, rows, quals = bw.cut_window(alns, 0, start, length, max_reads=max_reads)
        d = os.path.join(top, name)
        os.makedirs(d, exist_ok=True)
        bw.save_fixture(os.path.join(d, "window.npz"), ref_name, start, ref_seq[start:start + length].upper(), ids, rows, quals)
        t0 = time.time()
        with tempfile.TemporaryDirectory() as tmp:
            stem, f_reads, f_ref, f_q = bw.materialize(os.path.join(d, "window.npz"), tmp)
            out = os.path.join(tmp, "out") + "/"
            os.makedirs(out)
            with contextlib.redirect_stdout(io.StringIO()):
                run_dpm_mfa.main(f_reads, f_ref, f_q, out, n_starts, K, 0.0001, seed=seed)
            for suffix in ("support.fas", "cor.fas"):
                shutil.copy(os.path.join(out, f"{stem}.reads-{suffix}"), os.path.join(d, f"expected.reads-{suffix}"))
            ref_bin, _ = preparation.load_reference_seq(f_ref, "ACGT-")
            rb, w, q, _ = preparation.load_and_process_reads(f_reads, f_q, "ACGT-", True)
        entry = {"stem": stem, "n_raw": len(rows), "n_unique": int(rb.shape[0]), "length": length, "n_starts": n_starts,
                 "K": K, "seed": seed, "mean_quality": float(quals[quals > 2].mean()),
                 "frac_gap": float(np.mean([r.count("-") for r in rows]) / length),
                 "frac_N": float(np.mean([r.count("N") for r in rows]) / length)}
        if n_starts == 1:
            E = preparation.compute_reads_log_error_proba(q, rb, 5)
            with contextlib.redirect_stdout(io.StringIO()):
                _, res = cavi.run_cavi(K, 0.0001, "ACGT-", ref_bin, rb, w, E, 0, "./", 1e-3, False,
                                       default_rng(seed=seed), seed)
            entry.update(n_iterations=int(res["n_iterations"]), exit_message=res["exit_message"],
                         elbo=float(res["elbo"]), history_elbo=[float(x) for x in res["history_elbo"]])
        meta[name] = entry
        print("bam window", name, {k: v for k, v in entry.items() if k != "history_elbo"}, f"{time.time() - t0:.1f}s")
    import numpy, scipy
    meta["versions"] = {"numpy": numpy.__version__, "scipy": scipy.__version__}
    with open(os.path.join(top, "expected.json"), "w") as fh:
        json.dump(meta, fh, indent=1)


def _snv_rows(snp):
    return [[k.pos, k.var, v.chrom, v.haplotype_id, v.pos, v.ref, v.var, v.freq, v.support] for k, v in sorted(snp.items())]


def _cooc_rows(df):
    rows = []
    for rec in df.to_dict("records"):
        rows.append({k: (None if (isinstance(v, float) and v != v) else (v.item() if hasattr(v, "item") else v))
                     for k, v in rec.items()})
    return rows


def golden_snv():
    """Post-inference window parsing of the unmodified reference (viloca/shorah_snv.py); libshorah (the
    C++ strand-bias filter, not used by these functions) is stubbed so that the module imports."""
    import types
    sys.modules.setdefault("libshorah", types.ModuleType("libshorah"))
    from viloca import shorah_snv as ss
    mutations_in_haplotype = ss.__dict__["__mutations_in_haplotype"]
    out = {"windows": {}, "pairs": []}
    fixtures = [("data4", os.path.join(GOLD, "data4", "expected_seed42.reads-support.fas"),
                 os.path.join(GOLD, "data4", "w-HXB2-2335-2535.ref.fas"), "HXB2", "2335", "2535")]
    meta = json.load(open(os.path.join(GOLD, "bam_windows", "expected.json")))
    import bam_windows as bw
    for name, entry in meta.items():
        if name == "versions":
            continue
        _, chrom, beg, end = entry["stem"].rsplit("-", 3)[0:1] + entry["stem"].split("-", 1)[1].rsplit("-", 2)
        fixtures.append((name, os.path.join(GOLD, "bam_windows", name, "expected.reads-support.fas"), None, chrom, beg, end))
    for name, support, ref_file, chrom, beg, end in fixtures:
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                os.makedirs("haplotypes"); os.makedirs("raw_reads")
                stem = f"w-{chrom}-{beg}-{end}"
                shutil.copy(support, os.path.join("haplotypes", stem + ".reads-support.fas"))
                if ref_file is None:
                    _, _, ref_file, _ = bw.materialize(os.path.join(GOLD, "bam_windows", name, "window.npz"), tmp)
                shutil.copy(ref_file, os.path.join("raw_reads", stem + ".ref.fas"))
                line = f"{stem}.reads.fas\t{chrom}\t{beg}\t{end}\t100\n"
                per = {}
                for thr in (0.9, 0.0):
                    per[str(thr)] = _snv_rows(ss.parseWindow(line, False, 0, tmp, thr))
                cooc = ss.get_cooccuring_muts_haplo_df(os.path.join("haplotypes", stem + ".reads-support.fas"),
                                                       os.path.join("raw_reads", stem + ".ref.fas"), beg, end, chrom)
                out["windows"][name] = {"chrom": chrom, "beg": beg, "end": end, "snvs": per, "cooccurring": _cooc_rows(cooc)}
                print("snv golden", name, {k: len(v) for k, v in per.items()}, "cooccurring rows", len(cooc))
            finally:
                os.chdir(cwd)
    rng = default_rng(5)
    while len(out["pairs"]) < 300:
        n = int(rng.integers(4, 24))
        ref = "".join(rng.choice(list("ACGT-X"), size=n, p=[0.22, 0.22, 0.22, 0.22, 0.04, 0.08]))
        seq = []
        for r in ref:
            u = rng.random()
            if r == "X":
                seq.append(str(rng.choice(list("ACGTX-"), p=[0.1, 0.1, 0.1, 0.1, 0.5, 0.1])))
            else:
                seq.append(r if u < 0.7 else str(rng.choice(list("ACGT-"), p=[0.2, 0.2, 0.2, 0.2, 0.2])))
        seq = "".join(seq)
        start = int(rng.integers(1, 500))
        try:
            snp = {}
            pre = ss._preprocess_seq_with_X(ref, seq)
            tot = ss._compare_ref_to_read(ref, pre, start, snp, 3.0, 0.95, "chr", "hid")
            tot2 = ss._compare_ref_to_read(ref, pre, start, snp, 2.0, 0.5, "chr", "other")     # aggregation path
            cooc = mutations_in_haplotype(ref, seq, start, 3.0, 0.95, "chr", "hid")
        except (IndexError, AssertionError):
            continue
        out["pairs"].append({"ref": ref, "seq": seq, "preprocessed": pre, "start": start, "tot": [tot, tot2],
                             "snvs": _snv_rows(snp), "cooccurring": cooc})
    with open(os.path.join(GOLD, "snv_windows.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: goldens can only be regenerated where the reference is mounted")
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp()
    os.chdir(work)          # importing run_dpm_mfa creates viloca.log in the CWD
    golden_steps()
    golden_runs()
    golden_init()
    golden_prep()
    golden_data4()
    golden_lep_finish()
    golden_bam_windows()
    golden_snv()
    print("wrote", GOLD)
