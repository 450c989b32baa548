"""The oracle (oracle/cavi_oracle.py) against vectors produced by the unmodified reference
(oracle/make_golden.py).  The update equations call the same NumPy primitives in the same
order as the reference, so equality is exact; the full-run checks pin the loop control flow."""
import json
import os

import numpy as np
import pytest


import _bridge as br
from oracle import cavi_oracle as orc


def _json_file(path):
    with open(path) as fh:
        return json.load(fh)



def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=True)


def _init_from_golden(g, mode):
    K = int(g["K"])
    alpha = float(g["alpha0"]) * np.ones(K)
    st = {"alpha": alpha, "mean_log_pi": g["init_mean_log_pi"], "gamma_a": g["init_gamma_a"],
          "gamma_b": g["init_gamma_b"], "mean_log_gamma": tuple(g["init_mean_log_gamma"]),
          "mean_cluster": g["init_mean_cluster"], "mean_haplo": g["init_mean_haplo"],
          "lnB_alpha0": g["init_lnB_alpha0"], "betaln_a0_b0": g["init_betaln_a0_b0"]}
    if mode == "lep":
        st.update({"theta_c": g["init_theta_c"], "theta_d": g["init_theta_d"],
                   "mean_log_theta": tuple(g["init_mean_log_theta"]), "betaln_c0_d0": g["init_betaln_c0_d0"]})
    return st


@pytest.mark.parametrize("mode", ["uqs", "lep"])
def test_update_equations_bit_exact(golden_dir, mode):
    g = _load(golden_dir, f"{mode}_steps.npz")
    # dtypes as the reference pipeline produces them: int8 one-hots in uqs
    # (preparation.py:80,170), float64 in lep; einsum's summation order depends on them
    dt = np.int8 if mode == "uqs" else np.float64
    reads = orc.onehot_from_codes(g["codes"], dtype=dt)
    ref = orc.onehot_from_codes(g["ref_codes"], dtype=dt)
    w = g["weights"]
    if mode == "uqs":
        data = orc.reads_log_error_tensor(g["qualities"], reads)
        assert np.array_equal(data, g["E"])                      # Q-T1
    else:
        data = reads
    init = _init_from_golden(g, mode)
    # draw_init_state itself, from the stored seed
    drawn = orc.draw_init_state(mode, int(g["K"]), float(g["alpha0"]), reads.shape[1], reads.shape[0], ref,
                                np.random.default_rng(int(g["seed"])))
    assert np.array_equal(drawn["mean_cluster"], g["init_mean_cluster"])
    assert drawn["gamma_a"] == g["init_gamma_a"] and drawn["gamma_b"] == g["init_gamma_b"]
    assert np.array_equal(np.array(drawn["mean_log_gamma"]), g["init_mean_log_gamma"])
    assert np.array_equal(drawn["mean_haplo"], g["init_mean_haplo"])
    assert drawn["lnB_alpha0"] == g["init_lnB_alpha0"]
    if mode == "lep":
        assert np.array_equal(np.array(drawn["mean_log_theta"]), g["init_mean_log_theta"])
    cur = dict(init)
    for it in range(len(g["elbo"])):
        if it <= 1:
            orc._refresh_digamma_sums(mode, cur)
        assert cur["digamma_alpha_sum"] == g["in_digamma_alpha_sum"][it]
        assert cur["digamma_a_b_sum"] == g["in_digamma_a_b_sum"][it]
        if mode == "uqs":
            cur = orc.uqs_step(w, ref, data, init, cur)
            terms = orc.uqs_elbo_terms(w, ref, data, init, cur)
        else:
            cur = orc.lep_step(data, w, ref, init, cur)
            terms = orc.lep_elbo_terms(data, w, ref, init, cur)
        for key in ("alpha", "mean_log_pi", "mean_haplo", "mean_cluster"):
            assert np.array_equal(cur[key], g[key][it]), (key, it)
        assert cur["gamma_a"] == g["gamma_a"][it] and cur["gamma_b"] == g["gamma_b"][it]
        assert np.array_equal(np.array(cur["mean_log_gamma"]), g["mean_log_gamma"][it])
        if mode == "lep":
            assert cur["theta_c"] == g["theta_c"][it] and cur["theta_d"] == g["theta_d"][it]
            assert np.array_equal(np.array(cur["mean_log_theta"]), g["mean_log_theta"][it])
        # the reference's uninitialised-memory log (elbo_eqs.py:50-51) makes its entropy terms
        # equal to ours whenever the masked garbage is finite; allow 1e-12 there
        np.testing.assert_allclose(np.array(terms), g["terms"][it], rtol=1e-12, atol=0)
        elbo = terms[0]
        for t in terms[1:]:
            elbo = elbo + t
        np.testing.assert_allclose(elbo, g["elbo"][it], rtol=1e-12)


def _run_case(g, name, seed):
    class W:
        pass
    codes, quals = g[f"{name}_codes"], g[f"{name}_qualities"]
    reads = orc.onehot_from_codes(codes, dtype=np.int8)
    ref = orc.onehot_from_codes(g[f"{name}_ref_codes"], dtype=np.int8)
    E = orc.reads_log_error_tensor(quals, reads)
    K = int(g[f"{name}_K"])
    init = orc.draw_init_state("uqs", K, 1e-4, reads.shape[1], reads.shape[0], ref, np.random.default_rng(seed))
    return orc.run("uqs", g[f"{name}_weights"], ref, E, init, convergence_threshold=float(g[f"{name}_thr"]))


@pytest.mark.parametrize("name,seed", [("a", 27), ("b", 27), ("b", 28), ("b", 29), ("c", 42)])
def test_full_runs_match_reference_run_cavi(golden_dir, name, seed):
    g = _load(golden_dir, "uqs_run.npz")
    traj = _run_case(g, name, seed)
    key = f"{name}_s{seed}"
    assert traj.n_iterations == int(g[key + "_n_iterations"])
    assert traj.exit_message == str(g[key + "_exit_message"])
    assert traj.converged == bool(g[key + "_converged"])
    np.testing.assert_allclose(np.array(traj.history_elbo), g[key + "_history"], rtol=1e-12)
    assert np.array_equal(traj.state["mean_haplo"], g[key + "_mean_haplo"])
    assert np.array_equal(traj.state["mean_cluster"], g[key + "_mean_cluster"])


def test_soft_failure_case_is_pinned(golden_dir):
    """One of the golden runs ends with the reference's 'ELBO is decreasing' exit."""
    g = _load(golden_dir, "uqs_run.npz")
    assert str(g["b_s27_exit_message"]) == orc.MSG_DECREASING
    assert not bool(g["b_s27_converged"])


def test_data4_known_answer(golden_dir):
    """tests/data_4 window, Q = 30, seed 42 -> 54 iterations, ELBO
    -1943.9785808934962 (numpy 2.3.5 / scipy 1.18.1)."""
    d = os.path.join(golden_dir, "data4")
    meta = _json_file(os.path.join(d, "expected.json"))["seed42"]
    assert meta["n_iterations"] == 54 and meta["n_unique"] == 80 and meta["sum_weights"] == 87
    assert meta["elbo"] == pytest.approx(-1943.9785808934962, rel=1e-13)
    from viloca_b200.local_haplotype_inference import _common
    job = _common.WindowJob(mode="uqs", freads_in=os.path.join(d, "w-HXB2-2335-2535.reads.fas"),
                            fref_in=os.path.join(d, "w-HXB2-2335-2535.ref.fas"),
                            fname_qualities=os.path.join(d, "w-HXB2-2335-2535.qualities.npy"),
                            output_dir="./", n_starts=1, K=100, alpha0=1e-4, seed=42)
    # host preparation only (no engine): build the window tensors without touching CUDA
    ids, seqs = _common.read_window_fasta(job.freads_in)
    quals = np.load(job.fname_qualities, allow_pickle=True)
    p = _common.prepare_reads(ids, seqs, quals, "ACGT-", True)
    assert p.codes.shape == (80, 201) and int(p.weights.sum()) == 87
    ref_codes, _, _ = _common.load_reference(job.fref_in, "ACGT-", upper=True)
    reads = orc.onehot_from_codes(p.codes, dtype=np.int8)
    ref = orc.onehot_from_codes(ref_codes, dtype=np.int8)
    E = orc.reads_log_error_tensor(p.qualities, reads)
    init = orc.draw_init_state("uqs", 100, 1e-4, 201, 80, ref, np.random.default_rng(42))
    traj = orc.run("uqs", p.weights, ref, E, init)
    assert traj.n_iterations == 54 and traj.exit_message == orc.MSG_CONVERGED
    np.testing.assert_allclose(np.array(traj.history_elbo), np.array(meta["history_elbo"]), rtol=1e-12)
    groups = orc.unique_haplotypes(traj.state["mean_haplo"])
    assert len(groups) == 26


@pytest.mark.parametrize("name", ["amplicon_90", "sars_24520", "longreads_2900"])
def test_bam_fixture_windows(golden_dir, tmp_path, name):
    """Windows of real reads cut from the reference's BAM fixtures (oracle/bam_windows.py): the
    oracle follows the reference's cavi.run_cavi trajectory — same iteration count and exit, ELBO
    history to 1e-12 (real Phred qualities, deletions, N padding, Q~10 long reads)."""
    from oracle import bam_windows
    from viloca_b200.local_haplotype_inference import _common
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))[name]
    _, f_reads, f_ref, f_q = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(tmp_path))
    ids, seqs = _common.read_window_fasta(f_reads)
    p = _common.prepare_reads(ids, seqs, np.load(f_q, allow_pickle=True), "ACGT-", True)
    assert p.codes.shape == (meta["n_unique"], meta["length"]) and int(p.weights.sum()) == meta["n_raw"]
    ref_codes, _, _ = _common.load_reference(f_ref, "ACGT-", upper=True)
    reads = orc.onehot_from_codes(p.codes, dtype=np.int8)
    ref = orc.onehot_from_codes(ref_codes, dtype=np.int8)
    E = orc.reads_log_error_tensor(p.qualities, reads)
    init = orc.draw_init_state("uqs", meta["K"], 1e-4, meta["length"], meta["n_unique"], ref, np.random.default_rng(meta["seed"]))
    traj = orc.run("uqs", p.weights, ref, E, init)
    assert (traj.n_iterations, traj.exit_message) == (meta["n_iterations"], meta["exit_message"])
    np.testing.assert_allclose(np.array(traj.history_elbo), np.array(meta["history_elbo"]), rtol=1e-12)


def test_init_draw_order(golden_dir):
    g = _load(golden_dir, "init.npz")
    from viloca_b200 import init_state
    ref = orc.onehot_from_codes(np.array([0, 1, 2, 3, 4, 255, 0], dtype=np.uint8))
    for seed in (0, 27, 42):
        for mode in ("uqs", "lep"):
            st = orc.draw_init_state(mode, 9, 1e-4, 7, 13, ref, np.random.default_rng(seed))
            assert np.array_equal(st["mean_cluster"], g[f"{mode}_{seed}_mean_cluster"])
            z, sc = init_state.draw(mode, 9, 1e-4, 13, np.random.default_rng(seed))   # product host code
            assert np.array_equal(z, g[f"{mode}_{seed}_mean_cluster"])
            n = 4 if mode == "uqs" else 8
            assert np.array_equal(sc[:n], g[f"{mode}_{seed}_scalars"])


@pytest.mark.parametrize("name", ["a", "b"])
def test_lep_loop_and_finishing_match_reference(golden_dir, name):
    """learn_error_params end to end (tests/golden/lep_finish, generated by oracle/make_golden.py from the
    unmodified update_eqs / elbo_eqs under the as-written loop and the unmodified
    analyze_results.summarize_results): the oracle's loop takes the same number of updates and leaves
    the same ELBO history; its finishing steps give the same haplotypes, weights and assignments."""
    from viloca_b200.local_haplotype_inference import _common
    d = os.path.join(golden_dir, "lep_finish")
    meta = _json_file(os.path.join(d, "expected.json"))[name]
    ids, seqs = _common.read_window_fasta(os.path.join(d, meta["stem"] + ".reads.fas"))
    p = _common.prepare_reads(ids, seqs, None, "ACGT-", meta["unique"])
    assert p.codes.shape[0] == meta["n_rows"]
    ref_codes, _, _ = _common.load_reference(os.path.join(d, meta["stem"] + ".ref.fas"), "ACGT-", upper=False)
    reads = orc.onehot_from_codes(p.codes)
    ref = orc.onehot_from_codes(ref_codes)
    w = np.asarray(p.weights, dtype=np.float64)
    init = orc.draw_init_state("lep", meta["K"], 1e-4, p.codes.shape[1], p.codes.shape[0], ref,
                               np.random.default_rng(meta["seed"]))
    traj = orc.run("lep", w, ref, reads, init)
    assert (traj.n_updates, traj.n_iterations, traj.exit_message) == (meta["n_updates"], meta["n_iterations"], meta["exit_message"])
    np.testing.assert_allclose(np.array(traj.history_elbo), np.array(meta["history_elbo"]), rtol=1e-13)
    st = traj.state
    g = np.load(os.path.join(d, f"state_{name}.npz"))
    np.testing.assert_allclose(st["mean_cluster"], g["mean_cluster"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(st["mean_haplo"], g["mean_haplo"], rtol=1e-12, atol=0)
    groups = orc.unique_haplotypes(st["mean_haplo"])
    assert [s for s, _ in groups] == meta["haplotypes"]
    merged = orc.merge_clusters(st["mean_cluster"], groups)
    h_m = orc.lep_update_mean_haplo(reads, w, ref, merged, st["mean_log_theta"], st["mean_log_gamma"])
    np.testing.assert_allclose(np.array(orc.haplotype_posteriors(h_m, groups)), np.array(meta["posteriors"]), rtol=1e-11)
    member, weight = orc.assignments(merged, w)
    assert [int(x) for x in weight] == meta["weights"]
    for j in range(len(groups)):
        rows = np.nonzero(member[:, j])[0]
        assert [p.read_ids[n][0] for n in rows] == meta["assigned_unique"][j]
        assert [rid for n in rows for rid in p.read_ids[n]] == meta["assigned_all"][j]


def test_oracle_matches_reference_at_full_amplicon_shape(golden_dir):
    """The oracle at the configs[3] shape (N_raw = 20 000 -> 2 571 unique reads, L = 408, K = 100) against
    the summary of the UNMODIFIED reference's 12 iterations (oracle/make_golden_fullshape.py): the oracle
    is pinned at full size too, not only on the small fixtures."""
    import os
    import sys
    path = os.path.join(golden_dir, "fullshape", "amplicon_full.npz")
    if not os.path.exists(path):
        pytest.skip("fullshape goldens not generated")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _bridge as br
    from oracle import make_golden_fullshape as mg
    g = np.load(path, allow_pickle=True)
    from viloca_b200 import synthetic
    win = synthetic.make_window(synthetic.WORKLOADS["sars_amplicon"], 5, max_iter=4)
    assert (win.n_reads, win.length) == (int(g["n_reads"]), int(g["length"]))
    traj = br.oracle_run(win, max_iterations=4)                      # the first 4 of the 12 (a few seconds)
    assert np.array_equal(np.array(traj.history_elbo), g["r0_history"][:4])      # bit for bit, like the small fixtures
