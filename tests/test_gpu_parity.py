"""Parity of the CUDA path (through the C ABI) with the reference.

Tolerances (BASELINE.json north_star): responsibilities and ELBO within 1e-6 relative per
iteration in FP64, identical haplotype sequences.  The engine is checked at the much tighter
ENG_TOL as well, which is what a pure summation-order difference in FP64 produces.
"""
import dataclasses
import json
import os

import numpy as np
import pytest


import _bridge as br
from oracle import cavi_oracle as orc


def _text(path):
    with open(path) as fh:
        return fh.read()


def _json_file(path):
    with open(path) as fh:
        return json.load(fh)


pytestmark = pytest.mark.gpu

SPEC_TOL = 1e-6     # the stated bar
ENG_TOL = 1e-9      # what we actually hold


def _engine():
    from viloca_b200 import engine
    return engine


def _window_from_golden(g, mode, restarts_init, conv_thres=1e-3, max_iter=5000):
    from viloca_b200 import tensors
    engine = _engine()
    kw = {}
    if mode == "uqs":
        kw["log_hit"], kw["log_off"] = tensors.log_error_terms(g["qualities"])
    zs = np.stack([r[0] for r in restarts_init])
    scs = np.stack([r[1] for r in restarts_init])
    return engine.Window(mode=mode, codes=g["codes"], weights=g["weights"].astype(np.float64),
                         ref_codes=g["ref_codes"], init_cluster=zs, init_scalars=scs,
                         alpha0=float(g["alpha0"]), conv_thres=conv_thres, max_iter=max_iter, **kw)


def _assert_state_close(st, ref_state, mode, tol):
    assert np.array_equal(st["mean_cluster"] == 0, ref_state["mean_cluster"] == 0), "zero pattern of mean_cluster"
    assert br.rel_err(st["mean_cluster"], ref_state["mean_cluster"]) < tol
    assert br.rel_err(st["mean_haplo"], ref_state["mean_haplo"]) < tol
    assert br.rel_err(st["alpha"], ref_state["alpha"]) < tol
    assert br.rel_err(st["mean_log_pi"], ref_state["mean_log_pi"]) < tol
    assert br.rel_err(st["gamma_a"], ref_state["gamma_a"]) < tol
    assert br.rel_err(st["gamma_b"], ref_state["gamma_b"]) < tol
    assert br.rel_err(np.array(st["mean_log_gamma"]), np.array(ref_state["mean_log_gamma"])) < tol
    if mode == "lep":
        assert br.rel_err(st["theta_c"], ref_state["theta_c"]) < tol
        assert br.rel_err(st["theta_d"], ref_state["theta_d"]) < tol
        assert br.rel_err(np.array(st["mean_log_theta"]), np.array(ref_state["mean_log_theta"])) < tol


@pytest.mark.parametrize("mode", ["uqs", "lep"])
def test_teacher_forced_against_reference_vectors(golden_dir, mode):
    """State t of the REFERENCE in -> one device iteration -> compare with the reference's
    state t+1 and its ELBO (terms included)."""
    engine = _engine()
    g = np.load(os.path.join(golden_dir, f"{mode}_steps.npz"), allow_pickle=True)
    sc0 = np.zeros(8)
    sc0[0], sc0[1] = g["init_gamma_a"], g["init_gamma_b"]
    sc0[2:4] = g["init_mean_log_gamma"]
    sc0[4:6] = 1.0
    if mode == "lep":
        sc0[4], sc0[5] = g["init_theta_c"], g["init_theta_d"]
        sc0[6:8] = g["init_mean_log_theta"]
    win = _window_from_golden(g, mode, [(g["init_mean_cluster"], sc0)])
    n_iter = len(g["elbo"])
    with engine.CaviBatch([win]) as b:
        b.upload()
        for it in range(n_iter):
            if it == 0:
                z_in, lp_in = g["init_mean_cluster"], g["init_mean_log_pi"]
                mlg, mlt = g["init_mean_log_gamma"], (g["init_mean_log_theta"] if mode == "lep" else (0, 0))
            else:
                z_in, lp_in = g["mean_cluster"][it - 1], g["mean_log_pi"][it - 1]
                mlg, mlt = g["mean_log_gamma"][it - 1], (g["mean_log_theta"][it - 1] if mode == "lep" else (0, 0))
            sc = np.zeros(16)
            sc[3], sc[4] = mlg
            sc[7], sc[8] = mlt
            sc[9], sc[10] = g["in_digamma_alpha_sum"][it], g["in_digamma_a_b_sum"][it]
            if mode == "lep":
                sc[11] = g["in_digamma_c_d_sum"][it]
            b.set_state(0, 0, z_in, lp_in, sc, it)
            b.step(1)
            st = b.fetch(state=0)[0].state
            ref_state = {k: g[k][it] for k in ("mean_cluster", "mean_haplo", "alpha", "mean_log_pi", "gamma_a",
                                               "gamma_b", "mean_log_gamma")}
            if mode == "lep":
                ref_state.update({k: g[k][it] for k in ("theta_c", "theta_d", "mean_log_theta")})
            _assert_state_close(st, ref_state, mode, SPEC_TOL)
            _assert_state_close(st, ref_state, mode, ENG_TOL)
            assert abs(st["elbo"] - g["elbo"][it]) < ENG_TOL * abs(g["elbo"][it])
            got_terms = np.array(st["elbo_terms"])
            assert br.rel_err(got_terms, g["terms"][it][:4]) < ENG_TOL


def test_free_running_against_reference_run_cavi(golden_dir):
    """Whole trajectories of the reference's cavi.run_cavi: iteration counts, exit messages
    (including the soft failure 'ELBO is decreasing'), ELBO history, final haplotypes; the
    three restarts of case b run as one multi-restart window."""
    from viloca_b200 import init_state, tensors
    engine = _engine()
    g = np.load(os.path.join(golden_dir, "uqs_run.npz"), allow_pickle=True)
    wins, keys = [], []
    for name in ("a", "b", "c"):
        seeds = [int(s) for s in g[f"{name}_seeds"]]
        K = int(g[f"{name}_K"])
        n = g[f"{name}_codes"].shape[0]
        inits = [init_state.draw("uqs", K, 1e-4, n, np.random.default_rng(s)) for s in seeds]
        lh, lo = tensors.log_error_terms(g[f"{name}_qualities"])
        wins.append(engine.Window(mode="uqs", codes=g[f"{name}_codes"], weights=g[f"{name}_weights"].astype(float),
                                  ref_codes=g[f"{name}_ref_codes"], init_cluster=np.stack([i[0] for i in inits]),
                                  init_scalars=np.stack([i[1] for i in inits]), alpha0=1e-4, log_hit=lh, log_off=lo,
                                  conv_thres=float(g[f"{name}_thr"])))
        keys.append([f"{name}_s{s}" for s in seeds])
    with engine.CaviBatch(wins) as b:
        b.upload()
        b.run()
        for wi, ks in enumerate(keys):
            elbos = []
            for r, key in enumerate(ks):
                out = b.fetch(state=[r if i == wi else "none" for i in range(len(wins))])[wi]
                rr = out.restarts[r]
                assert rr.n_iterations == int(g[key + "_n_iterations"]), key
                assert rr.exit_message == str(g[key + "_exit_message"]), key
                assert rr.converged == bool(g[key + "_converged"]), key
                assert br.rel_err(rr.history_elbo, g[key + "_history"]) < ENG_TOL
                assert orc.haplotype_strings(out.state["mean_haplo"]) == orc.haplotype_strings(g[key + "_mean_haplo"])
                assert br.rel_err(out.state["mean_cluster"], g[key + "_mean_cluster"]) < SPEC_TOL
                elbos.append(float(g[key + "_elbo"]))
            # best-run selection of run_dpm_mfa.py:83-95
            assert b.fetch(state="none")[wi].best_restart == int(np.argmax(elbos))


def test_softmax_exponential_matches_host_exp():
    """The kernels' branch-free exp (x <= 0): <= 2 ulp in the normal range, correct rounding
    behaviour in the subnormal range, and exactly 0 where the reference's np.exp underflows."""
    from viloca_b200 import _capi as capi
    rng = np.random.default_rng(5)
    x = np.concatenate([
        -np.abs(rng.standard_cauchy(200000)) * 3, -rng.random(100000) * 760, np.linspace(-746.5, -744.0, 50001),
        np.linspace(-709.0, -707.0, 20001), -np.logspace(-300, 2.9, 20000), [0.0, -0.0, -np.inf, -745.1332191019412,
                                                                              -745.1332191019411, -1e-320, -800.0, -1e6]])
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    capi.check(capi.lib.vl_debug_exp(0, capi.f64_ptr(x), capi.f64_ptr(out), x.size))
    with np.errstate(under="ignore"):
        ref = np.exp(x)
    assert np.array_equal(out == 0, ref == 0)
    normal = ref > 2.3e-308
    ulp = np.spacing(ref[normal])
    assert np.max(np.abs(out[normal] - ref[normal]) / ulp) <= 2.0
    sub = (ref > 0) & ~normal
    assert np.max(np.abs(out[sub] - ref[sub])) <= 2 * 4.94e-324      # within one subnormal step of np.exp
    assert out[np.flatnonzero(x == 0)[0]] == 1.0


def _parse_support(path):
    from viloca_b200 import fasta
    import re
    out = []
    for rec in fasta.parse(path):
        m = re.search(r"posterior=(.*)\s*ave_reads=(.*)", rec.description)   # shorah_snv.py:224-236
        out.append((rec.id, rec.seq, float(m.group(1)), int(m.group(2))))
    return out


def test_data4_window_files_end_to_end(golden_dir, tmp_path):
    """The reference's fixture window through the drop-in entry point: same haplotypes, same
    ave_reads, posteriors equal to 1e-9, identical corrected-reads file."""
    from viloca_b200.local_haplotype_inference.use_quality_scores import run_dpm_mfa
    d = os.path.join(golden_dir, "data4")
    stem = "w-HXB2-2335-2535"
    out_dir = str(tmp_path) + "/"
    run_dpm_mfa.main(os.path.join(d, stem + ".reads.fas"), os.path.join(d, stem + ".ref.fas"),
                     os.path.join(d, stem + ".qualities.npy"), out_dir, 1, 100, 0.0001, seed=42)
    got = _parse_support(out_dir + stem + ".reads-support.fas")
    exp = _parse_support(os.path.join(d, "expected_seed42.reads-support.fas"))
    assert [(i, s, w) for i, s, _, w in got] == [(i, s, w) for i, s, _, w in exp]
    for (_, _, pg, _), (_, _, pe, _) in zip(got, exp):
        assert abs(pg - pe) <= ENG_TOL * abs(pe)
        assert (pg >= 0.9) == (pe >= 0.9)          # the threshold shorah_snv applies (-p 0.9)
    assert _text(out_dir + stem + ".reads-cor.fas") == _text(os.path.join(d, "expected_seed42.reads-cor.fas"))


BAM_WINDOWS = ["hiv_3120", "hiv_2500_3starts", "sars_24520", "amplicon_90", "longreads_2900"]


@pytest.mark.parametrize("name", BAM_WINDOWS)
def test_bam_fixture_windows_end_to_end(golden_dir, tmp_path, name):
    """Windows of real reads with real Phred qualities, deletions and N padding, cut from the
    reference's BAM fixtures tests/data_2,3,6,7 (oracle/bam_windows.py), through the drop-in
    entry point: the files the unmodified reference wrote for the same inputs are reproduced —
    identical haplotype sequences, ids, ave_reads and corrected reads, posteriors to 1e-9."""
    from oracle import bam_windows
    from viloca_b200.local_haplotype_inference.use_quality_scores import run_dpm_mfa
    d = os.path.join(golden_dir, "bam_windows", name)
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))[name]
    stem, f_reads, f_ref, f_q = bam_windows.materialize(os.path.join(d, "window.npz"), str(tmp_path))
    assert stem == meta["stem"]
    out_dir = os.path.join(str(tmp_path), "out") + "/"
    os.makedirs(out_dir)
    run_dpm_mfa.main(f_reads, f_ref, f_q, out_dir, meta["n_starts"], meta["K"], 0.0001, seed=meta["seed"])
    got = _parse_support(out_dir + stem + ".reads-support.fas")
    exp = _parse_support(os.path.join(d, "expected.reads-support.fas"))
    assert [(i, s, w) for i, s, _, w in got] == [(i, s, w) for i, s, _, w in exp]
    for (_, _, pg, _), (_, _, pe, _) in zip(got, exp):
        assert abs(pg - pe) <= ENG_TOL * abs(pe) + 1e-300
        assert (pg >= 0.9) == (pe >= 0.9)
    assert _text(out_dir + stem + ".reads-cor.fas") == _text(os.path.join(d, "expected.reads-cor.fas"))


def test_batched_hook_snv_tables_match_reference(golden_dir, tmp_path, monkeypatch):
    """All fixture windows as one batch through the shotgun hook (shotgun.py:524-538 replacement); the
    SNV tables and co-occurring-mutation rows built from the in-memory results equal what the
    reference's shorah_snv.parseWindow / get_cooccuring_muts_haplo_df give on the reference's files."""
    from oracle import bam_windows
    from viloca_b200 import shotgun_hook
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))
    gold = _json_file(os.path.join(golden_dir, "snv_windows.json"))["windows"]
    monkeypatch.chdir(tmp_path)
    runlist = []
    for name in BAM_WINDOWS:
        m = meta[name]
        stem, f_reads, _, _ = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(tmp_path))
        runlist.append((f_reads, 0, 0.0001, np.array([m["seed"]]), "use_quality_scores", m["K"], m["n_starts"], True, 1e-3, False))
    jobs, finished = shotgun_hook.run_dpm_batch(runlist, devices=[0], collect=True)
    tables = shotgun_hook.snv_tables(jobs, finished, threshold=0.9)
    for name, tab in zip(BAM_WINDOWS, tables):
        g = gold[name]
        assert tab["window"] == (g["chrom"], g["beg"], g["end"])
        got = [[k.pos, k.var, v.chrom, v.haplotype_id, v.pos, v.ref, v.var, v.freq, v.support] for k, v in sorted(tab["snvs"].items())]
        assert [r[:8] for r in got] == [r[:8] for r in g["snvs"]["0.9"]]          # identical SNV calls and frequencies
        for a, e in zip(got, g["snvs"]["0.9"]):
            assert abs(a[8] - e[8]) <= ENG_TOL * abs(e[8])
        exp = [{k: v for k, v in r.items() if v is not None} for r in g["cooccurring"]]
        assert len(tab["cooccurring"]) == len(exp)
        for a, e in zip(tab["cooccurring"], exp):
            assert {k: v for k, v in a.items() if k != "support"} == {k: v for k, v in e.items() if k != "support"}
            assert abs(a["support"] - e["support"]) <= ENG_TOL * abs(e["support"]) + 1e-300
        assert os.path.exists(meta[name]["stem"] + ".reads-support.fas")


@pytest.mark.parametrize("name", [n for n in BAM_WINDOWS if "starts" not in n])
def test_bam_fixture_windows_trajectory(golden_dir, tmp_path, name):
    """Same windows at the numeric seam: iteration count, exit message and the whole ELBO
    history of the reference's cavi.run_cavi."""
    from oracle import bam_windows
    from viloca_b200.local_haplotype_inference import _common
    d = os.path.join(golden_dir, "bam_windows", name)
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))[name]
    _, f_reads, f_ref, f_q = bam_windows.materialize(os.path.join(d, "window.npz"), str(tmp_path))
    job = _common.WindowJob(mode="uqs", freads_in=f_reads, fref_in=f_ref, fname_qualities=f_q, output_dir="./",
                            n_starts=1, K=meta["K"], alpha0=1e-4, seed=meta["seed"])
    job.load()
    assert job.window.n_reads == meta["n_unique"]
    out = _engine().run_windows([job.window])[0]
    r = out.restarts[0]
    assert (r.n_iterations, r.exit_message) == (meta["n_iterations"], meta["exit_message"])
    assert br.rel_err(r.history_elbo, np.array(meta["history_elbo"])) < ENG_TOL


def _spec(mode, K, called):
    from viloca_b200 import synthetic
    base = synthetic.WORKLOADS["hiv5k_uqs"]
    return dataclasses.replace(base, mode=mode, n_clusters=K, called_length=called)


@pytest.mark.parametrize("mode,n_raw,L,K", [("uqs", 400, 201, 100), ("lep", 400, 201, 100), ("uqs", 250, 90, 24),
                                            ("lep", 250, 90, 50), ("uqs", 200, 130, 128)])
def test_free_running_against_oracle(mode, n_raw, L, K):
    from viloca_b200 import synthetic
    engine = _engine()
    win = synthetic.make_window(_spec(mode, K, int(L * 0.75)), 3, n_reads=n_raw, length=L)
    traj = br.oracle_run(win)
    out = engine.run_windows([win])[0]
    r = out.restarts[0]
    assert (r.n_iterations, r.n_updates, r.exit_message) == (traj.n_iterations, traj.n_updates, traj.exit_message)
    assert br.rel_err(r.history_elbo, np.array(traj.history_elbo)) < ENG_TOL
    assert orc.haplotype_strings(out.state["mean_haplo"]) == orc.haplotype_strings(traj.state["mean_haplo"])
    assert np.array_equal(out.state["mean_cluster"] == 0, traj.state["mean_cluster"] == 0), "zero pattern of mean_cluster"
    assert br.rel_err(out.state["mean_cluster"], traj.state["mean_cluster"]) < SPEC_TOL


@pytest.mark.parametrize("mode,n_raw,L,K,R", [("uqs", 400, 201, 100, 1), ("lep", 300, 120, 24, 1), ("uqs", 300, 150, 30, 3),
                                              ("lep", 400, 201, 100, 2), ("uqs", 90, 40, 8, 1)])
def test_small_shapes_and_restarts_against_oracle(mode, n_raw, L, K, R):
    """Small windows (one read tile, K below one slot tile) and several restarts per window: every
    restart follows the reference's trajectory."""
    from viloca_b200 import synthetic
    engine = _engine()
    spec = dataclasses.replace(_spec(mode, K, int(L * 0.75)), n_starts=R)
    win = synthetic.make_window(spec, 5, n_reads=n_raw, length=L)
    with engine.CaviBatch([win]) as b:
        b.upload()
        b.run()
        for r in range(R):
            out = b.fetch(state=r)[0]
            traj = br.oracle_run(win, restart=r)
            rr = out.restarts[r]
            assert (rr.n_iterations, rr.n_updates, rr.exit_message) == (traj.n_iterations, traj.n_updates, traj.exit_message)
            assert br.rel_err(rr.history_elbo, np.array(traj.history_elbo)) < ENG_TOL
            assert orc.haplotype_strings(out.state["mean_haplo"]) == orc.haplotype_strings(traj.state["mean_haplo"])
            assert np.array_equal(out.state["mean_cluster"] == 0, traj.state["mean_cluster"] == 0)
            assert br.rel_err(out.state["mean_cluster"], traj.state["mean_cluster"]) < SPEC_TOL
            assert br.rel_err(out.state["mean_haplo"], traj.state["mean_haplo"]) < SPEC_TOL
            assert br.rel_err(out.state["alpha"], traj.state["alpha"]) < ENG_TOL


def test_fused_chunks_equal_per_update_launches():
    """vl_batch_run's one-launch-per-chunk kernel (persistent workers that claim tiles from per-restart
    scheduling words) does the same arithmetic as two launches per update: bit-identical results."""
    from viloca_b200 import synthetic
    engine = _engine()
    shapes = [("uqs", 400, 201, 100), ("lep", 200, 90, 40), ("uqs", 150, 64, 12), ("uqs", 1, 9, 5), ("lep", 300, 130, 128)]
    wins = [synthetic.make_window(_spec(m, K, max(4, int(L * 0.8))), 20 + i, n_reads=n, length=L)
            for i, (m, n, L, K) in enumerate(shapes)]
    fused = engine.run_windows(wins, fused_chunks=True)
    plain = engine.run_windows(wins, fused_chunks=False)
    for a, b in zip(fused, plain):
        assert a.restarts[0].n_updates == b.restarts[0].n_updates and a.restarts[0].status == b.restarts[0].status
        assert np.array_equal(a.restarts[0].history_elbo, b.restarts[0].history_elbo)
        assert np.array_equal(a.state["mean_cluster"], b.state["mean_cluster"])
        assert np.array_equal(a.state["mean_haplo"], b.state["mean_haplo"])
        assert np.array_equal(a.state["alpha"], b.state["alpha"])


def test_large_alpha0_keeps_every_cluster_alive():
    """alpha0 = 0.5: psi(alpha0) is not very negative, no responsibility underflows, every cluster
    stays in the layout for the whole run (widest kernel class throughout)."""
    from viloca_b200 import synthetic
    engine = _engine()
    spec = dataclasses.replace(_spec("uqs", 100, 60), alpha0=0.5)
    win = synthetic.make_window(spec, 6, n_reads=200, length=80, max_iter=60)
    traj = br.oracle_run(win, max_iterations=60)
    with engine.CaviBatch([win]) as b:
        b.upload()
        b.run()
        out = b.fetch()[0]
        stats = b.problem_stats()
    r = out.restarts[0]
    assert (r.n_updates, r.exit_message) == (traj.n_updates, traj.exit_message)      # a soft failure here: ELBO decreases
    assert br.rel_err(r.history_elbo, np.array(traj.history_elbo)) < ENG_TOL
    assert br.rel_err(out.state["mean_cluster"], traj.state["mean_cluster"]) < SPEC_TOL
    assert (out.state["mean_cluster"] > 0).all() and stats[0, 2] == 100     # 100 clusters in the layout, no ghost


def test_pinned_fetch_equals_pageable_fetch():
    from viloca_b200 import synthetic
    engine = _engine()
    wins = [synthetic.make_window(_spec("uqs", 30, 40), 40 + i, n_reads=150 + 50 * i, length=60) for i in range(3)]
    with engine.CaviBatch(wins) as b:
        b.upload()
        b.run()
        plain = b.fetch()
        pinned = b.fetch(pinned=True)
        again = b.fetch(pinned=True)
        for p, q, r in zip(plain, pinned, again):
            assert q.state["mean_cluster"] is r.state["mean_cluster"] or np.shares_memory(q.state["mean_cluster"], r.state["mean_cluster"])
            for key in ("mean_cluster", "mean_haplo", "alpha"):
                assert np.array_equal(p.state[key], q.state[key])
            assert np.array_equal(p.restarts[0].history_elbo, q.restarts[0].history_elbo)


def test_dead_cluster_skipping_is_exact():
    """Taking the clusters whose responsibilities are exactly 0 out of the contractions
    (VL_OPT_SKIP_DEAD, the default) changes nothing but the summation order: same iteration
    counts and exit states as the all-K computation, ELBO history equal to 1e-12, identical zero
    pattern of mean_cluster."""
    from viloca_b200 import synthetic
    engine = _engine()
    wins = [synthetic.make_window(_spec(m, K, int(L * 0.75)), 11 + i, n_reads=n, length=L)
            for i, (m, n, L, K) in enumerate([("uqs", 400, 201, 100), ("lep", 300, 120, 40), ("uqs", 200, 64, 128)])]
    skip = engine.run_windows(wins, skip_dead=True)
    dense = engine.run_windows(wins, skip_dead=False)
    for a, b in zip(skip, dense):
        ra, rb = a.restarts[0], b.restarts[0]
        assert (ra.n_iterations, ra.n_updates, ra.status) == (rb.n_iterations, rb.n_updates, rb.status)
        assert br.rel_err(ra.history_elbo, rb.history_elbo) < 1e-12
        assert np.array_equal(a.state["mean_cluster"] == 0, b.state["mean_cluster"] == 0)
        assert br.rel_err(a.state["mean_cluster"], b.state["mean_cluster"]) < 1e-9
        assert br.rel_err(a.state["mean_haplo"], b.state["mean_haplo"]) < 1e-9
        assert (a.state["mean_cluster"].sum(axis=0) > 0).sum() < a.state["mean_cluster"].shape[1]   # some cluster died


@pytest.mark.parametrize("name,n_raw,L,K,updates", [("sars_ont", 3000, 300, 100, 14), ("sars_amplicon", 6000, 410, 60, 12)])
def test_long_noisy_windows_against_oracle(name, n_raw, L, K, updates):
    """ONT-like reads (Q ~ 18, 5 % deletions: nearly every (position, base) pair becomes a dense
    column, thousands of sparse entries, no dedup effect) and deep amplicon windows: many read and
    column tiles per restart.  The first updates against the oracle, trajectory and state."""
    from viloca_b200 import synthetic
    engine = _engine()
    spec = dataclasses.replace(synthetic.WORKLOADS[name], n_clusters=K, n_starts=1, length_jitter=0)
    win = synthetic.make_window(spec, 2, n_reads=n_raw, length=L, max_iter=updates)
    assert win.n_reads > 700
    traj = br.oracle_run(win, max_iterations=updates)
    for opts in ({}, {"fused_chunks": False}):
        out = engine.run_windows([win], **opts)[0]
        r = out.restarts[0]
        assert r.n_updates == traj.n_updates == updates and r.status == 4
        assert br.rel_err(r.history_elbo, np.array(traj.history_elbo)) < ENG_TOL
        assert np.array_equal(out.state["mean_cluster"] == 0, traj.state["mean_cluster"] == 0)
        assert br.rel_err(out.state["mean_cluster"], traj.state["mean_cluster"]) < SPEC_TOL
        assert br.rel_err(out.state["mean_haplo"], traj.state["mean_haplo"]) < SPEC_TOL
        assert br.rel_err(out.state["alpha"], traj.state["alpha"]) < ENG_TOL


@pytest.mark.parametrize("n_raw,L,K", [(1400, 100, 24), (4200, 64, 12)])
def test_split_haplotype_tiles_run_to_convergence(n_raw, L, K):
    """Windows of >= 1024 padded reads share the reads of every haplotype tile among 2..8 CTAs (partial
    sums added in split order by the last CTA to arrive).  Whole runs against the oracle, alone and in
    a batch with a small window (bit-identical either way: the split is a property of the window)."""
    from viloca_b200 import synthetic
    engine = _engine()
    spec = dataclasses.replace(synthetic.WORKLOADS["sars_ont"], n_clusters=K, n_starts=1, length_jitter=0)
    win = synthetic.make_window(spec, 5, n_reads=n_raw, length=L)
    assert win.n_reads >= 1024
    small = synthetic.make_window(_spec("uqs", 20, 48), 1, n_reads=200, length=64)
    traj = br.oracle_run(win)
    alone = engine.run_windows([win])[0]
    batch = engine.run_windows([small, win])[1]
    r = alone.restarts[0]
    assert (r.n_iterations, r.n_updates, r.exit_message) == (traj.n_iterations, traj.n_updates, traj.exit_message)
    assert br.rel_err(r.history_elbo, np.array(traj.history_elbo)) < ENG_TOL
    assert orc.haplotype_strings(alone.state["mean_haplo"]) == orc.haplotype_strings(traj.state["mean_haplo"])
    assert br.rel_err(alone.state["mean_cluster"], traj.state["mean_cluster"]) < SPEC_TOL
    assert np.array_equal(alone.restarts[0].history_elbo, batch.restarts[0].history_elbo)
    assert np.array_equal(alone.state["mean_cluster"], batch.state["mean_cluster"])
    assert np.array_equal(alone.state["mean_haplo"], batch.state["mean_haplo"])


def test_ragged_mixed_batch_equals_single_window_runs():
    """Windows of different N, L, K and mode in one batch give bit-identical results to
    running each alone (problems are independent; the reductions are order-fixed)."""
    from viloca_b200 import synthetic
    engine = _engine()
    shapes = [("uqs", 300, 201, 100), ("lep", 120, 57, 10), ("uqs", 40, 16, 100), ("uqs", 500, 333, 30),
              ("lep", 64, 201, 100), ("uqs", 1, 9, 5)]
    wins = [synthetic.make_window(_spec(m, K, max(4, int(L * 0.8))), i, n_reads=n, length=L)
            for i, (m, n, L, K) in enumerate(shapes)]
    together = engine.run_windows(wins)
    for w, o in zip(wins, together):
        alone = engine.run_windows([w])[0]
        assert alone.restarts[0].n_updates == o.restarts[0].n_updates
        assert np.array_equal(alone.restarts[0].history_elbo, o.restarts[0].history_elbo)
        assert np.array_equal(alone.state["mean_cluster"], o.state["mean_cluster"])
        assert np.array_equal(alone.state["mean_haplo"], o.state["mean_haplo"])
    # and the single-read window agrees with the oracle
    traj = br.oracle_run(wins[-1])
    assert together[-1].restarts[0].n_updates == traj.n_updates


def test_edge_cases_all_N_columns_reference_N_and_iteration_cap():
    from viloca_b200 import synthetic
    engine = _engine()
    win = synthetic.make_window(_spec("uqs", 16, 40), 9, n_reads=150, length=50)
    win.codes[:, 7] = 255            # a column no read covers
    win.codes[3, :] = 255            # a read that is all N
    win.ref_codes[11] = 255          # N in the reference: all mass goes to gamma_b (update_eqs.py:103-112)
    win.ref_codes[7] = 255
    traj = br.oracle_run(win)
    out = engine.run_windows([win])[0]
    assert out.restarts[0].n_updates == traj.n_updates
    assert br.rel_err(out.restarts[0].history_elbo, np.array(traj.history_elbo)) < ENG_TOL
    assert br.rel_err(out.state["gamma_b"], traj.state["gamma_b"]) < ENG_TOL
    capped = dataclasses.replace(win, max_iter=5)
    out = engine.run_windows([capped])[0]
    assert out.restarts[0].n_updates == 5 and out.restarts[0].status == 4 and not out.restarts[0].converged
    t5 = br.oracle_run(win, max_iterations=5)
    assert br.rel_err(out.restarts[0].history_elbo, np.array(t5.history_elbo)) < ENG_TOL


def test_finishing_contraction_matches_oracle():
    """analyze_results.py:65-80: merged responsibilities and the haplotype table recomputed
    from them."""
    from viloca_b200 import synthetic
    engine = _engine()
    win = synthetic.make_window(_spec("uqs", 40, 60), 4, n_reads=300, length=80)
    w, ref, E = br.dense_inputs(win)
    with engine.CaviBatch([win]) as b:
        b.upload()
        b.run()
        st = b.fetch()[0].state
        groups = orc.unique_haplotypes(st["mean_haplo"])
        group_of = np.empty(win.n_clusters, dtype=np.int32)
        for j, (_, members) in enumerate(groups):
            group_of[members] = j
        zc, hm = b.merge_haplo(0, 0, group_of, len(groups))
    merged = orc.merge_clusters(st["mean_cluster"], groups)
    assert np.array_equal(zc, merged)                       # same additions in the same order
    h_ref = orc.uqs_update_mean_haplo(w, ref, E, merged, st["mean_log_gamma"])
    assert br.rel_err(hm, h_ref) < ENG_TOL
    post = orc.haplotype_posteriors(hm, groups)
    post_ref = orc.haplotype_posteriors(h_ref, groups)
    assert br.rel_err(np.array(post), np.array(post_ref)) < ENG_TOL


def test_full_size_invariants_and_determinism():
    """BASELINE config 2 at full size (N_raw = 5000, L = 201, K = 100): size-independent
    properties of a CAVI fixed point, and run-to-run bit reproducibility."""
    from viloca_b200 import synthetic
    engine = _engine()
    spec = synthetic.WORKLOADS["hiv5k_uqs"]
    wins = [synthetic.make_window(spec, i) for i in range(2)]
    res1 = engine.run_windows(wins)
    res2 = engine.run_windows(wins)
    for w, o, o2 in zip(wins, res1, res2):
        st, r = o.state, o.restarts[0]
        z, h = st["mean_cluster"], st["mean_haplo"]
        np.testing.assert_allclose(z.sum(axis=1), 1.0, rtol=0, atol=1e-12)
        np.testing.assert_allclose(h.sum(axis=2), 1.0, rtol=0, atol=1e-12)
        assert (z >= 0).all() and (h >= 0).all()
        # sum(alpha) = K*alpha0 + sum(w); a + b = a0 + b0 + K*L  (invariants of update_alpha / update_a_and_b)
        assert abs(st["alpha"].sum() - (w.n_clusters * w.alpha0 + w.weights.sum())) < 1e-8 * w.weights.sum()
        a0, b0 = w.init_scalars[0, 0], w.init_scalars[0, 1]
        assert abs(st["gamma_a"] + st["gamma_b"] - (a0 + b0 + w.n_clusters * w.length)) < 1e-7
        np.testing.assert_allclose(st["alpha"], w.alpha0 + (w.weights[:, None] * z).sum(axis=0), rtol=1e-12)
        hist = r.history_elbo
        assert r.status in (1, 2) and len(hist) == r.n_updates
        if r.status == 1:
            assert (np.diff(hist)[2:] > -1e-5).all()         # coordinate ascent: ELBO never drops by > 1e-5
            assert abs(hist[-1] - hist[-2]) < 1e-3 or r.n_updates == 10
        assert np.array_equal(hist, o2.restarts[0].history_elbo)
        assert np.array_equal(z, o2.state["mean_cluster"]) and np.array_equal(h, o2.state["mean_haplo"])
        # the dominant haplotypes are recovered
        found = set(orc.haplotype_strings(h))
        from viloca_b200 import tensors
        assert tensors.decode(w.true_haplotypes[0]) in found


def test_lep_entry_point_writes_outputs(golden_dir, tmp_path):
    from viloca_b200.local_haplotype_inference.learn_error_params import run_dpm_mfa
    d = os.path.join(golden_dir, "data4")
    stem = "w-HXB2-2335-2535"
    out_dir = str(tmp_path) + "/"
    run_dpm_mfa.main(os.path.join(d, stem + ".reads.fas"), os.path.join(d, stem + ".ref.fas"), out_dir, 2, 100, 0.0001,
                     seed=42)
    sup = _parse_support(out_dir + stem + ".reads-support.fas")
    assert len(sup) >= 1 and sum(w for *_, w in sup) >= 87          # ties may count a read twice
    assert all(len(s) == 201 for _, s, _, _ in sup)
    assert os.path.getsize(out_dir + stem + ".reads-cor.fas") > 0


@pytest.mark.parametrize("name", ["a", "b"])
def test_lep_entry_point_reproduces_reference_files(golden_dir, tmp_path, name):
    """learn_error_params through the drop-in entry point against files written by the UNMODIFIED
    analyze_results.summarize_results / haplotypes_to_fasta / correct_reads (tests/golden/lep_finish:
    the reference's equations under its as-written loop): identical haplotypes, ids, ave_reads and
    corrected reads, posteriors to 1e-9; the device finishing contraction runs in lep mode."""
    from viloca_b200.local_haplotype_inference.learn_error_params import run_dpm_mfa
    d = os.path.join(golden_dir, "lep_finish")
    meta = _json_file(os.path.join(d, "expected.json"))[name]
    stem = meta["stem"]
    out_dir = str(tmp_path) + "/"
    job = dict(freads_in=os.path.join(d, stem + ".reads.fas"), fref_in=os.path.join(d, stem + ".ref.fas"),
               output_dir=out_dir, n_starts=1, K=meta["K"], alpha0=0.0001, unique_modus=meta["unique"], seed=meta["seed"])
    (state, result), = run_dpm_mfa.main_batch([job])
    assert (result["n_iterations"], result["exit_message"]) == (meta["n_iterations"], meta["exit_message"])
    assert br.rel_err(np.array(result["history_elbo"]), np.array(meta["history_elbo"])) < ENG_TOL
    got = _parse_support(out_dir + stem + ".reads-support.fas")
    exp = _parse_support(os.path.join(d, f"expected_{name}.reads-support.fas"))
    assert [(i, s, w) for i, s, _, w in got] == [(i, s, w) for i, s, _, w in exp]
    for (_, _, pg, _), (_, _, pe, _) in zip(got, exp):
        assert abs(pg - pe) <= ENG_TOL * abs(pe) + 1e-300
        assert (pg >= 0.9) == (pe >= 0.9)
    assert _text(out_dir + stem + ".reads-cor.fas") == _text(os.path.join(d, f"expected_{name}.reads-cor.fas"))
    n_hap = len(meta["haplotypes"])
    assert [state[f"haplotype{k}"] for k in range(n_hap)] == meta["haplotypes"]
    assert [state[f"weight{k}"] for k in range(n_hap)] == meta["weights"]
    assert [state[f"assignedUniqueReads{k}"] for k in range(n_hap)] == meta["assigned_unique"]
    assert [state[f"assignedReads{k}"] for k in range(n_hap)] == meta["assigned_all"]
    assert br.rel_err(np.array([state[f"approximatePosterior{k}"] for k in range(n_hap)]), np.array(meta["posteriors"])) < ENG_TOL
    # the same run through the reference-shaped numeric seam (learn_error_params/cavi.py:71-85)
    from viloca_b200.local_haplotype_inference.learn_error_params import cavi, preparation
    reference_seq, _ = preparation.load_reference_seq(job["fref_in"])
    reference_binary = preparation.reference2binary(reference_seq, "ACGT-")
    reads_list = preparation.load_fasta2reads_list(job["freads_in"], "ACGT-", meta["unique"])
    reads_seq_binary, reads_weights = preparation.reads_list_to_array(reads_list)
    init, cur, res = cavi.run_cavi(meta["K"], 0.0001, "ACGT-", reference_binary, reference_seq, reads_list, reads_seq_binary,
                                   reads_weights, 0, out_dir, False, np.random.default_rng(seed=meta["seed"]), meta["seed"])
    assert (res["n_iterations"], res["exit_message"], res["exitflag"]) == (meta["n_iterations"], meta["exit_message"], 0)
    assert [cur[f"haplotype{k}"] for k in range(n_hap)] == meta["haplotypes"]
    assert [res[f"weight{k}"] for k in range(n_hap)] == meta["weights"]
    assert set(init) >= {"alpha", "mean_log_pi", "theta_c", "theta_d", "mean_log_theta", "gamma_a", "gamma_b", "mean_haplo", "mean_cluster"}


def test_reference_signature_run_cavi(golden_dir):
    """cavi.run_cavi with the reference's dense arguments (use_quality_scores/cavi.py:70-84)."""
    from viloca_b200.local_haplotype_inference.use_quality_scores import cavi
    g = np.load(os.path.join(golden_dir, "uqs_run.npz"), allow_pickle=True)
    reads = orc.onehot_from_codes(g["c_codes"], dtype=np.int8)
    ref = orc.onehot_from_codes(g["c_ref_codes"], dtype=np.int8)
    E = orc.reads_log_error_tensor(g["c_qualities"], reads)
    state, res = cavi.run_cavi(int(g["c_K"]), 1e-4, "ACGT-", ref, reads, g["c_weights"], E, 0, "./",
                               float(g["c_thr"]), False, np.random.default_rng(seed=42), 42)
    assert res["n_iterations"] == int(g["c_s42_n_iterations"]) and res["exit_message"] == "ELBO converged."
    assert br.rel_err(np.array(res["history_elbo"]), g["c_s42_history"]) < ENG_TOL
    for key in ("alpha", "mean_log_pi", "gamma_a", "gamma_b", "mean_log_gamma", "mean_haplo", "mean_cluster", "elbo"):
        assert key in state
    assert br.rel_err(state["alpha"], g["c_s42_alpha"]) < ENG_TOL


# ---- full-shape parity: BASELINE.json configs at their real sizes against summaries of the UNMODIFIED
# reference's results (oracle/make_golden_fullshape.py; the windows are regenerated from their seeds) ----

def _fullshape(golden_dir, name):
    path = os.path.join(golden_dir, "fullshape", name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    return np.load(path, allow_pickle=True)


def test_full_hiv5k_window_to_convergence_matches_reference_run_cavi(golden_dir):
    """configs[1] at full size: one whole hiv5k_uqs window through the reference's own cavi.run_cavi to the
    end of its loop (1 500+ iterations) -- identical iteration count and exit message, every ELBO of the
    history within ENG_TOL, identical haplotypes, responsibilities within SPEC_TOL."""
    from viloca_b200 import synthetic
    engine = _engine()
    g = _fullshape(golden_dir, "hiv5k_w96")
    win = synthetic.make_window(synthetic.WORKLOADS["hiv5k_uqs"], int(g["index"]))
    assert win.n_reads == int(g["n_reads"])
    out = engine.run_windows([win])[0]
    r = out.restarts[0]
    assert (r.n_iterations, r.exit_message, r.converged) == (int(g["n_iterations"]), str(g["exit_message"]), bool(g["converged"]))
    assert r.n_updates == len(g["history"]) > 1000
    assert br.rel_err(r.history_elbo, g["history"]) < ENG_TOL
    assert orc.haplotype_strings(out.state["mean_haplo"]) == orc.haplotype_strings(g["mean_haplo"])
    assert np.array_equal(out.state["mean_cluster"] == 0, g["mean_cluster"] == 0)
    assert br.rel_err(out.state["mean_cluster"], g["mean_cluster"]) < SPEC_TOL
    assert br.rel_err(out.state["alpha"], g["alpha"]) < ENG_TOL


@pytest.mark.parametrize("name", ["amplicon_full", "ont_full"])
def test_full_shape_windows_against_reference_updates(golden_dir, name):
    """configs[3] (N_raw = 20 000 -> ~2 500 unique reads, L ~ 400) and configs[4] shape (N = 20 000 unique
    ONT-like reads, L = 1 000, K = 100, 2 restarts: the widest kernel class with 8-way split haplotype tiles):
    the first iterations of every restart against the unmodified update_eqs.update / elbo_eqs.compute_elbo."""
    from oracle import make_golden_fullshape as mg
    engine = _engine()
    g = _fullshape(golden_dir, name)
    win, n_iter = mg.full_shape_windows()[name]
    assert (win.n_reads, win.length, win.n_starts, n_iter) == (int(g["n_reads"]), int(g["length"]), int(g["n_starts"]), int(g["n_iter"]))
    with engine.CaviBatch([win]) as batch:
        batch.upload()
        batch.run()
        for r in range(win.n_starts):
            out = batch.fetch(state=r)[0]
            res, st = out.restarts[r], out.state
            assert res.n_updates == n_iter and res.status == 4                     # stopped by the iteration cap
            assert br.rel_err(res.history_elbo, g[f"r{r}_history"]) < ENG_TOL
            z, h = st["mean_cluster"], st["mean_haplo"]
            assert int((z == 0).sum()) == int(g[f"r{r}_z_zeros"])
            assert br.rel_err(z[mg.sample_index(z.shape[0])], g[f"r{r}_z_rows"]) < SPEC_TOL
            assert br.rel_err(h[:, mg.sample_index(h.shape[1])], g[f"r{r}_h_positions"]) < SPEC_TOL
            assert br.rel_err(z.sum(axis=0), g[f"r{r}_z_colsum"], floor=1e-200) < SPEC_TOL
            assert br.rel_err(st["alpha"], g[f"r{r}_alpha"]) < ENG_TOL
            assert br.rel_err(st["mean_log_pi"], g[f"r{r}_mean_log_pi"]) < ENG_TOL
            assert br.rel_err(np.array([st["gamma_a"], st["gamma_b"]]), g[f"r{r}_gamma_ab"]) < ENG_TOL


def test_batched_hook_on_two_devices(golden_dir, tmp_path, monkeypatch):
    """shotgun_hook.run_dpm_batch with the windows sharded over two GPUs (one host thread per device):
    the same files and states as on one device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from oracle import bam_windows
    from viloca_b200 import shotgun_hook
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))
    runs = {}
    for tag, devices in (("one", [0]), ("two", [0, 1])):
        d = tmp_path / tag
        d.mkdir()
        monkeypatch.chdir(d)
        runlist = []
        for name in BAM_WINDOWS:
            m = meta[name]
            _, f_reads, _, _ = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(d))
            runlist.append((f_reads, 0, 0.0001, np.array([m["seed"]]), "use_quality_scores", m["K"], m["n_starts"], True, 1e-3, False))
        jobs, finished = shotgun_hook.run_dpm_batch(runlist, devices=devices, collect=True)
        runs[tag] = (d, [meta[n]["stem"] for n in BAM_WINDOWS], finished)
    (d1, stems, fin1), (d2, _, fin2) = runs["one"], runs["two"]
    for stem, (s1, r1), (s2, r2) in zip(stems, fin1, fin2):
        for suffix in (".reads-support.fas", ".reads-cor.fas"):
            assert _text(os.path.join(d1, stem + suffix)) == _text(os.path.join(d2, stem + suffix))
        assert r1["n_iterations"] == r2["n_iterations"] and r1["elbo"] == r2["elbo"]


def test_waves_equal_one_batch():
    """HBM-bounded waves (next wave uploaded on its own stream while this one runs) give bit for bit what
    one batch of all windows gives; a window larger than the budget is a wave of its own."""
    from viloca_b200 import synthetic
    engine = _engine()
    wins = [synthetic.make_window(_spec("uqs", 24, 60), i, n_reads=150 + 40 * i, length=80) for i in range(5)]
    wins.append(synthetic.make_window(_spec("lep", 16, 50), 7, n_reads=200, length=64))
    one = engine.run_windows(wins)
    sizes = [engine.window_workspace_bytes(w) for w in wins]
    stats = {}
    seen = []
    waves = engine.run_in_waves(wins, budget_bytes=int(1.5 * max(sizes)), stats=stats,
                                consume=lambda batch, idx, outs: seen.append((list(idx), len(batch.windows))))
    assert stats["waves"] >= 3 and [i for idx, _ in seen for i in idx] == list(range(len(wins)))
    assert all(n == len(idx) for idx, n in seen)
    for a, b in zip(one, waves):
        assert a.best_restart == b.best_restart
        assert [r.n_updates for r in a.restarts] == [r.n_updates for r in b.restarts]
        assert np.array_equal(a.restarts[0].history_elbo, b.restarts[0].history_elbo)
        assert np.array_equal(a.state["mean_cluster"], b.state["mean_cluster"])
        assert np.array_equal(a.state["mean_haplo"], b.state["mean_haplo"])
    tiny = engine.run_in_waves(wins[:2], budget_bytes=1)          # every window exceeds the budget: one wave each
    assert np.array_equal(tiny[1].state["mean_cluster"], one[1].state["mean_cluster"])


def test_in_memory_windows_through_the_batched_hook(golden_dir, tmp_path, monkeypatch):
    """the fixture windows handed to run_dpm_batch as arrays (no w-*.reads.fas / .qualities.npy
    round trip, results kept in memory) give the same states and the same SNV tables as the file-based path."""
    from oracle import bam_windows
    from viloca_b200 import shotgun_hook
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))
    monkeypatch.chdir(tmp_path)
    names = [n for n in BAM_WINDOWS if meta[n]["n_starts"] == 1 and meta[n]["K"] == meta[BAM_WINDOWS[0]]["K"]] or BAM_WINDOWS[:1]
    K, seed = meta[names[0]]["K"], meta[names[0]]["seed"]
    names = [n for n in names if meta[n]["seed"] == seed]
    runlist, arrays = [], []
    for name in names:
        npz = os.path.join(golden_dir, "bam_windows", name, "window.npz")
        stem, f_reads, _, _ = bam_windows.materialize(npz, str(tmp_path))
        runlist.append((f_reads, 0, 0.0001, np.array([seed]), "use_quality_scores", K, 1, True, 1e-3, False))
        z = np.load(npz)
        start, ref_seq = int(z["start"]), str(z["ref_seq"])
        arrays.append({"name": stem, "read_ids": [str(i) for i in z["ids"]], "read_seqs": [str(r) for r in z["rows"]],
                       "qualities": z["quals"].astype(np.int64), "ref": (str(z["ref_name"]), ref_seq)})
    jobs_f, fin_f = shotgun_hook.run_dpm_batch(runlist, devices=[0], collect=True)
    before = set(os.listdir(tmp_path))
    jobs_m = shotgun_hook.jobs_from_arrays(arrays, "use_quality_scores", K, 1, 0.0001, seed=seed)
    jobs_m, fin_m = shotgun_hook.run_dpm_batch(jobs_m, devices=[0], collect=True)
    assert set(os.listdir(tmp_path)) == before                        # nothing written
    for (sf, rf), (sm, rm) in zip(fin_f, fin_m):
        assert rf["n_iterations"] == rm["n_iterations"] and rf["elbo"] == rm["elbo"]
        assert np.array_equal(sf["mean_cluster"], sm["mean_cluster"])
        assert [sf[k] for k in sorted(sf) if k.startswith("haplotype")] == [sm[k] for k in sorted(sm) if k.startswith("haplotype")]
    tf, tm = shotgun_hook.snv_tables(jobs_f, fin_f), shotgun_hook.snv_tables(jobs_m, fin_m)
    assert [t["window"] for t in tf] == [t["window"] for t in tm]
    for a, b in zip(tf, tm):
        assert sorted((k.pos, k.var, v.freq, v.support) for k, v in a["snvs"].items()) == \
               sorted((k.pos, k.var, v.freq, v.support) for k, v in b["snvs"].items())
        assert a["cooccurring"] == b["cooccurring"]


def test_run_jobs_in_host_bounded_groups(golden_dir, tmp_path, monkeypatch):
    """A run whose windows do not fit host memory together is loaded, run and released group by group
    (forced here with a 1-byte budget): the same files and states as the all-at-once path."""
    from oracle import bam_windows
    from viloca_b200 import shotgun_hook
    from viloca_b200.local_haplotype_inference import _common
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))
    out = {}
    for tag, budget in (("all", None), ("groups", 1)):
        d = tmp_path / tag
        d.mkdir()
        monkeypatch.chdir(d)
        runlist = []
        for name in BAM_WINDOWS[:3]:
            m = meta[name]
            _, f_reads, _, _ = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(d))
            runlist.append((f_reads, 0, 0.0001, np.array([m["seed"]]), "use_quality_scores", m["K"], m["n_starts"], True, 1e-3, False))
        jobs = shotgun_hook.jobs_from_runlist(runlist)
        fin = _common.run_jobs(jobs, device=0, host_budget=budget)
        out[tag] = (d, [meta[n]["stem"] for n in BAM_WINDOWS[:3]], fin, jobs)
    (d1, stems, f1, _), (d2, _, f2, jobs2) = out["all"], out["groups"]
    assert all(j.window is None and j.ref_seq for j in jobs2)                    # released, but still usable for snv_tables
    for stem, (s1, r1), (s2, r2) in zip(stems, f1, f2):
        for suffix in (".reads-support.fas", ".reads-cor.fas"):
            assert _text(os.path.join(d1, stem + suffix)) == _text(os.path.join(d2, stem + suffix))
        assert r1["elbo"] == r2["elbo"] and np.array_equal(s1["mean_cluster"], s2["mean_cluster"])
    assert len(shotgun_hook.snv_tables(jobs2, f2)) == 3
