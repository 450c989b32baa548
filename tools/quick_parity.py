"""Ad-hoc GPU check: teacher-forced and free-running parity of the engine against the oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import _bridge as br
from viloca_b200 import synthetic, engine
import dataclasses

def case(mode, n_raw, L, K, T=6):
    spec = dataclasses.replace(synthetic.WORKLOADS["hiv5k_uqs"], mode=mode, n_clusters=K, called_length=max(8, int(L*0.75)))
    win = synthetic.make_window(spec, 0, n_reads=n_raw, length=L)
    print(f"--- {mode} N_raw={n_raw} N={win.n_reads} L={L} K={K}")
    w, ref, data = br.dense_inputs(win)
    init = br.oracle_init(win)
    step = br.orc.uqs_step if mode == "uqs" else br.orc.lep_step
    terms_fn = br.orc.uqs_elbo_terms if mode == "uqs" else br.orc.lep_elbo_terms
    cur = dict(init)
    with engine.CaviBatch([win]) as b:
        b.upload()
        for it in range(T):
            if it <= 1:
                br.orc._refresh_digamma_sums(mode, cur)
            b.set_state(0, 0, cur["mean_cluster"], cur["mean_log_pi"], br.scalars_of(cur, mode), it)
            nxt = step(data, w, ref, init, cur) if mode == "lep" else step(w, ref, data, init, cur)
            terms = terms_fn(data, w, ref, init, nxt) if mode == "lep" else terms_fn(w, ref, data, init, nxt)
            elbo = sum(terms[1:], terms[0])
            b.step(1)
            out = b.fetch(state=0)[0]
            st = out.state
            zero_same = np.array_equal(st["mean_cluster"] == 0, nxt["mean_cluster"] == 0)
            print(f" it={it} z={br.rel_err(st['mean_cluster'], nxt['mean_cluster']):.2e} zero_pattern={zero_same}"
                  f" h={br.rel_err(st['mean_haplo'], nxt['mean_haplo']):.2e}"
                  f" alpha={br.rel_err(st['alpha'], nxt['alpha']):.2e} mlp={br.rel_err(st['mean_log_pi'], nxt['mean_log_pi']):.2e}"
                  f" a={br.rel_err(st['gamma_a'], nxt['gamma_a']):.2e} b={br.rel_err(st['gamma_b'], nxt['gamma_b']):.2e}"
                  f" elbo={br.rel_err(st['elbo'], elbo):.2e} ({st['elbo']:.6f} vs {elbo:.6f})")
            cur = nxt
        # free run
        b.upload()
        t0 = time.time(); b.run(); t1 = time.time()
        out = b.fetch()[0]
    r = out.restarts[0]
    t2 = time.time(); traj = br.oracle_run(win); t3 = time.time()
    hg = br.orc.haplotype_strings(out.state["mean_haplo"]); ho = br.orc.haplotype_strings(traj.state["mean_haplo"])
    print(f" free-run: gpu n_iter={r.n_iterations} n_upd={r.n_updates} status={r.exit_message!r} elbo={r.elbo:.9f} ({t1-t0:.3f}s)"
          f" | oracle n_iter={traj.n_iterations} n_upd={traj.n_updates} msg={traj.exit_message!r} elbo={traj.elbo:.9f} ({t3-t2:.3f}s)"
          f" | hist_len {len(r.history_elbo)} vs {len(traj.history_elbo)} same_haplotypes={hg == ho}")
    m = min(len(r.history_elbo), len(traj.history_elbo))
    if m:
        print("  max rel elbo-history diff:", br.rel_err(r.history_elbo[:m], np.array(traj.history_elbo[:m])))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "small":
        case("uqs", 120, 40, 10, T=3)
        case("lep", 120, 40, 20, T=3)
        sys.exit(0)
    case("uqs", 300, 60, 10)
    case("lep", 300, 60, 10)
    case("uqs", 600, 201, 100)
    case("lep", 600, 201, 100)
