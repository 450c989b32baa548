"""Host-side logic that needs no GPU: preparation (against vectors from the reference's
preparation.py), FASTA conventions, the longest-first scheduler, synthetic generators."""
import json
import os

import numpy as np
import pytest


from viloca_b200 import fasta, scheduler, synthetic, tensors
from viloca_b200.local_haplotype_inference import _common
from viloca_b200.local_haplotype_inference.learn_error_params import preparation as prep_lep
from viloca_b200.local_haplotype_inference.use_quality_scores import preparation as prep_uqs


def _text(path):
    with open(path) as fh:
        return fh.read()


def _json_file(path):
    with open(path) as fh:
        return json.load(fh)


STEM = "w-REF-1-24"


@pytest.fixture(scope="module")
def prep(golden_dir):
    return np.load(os.path.join(golden_dir, "prep.npz"), allow_pickle=True), os.path.join(golden_dir, "prep_window")


@pytest.mark.parametrize("unique", [True, False])
def test_uqs_preparation_matches_reference(prep, unique):
    g, d = prep
    reads, w, q, mapping = prep_uqs.load_and_process_reads(os.path.join(d, STEM + ".reads.fas"),
                                                           os.path.join(d, STEM + ".qualities.npy"), "ACGT-", unique)
    tag = "unique" if unique else "nonunique"
    assert np.array_equal(reads, g[f"uqs_{tag}_reads"])
    assert np.array_equal(w, g[f"uqs_{tag}_weights"])
    assert np.array_equal(q, g[f"uqs_{tag}_qualities"])          # running-mean qualities, Q==0 -> 2: exact
    E = prep_uqs.compute_reads_log_error_proba(q, reads, 5)
    assert np.array_equal(E, g[f"uqs_{tag}_E"])                   # bit-identical log-error tensor
    if unique:
        assert ["|".join(i for i, _ in v) for v in mapping.values()] == list(g["uqs_unique_ids"])
        assert w.sum() == 40 and reads.shape[0] == 29
    else:
        assert [m[0] for m in mapping] == list(g["uqs_nonunique_ids"])


def test_reference_sequence_loading(prep):
    g, d = prep
    ref, ref_id = prep_uqs.load_reference_seq(os.path.join(d, STEM + ".ref.fas"), "ACGT-")
    assert np.array_equal(ref, g["uqs_ref"]) and ref_id == str(g["uqs_ref_id"])
    seq, _ = prep_lep.load_reference_seq(os.path.join(d, STEM + ".ref.fas"))
    assert np.array_equal(prep_lep.reference2binary(seq, "ACGT-"), g["lep_ref"])
    # reference tests/test_fasta2binary.py:19-63 (hard asserts of the reference's own suite)
    assert np.array_equal(prep_uqs.reference2binary("ACGT-", "ACGT-"), np.eye(5, dtype=np.int8))
    assert np.array_equal(prep_uqs.reference2binary("ANNT", "ACGT-"),
                          np.array([[1, 0, 0, 0, 0], [0] * 5, [0] * 5, [0, 0, 0, 1, 0]], dtype=np.int8))


@pytest.mark.parametrize("unique", [True, False])
def test_lep_preparation_matches_reference(prep, unique):
    g, d = prep
    rl = prep_lep.load_fasta2reads_list(os.path.join(d, STEM + ".reads.fas"), "ACGT-", unique)
    reads, w = prep_lep.reads_list_to_array(rl)
    tag = "unique" if unique else "nonunique"
    assert np.array_equal(reads, g[f"lep_{tag}_reads"])
    assert np.array_equal(w, g[f"lep_{tag}_weights"])
    assert [r.id for r in rl] == list(g[f"lep_{tag}_ids"])
    assert ["|".join(r.identical_reads) for r in rl] == list(g[f"lep_{tag}_identical"])


def test_dedup_weights_and_averaged_qualities():
    """The scenario of the reference's tests/test_fasta2binary.py:119-150: weights [1, 2] and
    averaged qualities."""
    ids, seqs = ["r1", "r2", "r3"], ["ACGT", "TTTT", "TTTT"]
    q = np.array([[40] * 4, [30] * 4, [20] * 4])
    p = _common.prepare_reads(ids, seqs, q, "ACGT-", True)
    assert list(p.weights) == [1, 2]
    assert np.array_equal(p.qualities, np.array([[40.0] * 4, [25.0] * 4]))
    assert p.read_ids == [["r1"], ["r2", "r3"]]
    p = _common.prepare_reads(ids, seqs, q, "ACGT-", False)
    assert list(p.weights) == [1, 1, 1] and p.codes.shape == (3, 4)


def test_synthetic_dedup_equals_host_preparation():
    spec = synthetic.WORKLOADS["hiv5k_uqs"]
    rng = np.random.default_rng(5)
    ref, haps = synthetic.make_haplotypes(rng, 40, 3)
    codes_raw, quals_raw = synthetic.sample_reads(rng, haps, np.array([0.5, 0.3, 0.2]), 200,
                                                  synthetic.WorkloadSpec(**{**spec.__dict__, "called_length": 30}))
    codes, w, q, _ = synthetic.dedup_reads(codes_raw, quals_raw)
    seqs = [tensors.decode(r) for r in codes_raw]
    p = _common.prepare_reads([str(i) for i in range(200)], seqs, quals_raw, "ACGT-", True)
    assert np.array_equal(p.codes, codes) and np.array_equal(p.weights, w) and np.array_equal(p.qualities, q)


def test_fasta_roundtrip_and_header_conventions(tmp_path):
    path = str(tmp_path / "x.fas")
    fasta.write([("haplotype0", "hap_0 | posterior=0.5 ave_reads=3", "A" * 130),
                 ("read7", "|posterior=1", "ACGT")], path)
    with open(path) as fh:
        text = fh.read().split("\n")
    assert text[0] == ">haplotype0 hap_0 | posterior=0.5 ave_reads=3"
    assert [len(x) for x in text[1:4]] == [60, 60, 10]
    assert text[4] == ">read7 |posterior=1"
    recs = list(fasta.parse(path))
    assert recs[0].id == "haplotype0" and recs[0].seq == "A" * 130
    assert recs[1].description == "read7 |posterior=1"


def test_encode_decode():
    codes = tensors.encode("ACGT-NacgtX")
    assert list(codes) == [0, 1, 2, 3, 4, 255, 255, 255, 255, 255, 255]
    assert tensors.decode(codes[:6]) == "ACGT-N"
    oh = tensors.onehot(codes[:6])
    assert oh.shape == (6, 5) and oh[5].sum() == 0 and np.array_equal(tensors.codes_from_onehot(oh), codes[:6])


def test_lpt_partition_is_balanced_and_deterministic():
    costs = [9, 7, 6, 5, 5, 4, 3, 2, 2, 1]
    bins = scheduler.lpt_partition(costs, 3)
    assert sorted(i for b in bins for i in b) == list(range(10))
    loads = [sum(costs[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= 2
    assert bins == scheduler.lpt_partition(costs, 3)
    assert scheduler.lpt_partition([], 2) == [[], []]
    assert scheduler.shard_for_rank(costs, 1, 3) == bins[1]
    with pytest.raises(ValueError):
        scheduler.lpt_partition(costs, 0)


def test_group_haplotypes_sorted_and_first_max_wins():
    h = np.zeros((3, 2, 5))
    h[0, 0, 1] = h[0, 0, 2] = 0.5            # tie -> first maximum ('C')
    h[0, 1, 0] = 1.0
    h[1] = h[0]
    h[2, 0, 0] = 1.0; h[2, 1, 0] = 1.0
    groups = _common.group_haplotypes(h)
    assert groups == [("AA", [2]), ("CA", [0, 1])]


def test_runlist_unpacking_matches_shotgun_run_dpm():
    """The 10-tuple of shotgun.win_to_run (shotgun.py:326) -> window jobs with the file names
    shotgun.run_dpm derives (shotgun.py:151-152)."""
    from viloca_b200 import shotgun_hook
    runlist = [("w-HXB2-1-201.reads.fas", 3000, 0.0001, np.array([42]), "use_quality_scores", 100, 3, True, 1e-3, False),
               ("w-HXB2-68-268.reads.fas", 3000, 0.0001, 7, "learn_error_params", 50, 1, False, 1e-3, False)]
    jobs = shotgun_hook.jobs_from_runlist(runlist)
    assert jobs[0].mode == "uqs" and jobs[0].fref_in == "w-HXB2-1-201.ref.fas"
    assert jobs[0].fname_qualities == "w-HXB2-1-201.qualities.npy" and jobs[0].seed == 42 and jobs[0].n_starts == 3
    assert jobs[0].output_name == "./w-HXB2-1-201.reads-"
    assert jobs[1].mode == "lep" and jobs[1].fname_qualities is None and jobs[1].K == 50 and not jobs[1].unique_modus
    with pytest.raises(ValueError):
        shotgun_hook.jobs_from_runlist([("x.reads.fas", 1, 0.1, 1, "shorah", 10, 1, True, 1e-3, False)])


# ---------------------------------------------------------------------------------------
# window-level SNV extraction
# ---------------------------------------------------------------------------------------

# the reference's own vectors, tests/test_shorah_snv.py:5-38 (start = 1, ave_reads 1, posterior 0.1)
REFERENCE_SNV_VECTORS = [
    ("AACTTA", "AGA--A", {(2, "G"): (2, "A", "G"), (3, "A"): (3, "C", "A"), (4, "--"): (3, "CTT", "A")}),
    ("AA--G", "A---T", {(2, "---"): (1, "AA--", "A"), (5, "T"): (5, "G", "T")}),
    ("AA--X-GC", "AA--X-TC", {(7, "T"): (6, "G", "T")}),
    ("AA-X-GCXGG", "AA-G-TCX-G", {(4, "G"): (3, "-", "-G"), (6, "T"): (5, "G", "T"), (9, "-"): (6, "CG", "C")}),
    ("AXXG", "AGGG", {(2, "GG"): (1, "A", "AGG")}),
    ("ACXXXXGT", "ACXGGCCT", {(4, "GGC"): (2, "C", "CGGC"), (7, "C"): (3, "G", "C")}),
    ("ACXGT", "ACCGT", {(3, "C"): (2, "C", "CC")}),
]


@pytest.mark.parametrize("ref,seq,spec", REFERENCE_SNV_VECTORS)
def test_compare_ref_to_read_reference_vectors(ref, seq, spec):
    from viloca_b200 import snv_windows as sw
    snp = {}
    tot = sw.compare_ref_to_read(ref, seq, 1, snp, 1, 0.1, "HXB2", "some-id")
    expected = {sw.SNP_id(pos=k[0], var=k[1]): sw.SNV("HXB2", "some-id", v[0], v[1], v[2], 1, 0.1) for k, v in spec.items()}
    assert snp == expected and tot == len(snp)


def _snv_rows(snp):
    return [[k.pos, k.var, v.chrom, v.haplotype_id, v.pos, v.ref, v.var, v.freq, v.support] for k, v in sorted(snp.items())]


def test_snv_random_pairs_match_reference_functions(golden_dir):
    """300 random reference / haplotype pairs with deletions and inserted 'X' columns, expected tables
    from the unmodified _compare_ref_to_read / __mutations_in_haplotype (oracle/make_golden.py)."""
    from viloca_b200 import snv_windows as sw
    pairs = _json_file(os.path.join(golden_dir, "snv_windows.json"))["pairs"]
    assert len(pairs) == 300
    for c in pairs:
        pre = sw.mark_inserted_gaps(c["ref"], c["seq"])
        assert pre == c["preprocessed"]
        snp = {}
        tot = [sw.compare_ref_to_read(c["ref"], pre, c["start"], snp, 3.0, 0.95, "chr", "hid"),
               sw.compare_ref_to_read(c["ref"], pre, c["start"], snp, 2.0, 0.5, "chr", "other")]
        assert tot == c["tot"] and _snv_rows(snp) == c["snvs"]
        rows = sw.cooccurring_mutations([sw.Haplotype("hid", c["seq"], 0.95, 3.0)], c["ref"], "chr", c["start"], 0)
        got = [{k: v for k, v in r.items() if k != "coverage"} for r in rows if "position" in r]
        exp = [dict(r, haplotype_id="hid-%d-0" % c["start"]) for r in c["cooccurring"]]
        assert got == exp


@pytest.mark.parametrize("name", ["data4", "hiv_3120", "hiv_2500_3starts", "sars_24520", "amplicon_90", "longreads_2900"])
def test_window_snvs_match_reference_on_fixture_windows(golden_dir, name):
    """parseWindow / get_cooccuring_muts_haplo_df of the reference on the expected support files of the
    fixture windows, against the in-memory functions fed with the same haplotypes."""
    from viloca_b200 import fasta, snv_windows as sw
    g = _json_file(os.path.join(golden_dir, "snv_windows.json"))["windows"][name]
    if name == "data4":
        support = os.path.join(golden_dir, "data4", "expected_seed42.reads-support.fas")
        ref_seq = next(iter(fasta.parse(os.path.join(golden_dir, "data4", "w-HXB2-2335-2535.ref.fas")))).seq
    else:
        support = os.path.join(golden_dir, "bam_windows", name, "expected.reads-support.fas")
        ref_seq = str(np.load(os.path.join(golden_dir, "bam_windows", name, "window.npz"))["ref_seq"])
    haps = sw.read_support_file(support)
    for thr, rows in g["snvs"].items():
        assert _snv_rows(sw.window_snvs(haps, ref_seq, g["chrom"], g["beg"], g["end"], float(thr))) == rows
    got = sw.cooccurring_mutations(haps, ref_seq, g["chrom"], g["beg"], g["end"])
    exp = [{k: v for k, v in r.items() if v is not None} for r in g["cooccurring"]]
    assert got == exp


def test_parallel_job_loading_equals_serial(golden_dir, tmp_path):
    """load_jobs spreads window preparation over worker processes (the reference's Pool over windows);
    the prepared tensors and seeded initial states are identical to loading in-process."""
    from oracle import bam_windows
    names = ["hiv_3120", "hiv_2500_3starts", "sars_24520", "amplicon_90", "longreads_2900"]
    meta = _json_file(os.path.join(golden_dir, "bam_windows", "expected.json"))

    def make_jobs():
        jobs = []
        for k, name in enumerate(names + names[:2]):
            d = tmp_path / f"{k}"
            d.mkdir(exist_ok=True)
            _, f_reads, f_ref, f_q = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(d))
            m = meta[name]
            jobs.append(_common.WindowJob(mode="uqs", freads_in=f_reads, fref_in=f_ref, fname_qualities=f_q, output_dir="./",
                                          n_starts=m["n_starts"], K=m["K"], alpha0=1e-4, seed=m["seed"]))
        return jobs

    serial = _common.load_jobs(make_jobs(), workers=1)
    pooled = _common.load_jobs(make_jobs(), workers=2)
    threaded = _common.load_jobs(make_jobs())                # small windows: the thread pool
    assert len(serial) == len(pooled) == len(threaded) == 7
    for a, b in list(zip(serial, pooled)) + list(zip(serial, threaded)):
        assert a.freads_in.split("/")[-1] == b.freads_in.split("/")[-1]
        for key in ("codes", "weights", "ref_codes", "init_cluster", "init_scalars", "log_hit", "log_off"):
            assert np.array_equal(getattr(a.window, key), getattr(b.window, key)), key
        assert a.prepared.read_ids == b.prepared.read_ids


# ---------------------------------------------------------------------------------------
# native read preparation against the Python restatement of the reference
# ---------------------------------------------------------------------------------------

def _same_prepared(a, b):
    assert a.n_raw == b.n_raw
    assert a.codes.dtype == b.codes.dtype and np.array_equal(a.codes, b.codes)
    assert a.weights.dtype == b.weights.dtype and np.array_equal(a.weights, b.weights)
    assert (a.qualities is None) == (b.qualities is None)
    if a.qualities is not None:
        assert a.qualities.dtype == b.qualities.dtype and np.array_equal(a.qualities, b.qualities)   # bit for bit
    assert a.read_ids == b.read_ids and a.sequences == b.sequences


@pytest.mark.parametrize("seed,n_raw,length,unique,with_q", [(0, 400, 31, True, True), (1, 150, 7, True, True),
                                                             (2, 300, 12, False, True), (3, 500, 9, True, False),
                                                             (4, 1, 5, True, True), (5, 64, 1, True, True)])
def test_native_read_preparation_equals_generic(tmp_path, seed, n_raw, length, unique, with_q):
    """Random windows with many duplicates, N, lower case, gaps, Q = 0, float or int qualities, CRLF."""
    rng = np.random.default_rng(seed)
    pool = ["".join(rng.choice(list("ACGT-Nacgt"), size=length, p=[.2, .2, .2, .2, .05, .05, .025, .025, .025, .025]))
            for _ in range(max(1, n_raw // 6))]
    seqs = [pool[int(rng.integers(len(pool)))] for _ in range(n_raw)]
    path = tmp_path / "w-X-1-9.reads.fas"
    eol = "\r\n" if seed % 2 else "\n"
    with open(path, "w", newline="") as fh:
        for i, s_ in enumerate(seqs):
            fh.write(f">read_{i}  {i % 7} extra words{eol}{s_}{eol}")
    quals = None
    if with_q:
        quals = rng.integers(0, 42, size=(n_raw, length)).astype(np.int64)
        if seed == 1:
            quals = quals.astype(np.float64) + 0.25
    generic = _common.prepare_reads(*_common.read_window_fasta(str(path)), quals, "ACGT-", unique)
    native = _common.prepare_reads_file(str(path), quals, "ACGT-", unique)
    _same_prepared(native, generic)


def test_native_read_preparation_on_reference_fixtures(golden_dir, tmp_path):
    """data_4 and the BAM-cut windows: the native reader gives the reference's prepared tensors."""
    from oracle import bam_windows
    for name in ("hiv_3120", "sars_24520", "amplicon_90", "longreads_2900"):
        _, f_reads, _, f_q = bam_windows.materialize(os.path.join(golden_dir, "bam_windows", name, "window.npz"), str(tmp_path))
        q = np.load(f_q, allow_pickle=True)
        _same_prepared(_common.prepare_reads_file(f_reads, q, "ACGT-", True),
                       _common.prepare_reads(*_common.read_window_fasta(f_reads), q, "ACGT-", True))
    d4 = os.path.join(golden_dir, "data4", "w-HXB2-2335-2535.reads.fas")
    q4 = np.load(os.path.join(golden_dir, "data4", "w-HXB2-2335-2535.qualities.npy"), allow_pickle=True)
    p = _common.prepare_reads_file(d4, q4, "ACGT-", True)
    assert p.codes.shape == (80, 201) and int(p.weights.sum()) == 87


def test_native_read_preparation_falls_back_and_reports(tmp_path):
    wrapped = tmp_path / "wrapped.fas"
    wrapped.write_text(">a 1\nACGT\nAC\n>b 2\nACGT\nAC\n")                 # multi-line records: generic reader
    p = _common.prepare_reads_file(str(wrapped), None, "ACGT-", True)
    assert p.sequences == ["ACGTAC"] and p.read_ids == [["a", "b"]]
    ragged = tmp_path / "ragged.fas"
    ragged.write_text(">a\nACGT\n>b\nACG\n")
    with pytest.raises(ValueError):
        _common.prepare_reads_file(str(ragged), None, "ACGT-", True)


def test_corrected_reads_writer_equals_generic_fasta_writer(tmp_path):
    """correct_reads formats the sequence lines once per haplotype; the file is what the generic
    record writer gives (analyze_results.py:9-23), wrapped lines, empty sequences and odd ids included."""
    rng = np.random.default_rng(3)
    state = {}
    for k, length in enumerate((201, 60, 61, 0, 1)):
        state[f"haplotype{k}"] = "".join(rng.choice(list("ACGT-"), length))
        state[f"assignedReads{k}"] = [f"read_{i}_{k}" for i in range(20)] + ["|posterior=1", "x y"]
    records = [(rid, "|posterior=1", state[f"haplotype{k}"]) for k in range(5) for rid in state[f"assignedReads{k}"]]
    fasta.write(records, str(tmp_path / "a.fas"))
    _common.correct_reads(state, str(tmp_path / "b.fas"))
    assert (tmp_path / "a.fas").read_text() == (tmp_path / "b.fas").read_text()


def test_plan_waves_keeps_order_and_respects_budget():
    import ast
    # plan_waves is pure host logic; engine.py itself imports the CUDA binding, so take the function's source
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "viloca_b200", "engine.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "plan_waves")
    ns = {}
    exec("from typing import Sequence\n" + ast.get_source_segment(src, fn), ns)
    plan = ns["plan_waves"]
    assert plan([5, 5, 5, 20, 1, 1], 10) == [[0, 1], [2], [3], [4, 5]]
    assert plan([], 10) == [] and plan([3], 1) == [[0]]
    sizes = [7, 3, 9, 2, 2, 8, 1]
    waves = plan(sizes, 12)
    assert [i for w in waves for i in w] == list(range(len(sizes)))
    assert all(sum(sizes[i] for i in w) <= 12 or len(w) == 1 for w in waves)


def test_in_memory_window_preparation_equals_file_preparation(tmp_path):
    """A window handed over as arrays (shotgun_hook.jobs_from_arrays) is prepared exactly like the same window
    read back from the files b2w would have written (b2w.py:604-616, 670-672)."""
    from viloca_b200 import shotgun_hook
    rng = np.random.default_rng(5)
    L, n_raw = 40, 60
    base = "".join(rng.choice(list("ACGT"), size=L))
    seqs = []
    for i in range(n_raw):
        s = list(base)
        for p in rng.choice(L, size=rng.integers(0, 3), replace=False):
            s[p] = rng.choice(list("ACGT-N"))
        seqs.append("".join(s))
    ids = [f"read{i} {100 + i}" for i in range(n_raw)]
    quals = rng.integers(0, 42, size=(n_raw, L)).astype(np.int64)
    name = "w-chr1-101-140"
    stem = str(tmp_path / name)
    with open(stem + ".reads.fas", "w") as fh:
        for i, s in zip(ids, seqs):
            fh.write(f">{i}\n{s}\n")
    with open(stem + ".ref.fas", "w") as fh:
        fh.write(f">chr1 101\n{base.lower()}\n")
    np.save(stem + ".qualities.npy", quals, allow_pickle=True)
    for inference, mode in (("use_quality_scores", "uqs"), ("learn_error_params", "lep")):
        run = (stem + ".reads.fas", 0, 1e-4, np.array([27]), inference, 12, 2, True, 1e-3, False)
        from_files = shotgun_hook.jobs_from_runlist([run])[0].load()
        mem = shotgun_hook.jobs_from_arrays([{"name": name, "read_ids": ids, "read_seqs": seqs, "qualities": quals,
                                              "ref": ("chr1", base.lower())}], inference, 12, 2, 1e-4, seed=27)[0].load()
        a, b = from_files.window, mem.window
        assert a.mode == b.mode == mode
        for field in ("codes", "weights", "ref_codes", "init_cluster", "init_scalars", "log_hit", "log_off"):
            x, y = getattr(a, field), getattr(b, field)
            assert (x is None and y is None) or np.array_equal(x, y), field
        assert from_files.prepared.read_ids == mem.prepared.read_ids
        assert from_files.ref_seq == mem.ref_seq
        assert shotgun_hook.window_coordinates(mem) == ("chr1", "101", "140")
        assert mem.write_files is False


def test_cluster_limit_is_reported_before_any_work():
    from viloca_b200 import shotgun_hook
    run = ("w-x-1-10.reads.fas", 0, 1e-4, np.array([27]), "use_quality_scores", 129, 1, True, 1e-3, False)
    with pytest.raises(ValueError, match="128"):
        shotgun_hook.jobs_from_runlist([run])
    with pytest.raises(ValueError, match="128"):
        shotgun_hook.jobs_from_arrays([], "use_quality_scores", 200, 1, 1e-4)


def test_host_memory_estimate_of_an_unloaded_window(tmp_path):
    """run_jobs decides from the read file's size alone whether the windows of a run fit host memory
    together; the estimate must bound what the loaded window really takes."""
    from viloca_b200 import shotgun_hook
    rng = np.random.default_rng(3)
    L, n_raw = 50, 400
    seqs = ["".join(rng.choice(list("ACGT"), size=L)) for _ in range(n_raw)]
    stem = str(tmp_path / "w-c-1-50")
    with open(stem + ".reads.fas", "w") as fh:
        for i, s in enumerate(seqs):
            fh.write(f">read{i} {i}\n{s}\n")
    with open(stem + ".ref.fas", "w") as fh:
        fh.write(f">c 1\n{seqs[0]}\n")
    np.save(stem + ".qualities.npy", rng.integers(2, 41, size=(n_raw, L)).astype(np.int64), allow_pickle=True)
    job = shotgun_hook.jobs_from_runlist([(stem + ".reads.fas", 0, 1e-4, np.array([27]), "use_quality_scores", 20, 3, True, 1e-3, False)])[0]
    est = _common.host_bytes_estimate(job)
    job.load()
    real = job.window.input_bytes() + job.prepared.qualities.nbytes
    assert real <= est <= 3 * real
    assert _common.host_bytes_estimate(job) >= job.window.input_bytes()       # loaded: the real figure
    _common.release_job(job)
    assert job.window is None and job.prepared.qualities is None and job.ref_seq
    assert _common.host_budget_bytes() > (1 << 28)


def test_bench_workload_description_and_lpt_costs():
    """bench.py's host logic: both arms print the same `config`; the longest-first cost follows the committed
    iteration table (profile-guided), the reference arm's genome work is derived from the same table."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    spec = synthetic.WORKLOADS["sars_amplicon"]
    cfg = bench.config_of(spec, spec.n_windows)
    assert cfg["workload"] == "sars_amplicon" and cfg["windows"] == 98 and "l2" in cfg and "model" not in cfg
    assert bench.config_of(spec, spec.n_windows) == cfg                      # one function, both arms
    costs, how = bench.predicted_costs(spec, spec.n_windows)
    assert how.startswith("profile-guided") and len(costs) == 98
    table = bench.load_iteration_table("sars_amplicon")
    assert int(np.argmax(costs)) == int(np.argmax(np.asarray(table["n_updates"]) * np.asarray(table["n_reads"]) * np.asarray(table["length"])))
    parts, _ = bench.shard_windows(spec, spec.n_windows, 8)
    assert sorted(i for p in parts for i in p) == list(range(98))
    heavy = int(np.argmax(costs))
    assert [p for p in parts if heavy in p][0] == [heavy]                     # the 17 178-update window gets a GPU of its own
    work, src = bench.genome_work(spec, spec.n_windows)
    assert work > 0 and "bench_iterations" in src
    costs_u, how_u = bench.predicted_costs(synthetic.WORKLOADS["sars_ont"], 120)
    assert how_u.startswith("uniform") and len(set(costs_u)) == 1
