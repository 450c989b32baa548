"""Window-level SNV extraction and the co-occurring-mutations table, from in-memory results.

SURVEY.md §8(f) row 4: the step right after the inference path.  The reference re-parses every
``haplotypes/w-*.reads-support.fas`` it has just written (viloca/shorah_snv.py:190-256 ``parseWindow``,
:486-526 ``get_cooccuring_muts_haplo_df``); here the same tables are built from the haplotypes the
engine returned, without the file round trip.  The file-based forms are kept for drop-in use.

Semantics restated from viloca/shorah_snv.py:95-187 (``_compare_ref_to_read``) and :407-484
(``__mutations_in_haplotype``), pinned by the reference's own vectors
(tests/test_shorah_snv.py:5-38 -> tests/test_host_logic.py) and by tables generated from the
unmodified reference functions on the committed fixture windows (oracle/make_golden.py).

Coordinates: ``start`` is the 1-based window begin; 'X' in the reference marks an inserted column
of the extended-window mode, which does not advance the reported reference position.
"""
from __future__ import annotations

import os
import re
from collections import namedtuple
from dataclasses import dataclass
from typing import Iterable, Iterator, Sequence

from . import fasta

SNP_id = namedtuple("SNP_id", ["pos", "var"])          # shorah_snv.py:65-68


@dataclass
class SNV:                                             # shorah_snv.py:70-78
    chrom: str
    haplotype_id: str
    pos: int
    ref: str
    var: str
    freq: float = 0.0
    support: float = 0.0


@dataclass
class Haplotype:
    """One record of a support file: id, sequence, posterior, ave_reads."""
    id: str
    seq: str
    posterior: float
    ave_reads: float


_HEADER = re.compile(r"posterior=(.*)\s*ave_reads=(.*)")      # shorah_snv.py:224


def _run_length(s: str, begin: int, char: str) -> int:
    end = begin
    while end < len(s) and s[end] == char:
        end += 1
    return end - begin


def _shared_x_before(ref: str, seq: str, last: int) -> int:
    """Number of columns directly before (and including) ``last`` where reference and haplotype both
    hold 'X'; 0 when the run reaches the window start (shorah_snv.py:88-92, including that case)."""
    i = last
    while i >= 0:
        if not (ref[i] == "X" and seq[i] == "X"):
            return last - i
        i -= 1
    return 0


def mark_inserted_gaps(ref: str, seq: str) -> str:
    """A '-' of the haplotype in an inserted column ('X' in the reference) means "no insertion":
    it becomes 'X' so that the column compares equal (shorah_snv.py:94-99)."""
    return "".join("X" if (c == "-" and r == "X") else c for r, c in zip(ref, seq))


@dataclass
class _Event:
    key: SNP_id          # position in window space + variant string (the aggregation key)
    pos: int             # reported position (reference space)
    ref: str
    var: str


def _events(ref: str, seq: str, start: int) -> Iterator[_Event]:
    """Differences between a haplotype and the reference slice, left to right.  Substitutions are
    single columns; a run of '-' in the haplotype (deletion) or of 'X' in the reference (insertion)
    is one event anchored at the preceding column, and further indel columns up to and including
    the column right after the run are not reported again (shorah_snv.py:119-127)."""
    if len(ref) != len(seq):
        raise ValueError("haplotype and reference slice differ in length")
    covered_to = -1
    inserted = 0                                      # 'X' columns of the reference seen so far
    for idx, r in enumerate(ref):
        h = seq[idx]
        if r != h:
            if h == "X":
                raise ValueError("'X' in a haplotype where the reference has a base")
            if h == "-" or r == "X":
                if idx > covered_to:
                    insertion = r == "X"
                    run = _run_length(ref if insertion else seq, idx, "X" if insertion else "-")
                    covered_to = idx + run
                    anchor = idx - 1 - _shared_x_before(ref, seq, idx - 1)    # -1 wraps around, as upstream
                    other = seq if insertion else ref
                    span = other[anchor] + other[idx: idx + run]
                    if insertion:
                        ev_ref, ev_var = ref[anchor], span
                    else:
                        ev_ref, ev_var = span, seq[anchor]
                    yield _Event(SNP_id(pos=start + idx, var=seq[idx:covered_to]), start + idx - 1 - inserted, ev_ref, ev_var)
            else:
                yield _Event(SNP_id(pos=start + idx, var=h), start + idx - inserted, r, h)
        if r == "X":
            inserted += 1


def compare_ref_to_read(ref: str, seq: str, start: int, snp: dict, av: float, post: float, chrom: str,
                        haplotype_id: str) -> int:
    """Adds the SNVs of one haplotype to ``snp`` (freq += ave_reads, support += posterior * ave_reads);
    returns how many it found.  Same contract as shorah_snv.py:101-187."""
    found = 0
    for ev in _events(ref, seq, start):
        found += 1
        hit = snp.get(ev.key)
        if hit is not None:
            hit.freq += av
            hit.support += post * av
        else:
            snp[ev.key] = SNV(chrom, haplotype_id, ev.pos, ev.ref, ev.var, av, post * av)
    return found


def window_snvs(haplotypes: Iterable[Haplotype], ref_slice: str, chrom: str, beg, end, threshold: float = 0.9) -> dict:
    """``parseWindow`` on in-memory haplotypes (shorah_snv.py:190-256): haplotypes below the posterior
    threshold are skipped, the rest contribute their SNVs; finally support /= freq, freq /= reads."""
    snp: dict = {}
    reads = 0.0
    ref_slice = ref_slice.upper()
    for hap in haplotypes:
        if hap.posterior < threshold:
            continue
        reads += hap.ave_reads
        hid = hap.id.split("|")[0] + "-" + str(beg) + "-" + str(end)
        compare_ref_to_read(ref_slice, mark_inserted_gaps(ref_slice, hap.seq.upper()), int(beg), snp, hap.ave_reads,
                            hap.posterior, chrom, hid)
    for v in snp.values():
        v.support /= v.freq
        v.freq /= reads
    return snp


def cooccurring_mutations(haplotypes: Iterable[Haplotype], ref_slice: str, chrom: str, beg, end) -> list:
    """Rows of cooccurring_mutations.csv for one window (shorah_snv.py:407-526): every mutation of
    every haplotype (no posterior filter), a "reference-<beg>-<end>" row for a haplotype equal to the
    reference; every row carries the window's total ``coverage`` (sum of ave_reads)."""
    rows, reads = [], 0.0
    ref_slice = ref_slice.upper()
    start = int(beg)
    for hap in haplotypes:
        reads += hap.ave_reads
        hid = hap.id.split("|")[0] + "-" + str(beg) + "-" + str(end)
        mine = [{"haplotype_id": hid, "chrom": chrom, "start": start, "position": ev.pos, "ref": ev.ref, "var": ev.var,
                 "reads": hap.ave_reads, "support": hap.posterior}
                for ev in _events(ref_slice, hap.seq.upper(), start)]
        if not mine:
            mine = [{"haplotype_id": "reference-" + str(beg) + "-" + str(end), "chrom": chrom, "start": start,
                     "reads": hap.ave_reads, "support": hap.posterior}]
        rows.extend(mine)
    for row in rows:
        row["coverage"] = reads
    return rows


def haplotypes_from_state(state: dict) -> list:
    """The haplotypes ``haplotypes_to_fasta`` would write for a finished window (analyze_results.py:26-53),
    taken from the summary dict instead of the file."""
    out = []
    k = 0
    while "haplotype" + str(k) in state:
        weight = state["weight" + str(k)]
        if weight != 0:
            out.append(Haplotype("haplotype" + str(k), state["haplotype" + str(k)],
                                 float(state["approximatePosterior" + str(k)]), float(weight)))
        k += 1
    return out


def read_support_file(path: str) -> list:
    out = []
    for rec in fasta.parse(path):
        m = _HEADER.search(rec.description)
        out.append(Haplotype(rec.id, rec.seq, float(m.group(1)), float(m.group(2))))
    return out


def parse_window(line: str, extended_window_mode: bool, exclude_non_var_pos_threshold: float, working_dir: str,
                 threshold: float = 0.9) -> dict:
    """File-based drop-in for ``parseWindow`` (same arguments; one line of coverage.txt)."""
    _, chrom, beg, end, _ = line.rstrip().split("\t")
    stem = "w-%s-%s-%s" % (chrom, beg, end)
    kind = "extended-ref" if extended_window_mode else ("envp-full-ref" if exclude_non_var_pos_threshold > 0 else "ref")
    support = os.path.join(working_dir, "haplotypes", stem + ".reads-support.fas")
    ref_file = os.path.join("raw_reads", f"{stem}.{kind}.fas")
    if not (os.path.exists(support) and os.path.exists(ref_file)):
        return {}                                   # shorah_snv.py:245-247: logged and skipped upstream
    refs = {rec.id: rec.seq.upper() for rec in fasta.parse(ref_file)}
    return window_snvs(read_support_file(support), refs[chrom], chrom, beg, end, threshold)
