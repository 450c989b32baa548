"""Real-read window fixtures from the reference's BAM test data (TEST INFRASTRUCTURE).

A small pure-Python BGZF/BAM reader plus a window cutter that follows the conventions of the
reference's window builder (viloca/b2w.py:69-121 CIGAR walk, :440-470 padding):

  * M/=/X copy read bases, D puts '-' with quality 2, I and S are dropped, N (skip) pads '-'
  * a read enters a window if it has at least floor(min_ext * L) aligned positions inside it
    (b2w.py:424-427, 559)
  * positions of the window the read does not reach are 'N' with quality 2 (b2w.py:453-465)
  * window read header ">{query_name} {first_aligned_pos}" (b2w.py:508)

It is NOT a re-implementation of b2w (no insertion handling, no extended windows, no
position filter, no paired-read merging): its only job is to turn the reference's own BAM
fixtures into window files with real bases and real qualities that oracle/make_golden.py feeds
to the unmodified reference `run_dpm_mfa.main`, so that the expected outputs committed under
tests/golden/bam_windows/ belong to exactly the inputs committed next to them.
"""
from __future__ import annotations

import gzip
import struct

import numpy as np

SEQ_CODE = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"


def read_bam(path):
    """-> (references [(name, length)], alignments [dict])."""
    with gzip.open(path, "rb") as fh:          # BGZF = concatenated gzip members
        data = fh.read()
    assert data[:4] == b"BAM\x01", "not a BAM file"
    off = 4
    (l_text,) = struct.unpack_from("<i", data, off)
    off += 4 + l_text
    (n_ref,) = struct.unpack_from("<i", data, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, off)
        name = data[off + 4: off + 4 + l_name - 1].decode()
        (l_ref,) = struct.unpack_from("<i", data, off + 4 + l_name)
        refs.append((name, l_ref))
        off += 8 + l_name
    alns = []
    while off < len(data):
        (block_size,) = struct.unpack_from("<i", data, off)
        start = off + 4
        ref_id, pos, l_read_name, mapq, _bin, n_cigar, flag, l_seq, _nref, _npos, _tlen = struct.unpack_from(
            "<iiBBHHHiiii", data, start)
        p = start + 32
        qname = data[p: p + l_read_name - 1].decode()
        p += l_read_name
        cigar = [(CIGAR_OPS[v & 0xF], v >> 4) for v in struct.unpack_from(f"<{n_cigar}I", data, p)]
        p += 4 * n_cigar
        packed = np.frombuffer(data, dtype=np.uint8, count=(l_seq + 1) // 2, offset=p)
        nib = np.empty(2 * packed.size, dtype=np.uint8)
        nib[0::2] = packed >> 4
        nib[1::2] = packed & 0xF
        seq = "".join(SEQ_CODE[c] for c in nib[:l_seq])
        p += (l_seq + 1) // 2
        qual = np.frombuffer(data, dtype=np.uint8, count=l_seq, offset=p).astype(np.int64)
        alns.append({"qname": qname, "flag": flag, "ref_id": ref_id, "pos": pos, "mapq": mapq,
                     "cigar": cigar, "seq": seq, "qual": qual})
        off = start + block_size
    return refs, alns


def aligned_row(aln):
    """Reference-space bases and qualities of one alignment: (first_pos, str, int64[...])."""
    bases, quals = [], []
    r = 0
    seq, qual = aln["seq"], aln["qual"]
    have_q = qual.size > 0 and not np.all(qual == 255)
    for op, n in aln["cigar"]:
        if op in "M=X":
            bases.append(seq[r: r + n])
            quals.extend(qual[r: r + n] if have_q else [30] * n)
            r += n
        elif op in "DN":
            bases.append("-" * n)
            quals.extend([2] * n)
        elif op in "IS":
            r += n
    return aln["pos"], "".join(bases), np.asarray(quals, dtype=np.int64)


def cut_window(alns, ref_id, start, length, min_ext=0.85, max_reads=None):
    """Rows of the window [start, start+length) (0-based): ids, strings, (n, length) int64."""
    need = int(np.floor(min_ext * length))
    ids, rows, quals = [], [], []
    for aln in alns:
        if aln["ref_id"] != ref_id or aln["flag"] & (0x4 | 0x100 | 0x800) or not aln["cigar"]:
            continue
        first, bases, q = aligned_row(aln)
        lo, hi = max(first, start), min(first + len(bases), start + length)
        if hi - lo < need:
            continue
        row = "N" * (lo - start) + bases[lo - first: hi - first] + "N" * (start + length - hi)
        qq = np.full(length, 2, dtype=np.int64)
        qq[lo - start: hi - start] = q[lo - first: hi - first]
        row = "".join(c if c in "ACGT-" else "N" for c in row.upper())
        ids.append(f"{aln['qname']} {first}")
        rows.append(row)
        quals.append(qq)
        if max_reads is not None and len(rows) >= max_reads:
            break
    return ids, rows, (np.stack(quals) if quals else np.zeros((0, length), dtype=np.int64))


def read_fasta(path):
    name, chunks, out = None, [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if name is not None:
                out.append((name, "".join(chunks)))
            name, chunks = line[1:].split()[0], []
        elif line:
            chunks.append(line)
    if name is not None:
        out.append((name, "".join(chunks)))
    return out


def coverage(alns, ref_id, ref_len):
    cov = np.zeros(ref_len + 1, dtype=np.int64)
    for aln in alns:
        if aln["ref_id"] != ref_id or aln["flag"] & (0x4 | 0x100 | 0x800) or not aln["cigar"]:
            continue
        first, bases, _ = aligned_row(aln)
        cov[first] += 1
        cov[min(first + len(bases), ref_len)] -= 1
    return np.cumsum(cov)[:ref_len]


def save_fixture(path, ref_name, start, ref_seq, ids, rows, quals):
    """Compact committed form of one window (expanded again by `materialize`)."""
    np.savez_compressed(path, ref_name=np.array(ref_name), start=np.array(start), ref_seq=np.array(ref_seq),
                        ids=np.array(ids), rows=np.array(rows), quals=quals.astype(np.uint8))


def materialize(npz_path, out_dir):
    """Write the three files the reference's run_dpm_mfa.main reads for a window
    (b2w.py:508, 604-616: w-{ref}-{beg}-{end}.reads.fas / .ref.fas / .qualities.npy) and
    return (stem, reads.fas, ref.fas, qualities.npy)."""
    import os
    z = np.load(npz_path)
    name, start, ref_seq = str(z["ref_name"]), int(z["start"]), str(z["ref_seq"])
    stem = f"w-{name}-{start + 1}-{start + len(ref_seq)}"
    f_reads = os.path.join(out_dir, stem + ".reads.fas")
    f_ref = os.path.join(out_dir, stem + ".ref.fas")
    f_q = os.path.join(out_dir, stem + ".qualities.npy")
    with open(f_reads, "w") as fh:
        for i, r in zip(z["ids"], z["rows"]):
            fh.write(f">{i}\n{r}\n")
    with open(f_ref, "w") as fh:
        fh.write(f">{name} {start + 1}\n{ref_seq}\n")
    np.save(f_q, z["quals"].astype(np.int64), allow_pickle=True)
    return stem, f_reads, f_ref, f_q
