"""Window files -> tensors, with the return types of the reference's
use_quality_scores/preparation.py (dense one-hot arrays) for callers that want them.  The
engine itself consumes the compact form (``_common.prepare_reads``)."""
from __future__ import annotations

import numpy as np

from ... import tensors
from .. import _common


def reference2binary(reference_seq, alphabet):
    """preparation.py:167-174: (L, B) int8 one-hot, zero row outside the alphabet."""
    return tensors.onehot(tensors.encode(str(reference_seq), alphabet), len(alphabet), np.int8)


def load_reference_seq(reference_file, alphabet):
    """preparation.py:163-165: first record, upper-cased -> (reference_binary, id)."""
    codes, ref_id, _ = _common.load_reference(reference_file, alphabet, upper=True)
    return tensors.onehot(codes, len(alphabet), np.int8), ref_id


def load_and_process_reads(fname_fasta, fname_qualities, alphabet, unique_modus):
    """preparation.py:18-90 -> (reads_seq_binary, reads_weights, qualities, read_mapping)."""
    with open(fname_qualities, "rb") as fh:
        qualities = np.load(fh, allow_pickle=True)
    ids, seqs = _common.read_window_fasta(fname_fasta)
    p = _common.prepare_reads(ids, seqs, qualities, alphabet, unique_modus)
    if unique_modus:
        reads = tensors.onehot(p.codes, len(alphabet), np.int8)
        mapping = {s: [(i, 1) for i in ids_] for s, ids_ in zip(p.sequences, p.read_ids)}
    else:
        reads = tensors.onehot(p.codes, len(alphabet), np.float64)
        mapping = [(ids_[0], 1, []) for ids_ in p.read_ids]
    return reads, p.weights, p.qualities, mapping


def compute_reads_log_error_proba(qualities, reads_seq_binary, size_alphabet):
    """preparation.py:123-157: dense (N, L, B) FP64 tensor.  Provided for API parity; the
    engine keeps only the two distinct values per (read, position)."""
    log_hit, log_off = tensors.log_error_terms(qualities, size_alphabet)
    reads = np.asarray(reads_seq_binary)
    out = np.where(reads > 0, log_hit[:, :, None], log_off[:, :, None])
    out[reads.sum(axis=2) == 0] = 0
    return out
