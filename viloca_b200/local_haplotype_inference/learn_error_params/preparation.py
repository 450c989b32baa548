"""Window files -> tensors with the interface of learn_error_params/preparation.py."""
from __future__ import annotations

import numpy as np

from ... import tensors
from .. import _common


class Read:
    """learn_error_params/preparation.py:6-22."""

    def __init__(self, seq_string, seq_id):
        self.seq_string = seq_string
        self.weight = 1
        self.id = seq_id
        self.seq_binary = []
        self.identical_reads = []
        self.n_non_N = len(seq_string) - str(seq_string).count("N")

    def seq2binary(self, alphabet):
        self.seq_binary = tensors.onehot(tensors.encode(str(self.seq_string), alphabet), len(alphabet), np.float64)


def load_fasta2reads_list(reads_fasta_file, alphabet, unique_modus):
    """preparation.py:51-60.  The reference's O(N^2) Hamming scan (:79-94) keeps the first
    occurrence of each distinct string, adds up multiplicities and records the ids of the
    later copies; hashing gives the same list in the same order."""
    ids, seqs = _common.read_window_fasta(reads_fasta_file)
    p = _common.prepare_reads(ids, seqs, None, alphabet, unique_modus)
    out = []
    for r, s in enumerate(p.sequences):
        rd = Read(s, p.read_ids[r][0])
        rd.seq2binary(alphabet)
        rd.weight = int(p.weights[r])
        rd.identical_reads = list(p.read_ids[r][1:])
        out.append(rd)
    return out


def reads_list_to_array(reads_list):
    """preparation.py:25-48."""
    return (np.asarray([r.seq_binary for r in reads_list]), np.asarray([r.weight for r in reads_list]))


def load_reference_seq(reference_file):
    """preparation.py:97-99 -> (sequence, id); not upper-cased."""
    _, ref_id, seq = _common.load_reference(reference_file, "ACGT-", upper=False)
    return seq, ref_id


def reference2binary(reference_seq, alphabet):
    """preparation.py:102-108: float one-hot; lower-case bases fall outside the alphabet."""
    return tensors.onehot(tensors.encode(str(reference_seq), alphabet), len(alphabet), np.float64)
