"""Compact window tensors: what the kernels read instead of the dense one-hot arrays.

The reference turns every read into a one-hot (N, L, B) array and a dense FP64
``reads_log_error_proba`` (N, L, B) (use_quality_scores/preparation.py:48-157).  Only two
distinct values exist per (read, position), so the engine keeps the base code and those two
logs; the kernels rebuild the one-hot, quality-weighted operand in registers.
"""
from __future__ import annotations

import numpy as np

ALPHABET = "ACGT-"
N_BASES = len(ALPHABET)
CODE_N = 255

_LUT = np.full(256, CODE_N, dtype=np.uint8)
for _i, _ch in enumerate(ALPHABET):
    _LUT[ord(_ch)] = _i


def encode(seq: str | bytes, alphabet: str = ALPHABET) -> np.ndarray:
    """Sequence -> base codes; anything outside the alphabet (e.g. 'N') -> 255, which is the
    all-zero one-hot row of preparation.py:58-63.  Reads are case-sensitive like the
    reference (``alphabet_dict.get``)."""
    if alphabet != ALPHABET:
        lut = np.full(256, CODE_N, dtype=np.uint8)
        for i, ch in enumerate(alphabet):
            lut[ord(ch)] = i
    else:
        lut = _LUT
    raw = seq.encode("ascii", "replace") if isinstance(seq, str) else seq
    return lut[np.frombuffer(raw, dtype=np.uint8)]


def decode(codes: np.ndarray, alphabet: str = ALPHABET) -> str:
    table = np.array(list(alphabet) + ["N"])
    idx = np.where(np.asarray(codes) < len(alphabet), codes, len(alphabet))
    return "".join(table[idx])


def log_error_terms(qualities: np.ndarray, n_bases: int = N_BASES):
    """log theta and log((1-theta)/(B-1)) with theta = 1 - 10^(-Q/10), evaluated with the same
    NumPy expression order as compute_reads_log_error_proba (preparation.py:134-138) so that
    the entries are bit-identical to the reference's E tensor."""
    q = np.asarray(qualities, dtype=np.float64)
    theta = 1 - 10 ** (-q / 10)
    return np.log(theta), np.log((1 - theta) / (n_bases - 1))


def onehot(codes: np.ndarray, n_bases: int = N_BASES, dtype=np.int8) -> np.ndarray:
    """(…,) codes -> (…, B) one-hot with zero rows for 255 (the reference's layout)."""
    codes = np.asarray(codes)
    out = np.zeros(codes.shape + (n_bases,), dtype=dtype)
    ok = codes < n_bases
    out[np.nonzero(ok) + (codes[ok].astype(np.int64),)] = 1
    return out


def codes_from_onehot(arr: np.ndarray) -> np.ndarray:
    """Inverse of :func:`onehot` for arrays coming from reference-style callers."""
    arr = np.asarray(arr)
    codes = np.argmax(arr, axis=-1).astype(np.uint8)
    codes[arr.sum(axis=-1) == 0] = CODE_N
    return codes
