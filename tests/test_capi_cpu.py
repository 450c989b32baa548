"""CPU-side checks of the C ABI: the library loads, exports every symbol include/*.h declares,
and validates its arguments without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from viloca_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_text():
    with open(os.path.join(ROOT, "include", "viloca_b200.h")) as fh:
        return fh.read()


def _declared_symbols():
    text = _header_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 13
    for name in names:
        assert hasattr(capi.lib, name), name
    assert sorted(capi.EXPORTS) == names
    assert capi.lib.vl_abi_version() == capi.ABI_VERSION


def test_struct_layout_matches_header_constants():
    text = _header_text()
    for macro, value in (("VL_N_BASES", capi.N_BASES), ("VL_CODE_N", capi.CODE_N), ("VL_MAX_CLUSTERS", capi.MAX_CLUSTERS),
                         ("VL_INIT_STRIDE", capi.INIT_STRIDE), ("VL_SCALAR_STRIDE", capi.SCALAR_STRIDE),
                         ("VL_STATUS_MAXITER", capi.STATUS_MAXITER), ("VL_MODE_LEP", capi.MODE_LEP)):
        m = re.search(rf"#define {macro}\s+(\d+)", text)
        assert m and int(m.group(1)) == value, macro
    assert C.sizeof(capi.WindowDesc) == 6 * 4 + 2 * 8 + 7 * 8
    assert C.sizeof(capi.WindowResult) == 7 * 8 + 3 * 4 + 4 + 5 * 8


_KEEP = []


def _desc(n=10, l=20, k=8, r=1, mode=0, max_iter=100, codes="zeros"):
    d = capi.WindowDesc()
    d.mode, d.n_reads, d.length, d.n_clusters, d.n_starts, d.max_iter = mode, n, l, k, r, max_iter
    d.alpha0, d.conv_thres = 1e-4, 1e-3
    if codes is not None:
        arr = np.zeros((max(n, 1), max(l, 1)), dtype=np.uint8) if isinstance(codes, str) else codes
        _KEEP.append(arr)
        d.codes = capi.u8_ptr(arr)
    return d


def test_workspace_bytes_scales_and_validates():
    nbytes = C.c_size_t(0)
    descs = (capi.WindowDesc * 1)(_desc())
    assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) == 0
    small = nbytes.value
    descs = (capi.WindowDesc * 1)(_desc(n=1000, l=201, k=100, r=3))
    assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) == 0
    big = nbytes.value
    assert big > small > 0
    # dominant terms: one FP64 (N, L) operand + R responsibility matrices
    assert big > 8 * 1000 * 201 + 3 * 8 * 1000 * 100
    # the operand is sized from the read codes.  Rare off-majority entries go to the sparse
    # lists (4 + 8 + 4 + 8 bytes each) ...
    rng = np.random.default_rng(0)
    rare = np.zeros((1000, 201), dtype=np.uint8)
    flip = rng.random(rare.shape) < 0.002
    rare[flip] = rng.integers(1, 5, size=int(flip.sum()))
    descs = (capi.WindowDesc * 1)(_desc(n=1000, l=201, k=100, r=3, codes=rare))
    assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) == 0
    assert abs((nbytes.value - big) - 24 * int(flip.sum())) <= 2048
    # ... while (position, base) pairs that many reads carry become extra dense columns
    noisy = np.zeros((1000, 201), dtype=np.uint8)
    flip = rng.random(noisy.shape) < 0.2
    noisy[flip] = rng.integers(1, 5, size=int(flip.sum()))
    descs = (capi.WindowDesc * 1)(_desc(n=1000, l=201, k=100, r=3, codes=noisy))
    assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) == 0
    assert nbytes.value - big > 4 * 201 * 1000 * 8
    for bad in (_desc(k=capi.MAX_CLUSTERS + 1), _desc(n=0), _desc(mode=7), _desc(max_iter=0), _desc(codes=None)):
        descs = (capi.WindowDesc * 1)(bad)
        assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) != 0
        assert len(capi.lib.vl_last_error()) > 0
    assert capi.lib.vl_batch_workspace_bytes(None, 0, C.byref(nbytes)) != 0


def _plan_numpy(codes):
    """NumPy restatement of the dense-column rule documented at vl_window_plan."""
    n, length = codes.shape
    cnt = np.stack([(codes == b).sum(axis=0) for b in range(5)], axis=1)
    thr = min(max(n // 128, 2), 64)
    tiles = []
    for l in range(length):
        best = int(np.argmax(cnt[l]))
        bases = [b for b in range(5) if b == best or cnt[l, b] >= thr]
        for t in tiles:                           # first fit in position order
            if len(t) + len(bases) <= 16:
                t.extend(l * 5 + b for b in bases)
                break
        else:
            tiles.append([l * 5 + b for b in bases])
    cols = []
    for t in tiles:
        cols += t + [-1] * (16 - len(t))
    dense = set(c for c in cols if c >= 0)
    sparse = sum(int(cnt[l, b]) for l in range(length) for b in range(5) if l * 5 + b not in dense)
    return cols, sparse


@pytest.mark.parametrize("n,length,p_err", [(300, 201, 0.01), (40, 16, 0.2), (1000, 57, 0.05), (5, 1, 0.5),
                                            (700, 33, 0.6), (600, 100, 0.8)])
def test_window_plan_matches_documented_rule(n, length, p_err):
    rng = np.random.default_rng(n + length)
    codes = np.tile(rng.integers(0, 4, size=length).astype(np.uint8), (n, 1))
    flip = rng.random(codes.shape) < p_err
    codes[flip] = rng.integers(0, 5, size=int(flip.sum()))
    codes[rng.random(codes.shape) < 0.1] = 255
    if length > 3:
        codes[:, 2] = 255                         # a position no read covers still gets its column
    d = _desc(n=n, l=length, codes=np.ascontiguousarray(codes))
    ncols, nsp = C.c_int32(0), C.c_int64(0)
    cap = capi.lib.vl_window_plan_capacity(length)
    assert cap == 16 * (-(-5 * length // 12) + 1)
    col_lb = np.full(cap, -7, dtype=np.int32)
    assert capi.lib.vl_window_plan(C.byref(d), C.byref(ncols), C.byref(nsp), capi.i32_ptr(col_lb)) == 0
    exp_cols, exp_sparse = _plan_numpy(codes)
    assert ncols.value == len(exp_cols) and ncols.value % 16 == 0
    assert col_lb[:ncols.value].tolist() == exp_cols
    assert nsp.value == exp_sparse
    # every position owns at least one column and its columns share a tile
    tiles = {}
    for j, lb in enumerate(exp_cols):
        if lb >= 0:
            tiles.setdefault(lb // 5, set()).add(j // 16)
    assert sorted(tiles) == list(range(length)) and all(len(t) == 1 for t in tiles.values())


def test_create_rejects_bad_workspace():
    handle = C.c_void_p()
    descs = (capi.WindowDesc * 1)(_desc())
    assert capi.lib.vl_batch_create(C.byref(handle), 0, descs, 1, None, 0, None) != 0
    assert b"workspace" in capi.lib.vl_last_error()
    with pytest.raises(capi.EngineError):
        capi.check(capi.lib.vl_batch_upload(None, descs, 1))


def test_engine_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from viloca_b200 import engine, synthetic
    import dataclasses
    spec = dataclasses.replace(synthetic.WORKLOADS["hiv5k_uqs"], n_clusters=8, called_length=20)
    win = synthetic.make_window(spec, 0, n_reads=30, length=24)
    with pytest.raises(capi.EngineError):
        engine.CaviBatch([win])


def test_reads_reader_error_behaviour():
    """vl_reads_parse / vl_reads_fill report bad input through the int return + vl_last_error, like
    the rest of the ABI; a failed parse leaves no handle behind."""
    h, a, b, c = C.c_void_p(), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    assert capi.lib.vl_reads_parse(None, 0, 1, C.byref(h), C.byref(a), C.byref(b), C.byref(c)) == 1
    assert "NULL" in capi.last_error()
    assert capi.lib.vl_reads_parse(b"", 0, 1, C.byref(h), C.byref(a), C.byref(b), C.byref(c)) == 1
    assert "without reads" in capi.last_error()
    text = b">r1 5\nACGT\n>r2 7\nACG\n"
    assert capi.lib.vl_reads_parse(text, len(text), 1, C.byref(h), C.byref(a), C.byref(b), C.byref(c)) == 1
    assert "equal length" in capi.last_error()
    text = b">r1 5\nACGTN\n>r2 7\nACGTN\n>r3 9\nacgt-\n"
    assert capi.lib.vl_reads_parse(text, len(text), 1, C.byref(h), C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (3, 2, 5)
    codes = np.zeros((2, 5), dtype=np.uint8)
    w = np.zeros(2)
    assert capi.lib.vl_reads_fill(h, b"ACGT-", None, 1, capi.u8_ptr(codes), capi.f64_ptr(w), None, None, None, None) == 1
    assert capi.lib.vl_reads_fill(h, b"ACGT-", None, 0, capi.u8_ptr(codes), capi.f64_ptr(w), None, None, None, None) == 0
    capi.lib.vl_reads_free(h)
    assert codes.tolist() == [[0, 1, 2, 3, 255], [255, 255, 255, 255, 4]] and w.tolist() == [2.0, 1.0]
