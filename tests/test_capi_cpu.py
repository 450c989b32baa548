"""CPU-side checks of the C ABI: the library loads, exports every symbol include/*.h declares,
and validates its arguments without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from viloca_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "viloca_b200.h")).read()
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
    text = open(os.path.join(ROOT, "include", "viloca_b200.h")).read()
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
    # the minority lists are sized from the read codes: entries off the per-position majority
    rng = np.random.default_rng(0)
    noisy = np.zeros((1000, 201), dtype=np.uint8)
    flip = rng.random(noisy.shape) < 0.2
    noisy[flip] = rng.integers(1, 5, size=int(flip.sum()))
    descs = (capi.WindowDesc * 1)(_desc(n=1000, l=201, k=100, r=3, codes=noisy))
    assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) == 0
    extra = nbytes.value - big
    nnz = int(flip.sum())
    assert abs(extra - 24 * nnz) <= 2048      # 4 + 8 + 4 + 8 bytes per entry, 256-byte aligned arrays
    for bad in (_desc(k=capi.MAX_CLUSTERS + 1), _desc(n=0), _desc(mode=7), _desc(max_iter=0), _desc(codes=None)):
        descs = (capi.WindowDesc * 1)(bad)
        assert capi.lib.vl_batch_workspace_bytes(descs, 1, C.byref(nbytes)) != 0
        assert len(capi.lib.vl_last_error()) > 0
    assert capi.lib.vl_batch_workspace_bytes(None, 0, C.byref(nbytes)) != 0


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
