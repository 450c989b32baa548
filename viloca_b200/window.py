"""One inference window as data (no CUDA binding is imported here, so host-only code -- the
synthetic generators, the reference arm of bench.py -- can build windows without loading the
engine library).

A :class:`Window` carries the tensors the reference's ``run_cavi`` receives
(use_quality_scores/cavi.py:70-84) in compact form plus the seeded initial states of its
restarts.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _consts

DEFAULT_MAX_ITER = 50000   # safety cap; the reference's loop has none (cavi.py:120)


@dataclass
class Window:
    """One inference window (one call of ``run_dpm_mfa.main`` in the reference)."""
    mode: str                      # "uqs" = use_quality_scores, "lep" = learn_error_params
    codes: np.ndarray              # (N, L) uint8, 0..4 = "ACGT-", 255 otherwise
    weights: np.ndarray            # (N,) read multiplicities
    ref_codes: np.ndarray          # (L,) uint8
    init_cluster: np.ndarray       # (R, N, K) initial mean_cluster per restart
    init_scalars: np.ndarray       # (R, 8) a0, b0, log g0, log(1-g0), c0, d0, log t0, log(1-t0)
    alpha0: float
    log_hit: Optional[np.ndarray] = None   # (N, L) log theta              (uqs)
    log_off: Optional[np.ndarray] = None   # (N, L) log((1-theta)/(B-1))   (uqs)
    conv_thres: float = 1e-3
    max_iter: int = DEFAULT_MAX_ITER
    name: str = ""

    def __post_init__(self):
        self.codes = np.ascontiguousarray(self.codes, dtype=np.uint8)
        self.weights = np.ascontiguousarray(self.weights, dtype=np.float64)
        self.ref_codes = np.ascontiguousarray(self.ref_codes, dtype=np.uint8)
        self.init_cluster = np.ascontiguousarray(self.init_cluster, dtype=np.float64)
        self.init_scalars = np.ascontiguousarray(self.init_scalars, dtype=np.float64)
        if self.init_cluster.ndim == 2:
            self.init_cluster = self.init_cluster[None]
        if self.init_scalars.ndim == 1:
            self.init_scalars = self.init_scalars[None]
        if self.mode not in ("uqs", "lep"):
            raise ValueError(f"mode must be 'uqs' or 'lep', got {self.mode!r}")
        n, l = self.codes.shape
        if self.weights.shape != (n,) or self.ref_codes.shape != (l,):
            raise ValueError("weights / ref_codes do not match codes")
        r = self.init_cluster.shape[0]
        if self.init_cluster.shape[1] != n or self.init_scalars.shape != (r, _consts.INIT_STRIDE):
            raise ValueError("init_cluster must be (R, N, K) and init_scalars (R, 8)")
        if self.mode == "uqs":
            if self.log_hit is None or self.log_off is None:
                raise ValueError("uqs window needs log_hit and log_off")
            self.log_hit = np.ascontiguousarray(self.log_hit, dtype=np.float64)
            self.log_off = np.ascontiguousarray(self.log_off, dtype=np.float64)
            if self.log_hit.shape != (n, l) or self.log_off.shape != (n, l):
                raise ValueError("log_hit / log_off must be (N, L)")

    @property
    def n_reads(self) -> int:
        return self.codes.shape[0]

    @property
    def length(self) -> int:
        return self.codes.shape[1]

    @property
    def n_clusters(self) -> int:
        return self.init_cluster.shape[2]

    @property
    def n_starts(self) -> int:
        return self.init_cluster.shape[0]

    def pin(self) -> "Window":
        """Re-home the host arrays in pinned memory (PyTorch owns it) so that the H2D copies
        of vl_batch_upload are true asynchronous DMA transfers."""
        import torch
        self._pinned = []
        for name in ("codes", "weights", "ref_codes", "init_cluster", "init_scalars", "log_hit", "log_off"):
            a = getattr(self, name)
            if a is None:
                continue
            t = torch.from_numpy(a).pin_memory()
            self._pinned.append(t)
            setattr(self, name, t.numpy())
        return self

    def cost(self) -> float:
        """Relative cost estimate used by the longest-first scheduler."""
        return float(self.n_reads) * self.length * self.n_clusters * self.n_starts

    def input_bytes(self) -> int:
        b = self.codes.nbytes + self.weights.nbytes + self.ref_codes.nbytes
        b += self.init_cluster.nbytes + self.init_scalars.nbytes
        if self.mode == "uqs":
            b += self.log_hit.nbytes + self.log_off.nbytes
        return b

    def desc(self):
        """The window as the C ABI's ``vl_window_desc`` (loads the engine library)."""
        from . import _capi as capi
        d = capi.WindowDesc()
        d.mode = capi.MODE_UQS if self.mode == "uqs" else capi.MODE_LEP
        d.n_reads, d.length = self.n_reads, self.length
        d.n_clusters, d.n_starts = self.n_clusters, self.n_starts
        d.max_iter = int(self.max_iter)
        d.alpha0, d.conv_thres = float(self.alpha0), float(self.conv_thres)
        d.codes = capi.u8_ptr(self.codes)
        d.log_hit = capi.f64_ptr(self.log_hit if self.mode == "uqs" else None)
        d.log_off = capi.f64_ptr(self.log_off if self.mode == "uqs" else None)
        d.weights = capi.f64_ptr(self.weights)
        d.ref_codes = capi.u8_ptr(self.ref_codes)
        d.init_cluster = capi.f64_ptr(self.init_cluster)
        d.init_scalars = capi.f64_ptr(self.init_scalars)
        return d
