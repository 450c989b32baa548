"""Batched CAVI on one GPU: the Python face of the C ABI.

A :class:`Window` carries the tensors the reference's ``run_cavi`` receives
(use_quality_scores/cavi.py:70-84) in compact form plus the seeded initial states of its
restarts; :class:`CaviBatch` runs all (window, restart) problems of a list of windows on one
device.  PyTorch is used only to own the device workspace and pinned host buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _capi as capi

from .window import DEFAULT_MAX_ITER, Window  # noqa: F401  (re-exported)


@dataclass
class RestartOutcome:
    """``dict_result`` of one ``run_cavi`` call (use_quality_scores/cavi.py:190-206)."""
    elbo: float
    n_iterations: int
    n_updates: int
    status: int
    converged: bool
    history_elbo: np.ndarray

    @property
    def exit_message(self) -> str:
        return capi.STATUS_MESSAGE[self.status]


@dataclass
class WindowOutcome:
    restarts: list
    best_restart: int
    state_restart: int
    state: dict = field(default_factory=dict)   # keys of state_curr_dict for state_restart


class CaviBatch:
    """All (window, restart) problems of ``windows`` on one device."""

    def __init__(self, windows: Sequence[Window], device: int = 0, stream: Optional[int] = None,
                 skip_dead: bool = True, fused_chunks: bool = True):
        if not windows:
            raise ValueError("empty batch")
        self.windows = list(windows)
        self.device = int(device)
        self._descs = (capi.WindowDesc * len(self.windows))(*[w.desc() for w in self.windows])
        nbytes = C.c_size_t(0)
        capi.check(capi.lib.vl_batch_workspace_bytes(self._descs, len(self.windows), C.byref(nbytes)))
        self.workspace_bytes = int(nbytes.value)
        if not torch.cuda.is_available():
            raise capi.EngineError("no CUDA device: the CAVI engine has no CPU fallback")
        self._workspace = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8,
                                      device=torch.device("cuda", self.device))
        ptr = (self._workspace.data_ptr() + 255) // 256 * 256
        self._handle = C.c_void_p()
        capi.check(capi.lib.vl_batch_create(C.byref(self._handle), self.device, self._descs,
                                            len(self.windows), C.c_void_p(ptr),
                                            self.workspace_bytes, C.c_void_p(stream or 0)))
        self.n_problems = sum(w.n_starts for w in self.windows)
        self.launches = 0
        if not skip_dead:
            capi.check(capi.lib.vl_batch_set_option(self._handle, capi.OPT_SKIP_DEAD, 0))
        capi.check(capi.lib.vl_batch_set_option(self._handle, capi.OPT_FUSED_CHUNKS, 1 if fused_chunks else 0))

    # -- lifetime ------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            capi.lib.vl_batch_destroy(self._handle)
            self._handle = C.c_void_p()
        self._workspace = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data path -----------------------------------------------------------------------
    def input_bytes(self) -> int:
        return sum(w.input_bytes() for w in self.windows)

    def upload(self) -> None:
        capi.check(capi.lib.vl_batch_upload(self._handle, self._descs, len(self.windows)))

    def run(self) -> int:
        n = C.c_int64(0)
        capi.check(capi.lib.vl_batch_run(self._handle, C.byref(n)))
        self.launches += int(n.value)
        return int(n.value)

    def step(self, n_updates: int = 1) -> int:
        n = C.c_int64(0)
        capi.check(capi.lib.vl_batch_step(self._handle, int(n_updates), C.byref(n)))
        self.launches += int(n.value)
        return int(n.value)

    def profile_step(self, n_updates: int):
        """Per-kernel device time (ms) and launch counts over ``n_updates`` iterations:
        {"haplo": (ms, n), "cluster": (ms, n)}."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        capi.check(capi.lib.vl_batch_profile_step(self._handle, int(n_updates), ms, cnt))
        self.launches += int(sum(cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("haplo", "cluster"))}

    def problem_stats(self) -> np.ndarray:
        """(n_problems, 4) int32: running, 8-cluster tiles, clusters in the layout, updates."""
        out = np.zeros((self.n_problems, 4), dtype=np.int32)
        capi.check(capi.lib.vl_batch_problem_stats(self._handle, capi.i32_ptr(out), self.n_problems))
        return out

    def last_device_ms(self) -> float:
        return float(capi.lib.vl_batch_last_device_ms(self._handle))

    def set_state(self, window: int, restart: int, mean_cluster, mean_log_pi, scalars, it: int) -> None:
        mc = np.ascontiguousarray(mean_cluster, dtype=np.float64)
        lp = np.ascontiguousarray(mean_log_pi, dtype=np.float64)
        sc = np.ascontiguousarray(scalars, dtype=np.float64)
        assert sc.shape == (capi.SCALAR_STRIDE,)
        capi.check(capi.lib.vl_batch_set_state(self._handle, window, restart, capi.f64_ptr(mc),
                                               capi.f64_ptr(lp), capi.f64_ptr(sc), int(it)))

    def merge_haplo(self, window: int, restart: int, group_of: np.ndarray, n_groups: int):
        """analyze_results.py:65-80 on the device: merged responsibilities (N, G) and the
        haplotype table (G, L, B) recomputed from them."""
        w = self.windows[window]
        g = np.ascontiguousarray(group_of, dtype=np.int32)
        zc = np.empty((w.n_reads, n_groups), dtype=np.float64)
        hm = np.empty((n_groups, w.length, capi.N_BASES), dtype=np.float64)
        capi.check(capi.lib.vl_batch_merge_haplo(self._handle, window, restart, capi.i32_ptr(g),
                                                 int(n_groups), capi.f64_ptr(zc), capi.f64_ptr(hm)))
        return zc, hm

    def _buffer(self, key, shape, dtype, pinned: bool, zero: bool = False) -> np.ndarray:
        """Result buffer: a fresh pageable array, or (``pinned``) a page-locked one owned by the
        batch and REUSED by the next pinned fetch, so that the device -> host copies are plain
        DMA transfers."""
        if not pinned:
            return np.zeros(shape, dtype) if zero else np.empty(shape, dtype)
        cache = self.__dict__.setdefault("_pinned_results", {})
        arr = cache.get(key)
        if arr is None or arr.shape != tuple(np.atleast_1d(shape)) or arr.dtype != np.dtype(dtype):
            t = torch.empty(tuple(np.atleast_1d(shape)), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            arr = t.numpy()
            cache[key] = arr
            cache[("tensor",) + key] = t
            if zero:
                arr[...] = 0
        return arr

    def fetch(self, state: str | int = "best", want_cluster: bool = True, want_haplo: bool = True,
              want_history: bool = True, pinned: bool = False) -> list:
        """Copy results to the host.  ``state`` = "best" (max-ELBO restart), "none", a restart
        index, or a per-window list of those.  ``pinned=True`` returns views of page-locked
        buffers owned by the batch (valid until the next pinned fetch)."""
        nw = len(self.windows)
        states = list(state) if isinstance(state, (list, tuple)) else [state] * nw
        res = (capi.WindowResult * nw)()
        keep = []
        for i, w in enumerate(self.windows):
            r_ = w.n_starts
            state = states[i]
            bufs = {
                "elbo": np.empty(r_, np.float64), "n_iterations": np.empty(r_, np.int32),
                "n_updates": np.empty(r_, np.int32), "status": np.empty(r_, np.int32),
                "converged": np.empty(r_, np.int32), "history_len": np.empty(r_, np.int32),
            }
            res[i].elbo = capi.f64_ptr(bufs["elbo"])
            for key in ("n_iterations", "n_updates", "status", "converged", "history_len"):
                setattr(res[i], key, capi.i32_ptr(bufs[key]))
            if want_history:
                bufs["history"] = self._buffer((i, "history"), (r_, w.max_iter), np.float64, pinned, zero=True)
                res[i].history = capi.f64_ptr(bufs["history"])
                res[i].history_stride = w.max_iter
            if state == "none":
                res[i].state_restart = -1
            else:
                res[i].state_restart = -1 if state == "best" else int(state)
                k, l, n = w.n_clusters, w.length, w.n_reads
                if want_haplo:
                    bufs["mean_haplo"] = self._buffer((i, "mean_haplo"), (k, l, capi.N_BASES), np.float64, pinned)
                    res[i].mean_haplo = capi.f64_ptr(bufs["mean_haplo"])
                if want_cluster:
                    bufs["mean_cluster"] = self._buffer((i, "mean_cluster"), (n, k), np.float64, pinned)
                    res[i].mean_cluster = capi.f64_ptr(bufs["mean_cluster"])
                if pinned:      # (a pageable destination would turn these two copies into blocking ones)
                    small = self._buffer(("small",), (nw, 2, capi.MAX_CLUSTERS), np.float64, True)
                    bufs["alpha"], bufs["mean_log_pi"] = small[i, 0, :k], small[i, 1, :k]
                else:
                    bufs["alpha"] = np.empty(k, np.float64)
                    bufs["mean_log_pi"] = np.empty(k, np.float64)
                bufs["scalars"] = np.empty(capi.SCALAR_STRIDE, np.float64)
                res[i].alpha = capi.f64_ptr(bufs["alpha"])
                res[i].mean_log_pi = capi.f64_ptr(bufs["mean_log_pi"])
                res[i].scalars = capi.f64_ptr(bufs["scalars"])
            keep.append(bufs)
        capi.check(capi.lib.vl_batch_fetch(self._handle, res, nw))
        out = []
        for i, w in enumerate(self.windows):
            b = keep[i]
            state = states[i]
            restarts = []
            for r in range(w.n_starts):
                hl = int(b["history_len"][r])
                hist = b["history"][r, :hl].copy() if want_history else np.empty(0)   # (a copy: safe with pinned buffers)
                restarts.append(RestartOutcome(
                    elbo=float(b["elbo"][r]), n_iterations=int(b["n_iterations"][r]),
                    n_updates=int(b["n_updates"][r]), status=int(b["status"][r]),
                    converged=bool(b["converged"][r]), history_elbo=hist))
            best = int(res[i].best_restart)
            sr = best if state in ("best", "none") else int(state)
            st = {}
            if state != "none":
                sc = b["scalars"]
                st = {
                    "alpha": b["alpha"], "mean_log_pi": b["mean_log_pi"],
                    "gamma_a": sc[capi.S_GAMMA_A], "gamma_b": sc[capi.S_GAMMA_B],
                    "mean_log_gamma": (sc[capi.S_MLG0], sc[capi.S_MLG1]),
                    "digamma_alpha_sum": sc[capi.S_DSUM_ALPHA], "digamma_a_b_sum": sc[capi.S_DSUM_AB],
                    "elbo": sc[capi.S_ELBO],
                    "elbo_terms": (sc[capi.S_TERM_DATA], sc[capi.S_TERM_PI], sc[capi.S_TERM_CLUSTER],
                                   sc[capi.S_TERM_HAPLO]),
                }
                if w.mode == "lep":
                    st.update({"theta_c": sc[capi.S_THETA_C], "theta_d": sc[capi.S_THETA_D],
                               "mean_log_theta": (sc[capi.S_MLT0], sc[capi.S_MLT1]),
                               "digamma_c_d_sum": sc[capi.S_DSUM_CD]})
                if want_haplo:
                    st["mean_haplo"] = b["mean_haplo"]
                if want_cluster:
                    st["mean_cluster"] = b["mean_cluster"]
            out.append(WindowOutcome(restarts=restarts, best_restart=best, state_restart=sr, state=st))
        return out


def run_windows(windows: Sequence[Window], device: int = 0, **options) -> list:
    """upload + run + fetch of one batch: the end-to-end call with host buffers.  ``options``
    are the keyword options of :class:`CaviBatch` (skip_dead, fused_chunks)."""
    with CaviBatch(windows, device=device, **options) as batch:
        batch.upload()
        batch.run()
        return batch.fetch()


# ------------------------------------------------------------------------------------------
# Waves: batches bounded by the HBM that is free, the next one uploaded while this one runs
# ------------------------------------------------------------------------------------------

def window_workspace_bytes(window: Window) -> int:
    """Device workspace one window needs (all its restarts), from the library's own layout."""
    desc = (capi.WindowDesc * 1)(window.desc())
    nbytes = C.c_size_t(0)
    capi.check(capi.lib.vl_batch_workspace_bytes(desc, 1, C.byref(nbytes)))
    return int(nbytes.value)


def plan_waves(sizes: Sequence[int], budget: int) -> list:
    """Consecutive groups of windows (the caller's order is kept) whose workspaces fit ``budget``
    bytes together; a window larger than the budget is a wave of its own.  The reference never
    holds more than ``max_proc`` windows at a time (shotgun.py:527-538); here the bound is HBM."""
    waves, cur, used = [], [], 0
    for i, s in enumerate(sizes):
        if cur and used + s > budget:
            waves.append(cur)
            cur, used = [], 0
        cur.append(i)
        used += int(s)
    if cur:
        waves.append(cur)
    return waves


def run_in_waves(windows: Sequence[Window], device: int = 0, budget_bytes: Optional[int] = None, consume=None,
                 fetch_options: Optional[dict] = None, stats: Optional[dict] = None, **options) -> list:
    """upload + run + fetch of ``windows`` in HBM-bounded waves on one GPU.  While wave i runs its
    loops, wave i + 1 is created and uploaded by a helper thread on its own stream (host planning,
    H2D copies and the operand kernels overlap the run), so two waves are resident at a time and each
    may take up to ``budget_bytes`` (default: 45 % of the HBM that is free now).  ``consume(batch,
    indices, outcomes)`` is called while a wave's batch is still alive (the finishing contraction
    needs it); the outcomes are returned in the order of ``windows``."""
    from concurrent.futures import ThreadPoolExecutor
    windows = list(windows)
    if not windows:
        return []
    if not torch.cuda.is_available():
        raise capi.EngineError("no CUDA device: the CAVI engine has no CPU fallback")
    dev = torch.device("cuda", int(device))
    if budget_bytes is None:
        free, _total = torch.cuda.mem_get_info(dev)
        budget_bytes = int(0.45 * free)
    sizes = [window_workspace_bytes(w) for w in windows]
    waves = plan_waves(sizes, budget_bytes)
    if stats is not None:
        stats.update({"waves": len(waves), "budget_bytes": int(budget_bytes), "workspace_bytes": int(sum(sizes)),
                      "largest_wave_bytes": max(sum(sizes[i] for i in w) for w in waves),
                      "device_ms": 0.0, "launches": 0, "upload_s": 0.0, "run_s": 0.0, "fetch_s": 0.0})

    def prepare(idx):
        import time as _t
        t0 = _t.perf_counter()
        with torch.cuda.device(dev):
            stream = torch.cuda.Stream(device=dev)
            batch = CaviBatch([windows[i] for i in idx], device=device, stream=stream.cuda_stream, **options)
            batch._stream_keepalive = stream
            batch.upload()
            stream.synchronize()
        return batch, _t.perf_counter() - t0

    import time
    results = [None] * len(windows)
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = pool.submit(prepare, waves[0])
        for k, idx in enumerate(waves):
            batch, up_s = pending.result()
            pending = pool.submit(prepare, waves[k + 1]) if k + 1 < len(waves) else None
            try:
                t0 = time.perf_counter()
                n = batch.run()
                t1 = time.perf_counter()
                outcomes = batch.fetch(**(fetch_options or {}))
                t2 = time.perf_counter()
                if stats is not None:
                    stats["device_ms"] += batch.last_device_ms(); stats["launches"] += n
                    stats["upload_s"] += up_s; stats["run_s"] += t1 - t0; stats["fetch_s"] += t2 - t1
                if consume is not None:
                    consume(batch, idx, outcomes)
                for i, o in zip(idx, outcomes):
                    results[i] = o
            finally:
                batch.close()
    return results
