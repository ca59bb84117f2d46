"""ctypes binding of the C++ host layer's C entry (pairhmm/client/pairhmm_c_api.h): one batch through
PairHMMClient + PairHMMWorker over the task plugin libPairHMMTask.so -- the reference's own client path
(/root/reference/pairhmm/client/PairHMMWorker.cpp:157-271).  Used by the tests and by bench.py (`e2e.plugin_value`).
No fallback: without the built library or without a GPU it raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import Batch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "pairhmm", "lib", "libpairhmm_host.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: build it with __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.pairhmm_worker_forward.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.POINTER(C.c_int), C.c_char_p, C.c_int]
        L.pairhmm_worker_forward.restype = C.c_int
        L.pairhmm_worker_shutdown.argtypes = []; L.pairhmm_worker_shutdown.restype = None
        _lib = L
    return _lib


class WorkerJob:
    """A batch laid out once for repeated calls (the timing loop of bench.py)."""

    def __init__(self, b: Batch):
        self.arrs = [np.ascontiguousarray(x, dtype=np.uint8) for x in (b.rs, b.q, b.i, b.d, b.c, b.hap)]
        self.ro = np.ascontiguousarray(b.read_off, dtype=np.int32)
        self.ho = np.ascontiguousarray(b.hap_off, dtype=np.int32)
        self.num_read, self.num_hap = b.num_read, b.num_hap
        self.out = np.empty(b.num_pairs, dtype=np.float64)


def worker_forward(job: WorkerJob | Batch, out: np.ndarray | None = None):
    """-> (log10 likelihoods [num_read * num_hap], pairs that took the double re-run)."""
    L = load()
    if isinstance(job, Batch):
        job = WorkerJob(job)
    if out is None:
        out = job.out
    n = C.c_int()
    err = C.create_string_buffer(512)
    a = job.arrs
    rc = L.pairhmm_worker_forward(job.num_read, job.ro.ctypes.data, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
                                  a[3].ctypes.data, a[4].ctypes.data, job.num_hap, job.ho.ctypes.data, a[5].ctypes.data,
                                  out.ctypes.data, C.byref(n), err, len(err))
    if rc != 0:
        raise RuntimeError("pairhmm_worker_forward: " + err.value.decode(errors="replace"))
    return out, int(n.value)


def shutdown() -> None:
    if _lib is not None:
        _lib.pairhmm_worker_shutdown()
