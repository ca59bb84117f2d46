"""ctypes binding of the Smith-Waterman aligner (include/smithwaterman_cuda.h) and a seeded generator of
haplotype-to-reference pairs.  No fallback: a missing library or GPU raises."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from .engine import load_library

SW_OK = 0
SOFTCLIP, INDEL, LEADING_INDEL, IGNORE = 0, 1, 2, 3
DEFAULT_WEIGHTS = (200, -150, -260, -11)      # W_MATCH, W_MISMATCH, W_OPEN, W_EXTEND (htc-sw/host/common.h:15-18)
SW_EXPORTS = ["sw_create", "sw_destroy", "sw_last_error", "sw_align_batch", "sw_get_stats"]
_OPS = {0: "M", 1: "I", 2: "D", 4: "S"}


class SwStats(C.Structure):
    _fields_ = [("pairs", C.c_uint64), ("cells", C.c_uint64), ("bytes_backtrack", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("chunks", C.c_uint32), ("ms_kernel", C.c_float), ("ms_total", C.c_float)]


class SwError(RuntimeError):
    pass


def _lib():
    L = load_library()
    if not getattr(L, "_sw_ready", False):
        vp, u32 = C.c_void_p, C.c_uint32
        L.sw_create.argtypes = [C.c_int, C.POINTER(vp)]; L.sw_create.restype = C.c_int
        L.sw_destroy.argtypes = [vp]; L.sw_destroy.restype = None
        L.sw_last_error.argtypes = [vp]; L.sw_last_error.restype = C.c_char_p
        L.sw_align_batch.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32, vp, vp, vp, vp]
        L.sw_align_batch.restype = C.c_int
        L.sw_get_stats.argtypes = [vp, C.POINTER(SwStats)]; L.sw_get_stats.restype = C.c_int
        L._sw_ready = True
    return L


def cigar_string(elems) -> str:
    return "".join(f"{n}{_OPS.get(s, '?')}" for n, s in elems)


class SmithWaterman:
    """One GPU context (sw_ctx).  align(pairs) -> list of (alignment_offset, [(length, state), ...], score)."""

    def __init__(self, device: int = -1):
        self.lib = _lib()
        h = C.c_void_p()
        rc = self.lib.sw_create(device, C.byref(h))
        if rc != SW_OK:
            raise SwError(self.lib.sw_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.sw_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align(self, pairs: Sequence[tuple[bytes, bytes]], strategy: int = SOFTCLIP, weights=DEFAULT_WEIGHTS, cigar_cap: int = 64):
        n = len(pairs)
        # references are stored once each (the common case is one reference against many alternates)
        ref_pos, blob1, s1, l1 = {}, bytearray(), np.empty(n, np.uint32), np.empty(n, np.uint32)
        blob2, s2, l2 = bytearray(), np.empty(n, np.uint32), np.empty(n, np.uint32)
        for k, (r, a) in enumerate(pairs):
            if r not in ref_pos:
                ref_pos[r] = len(blob1); blob1 += r
            s1[k], l1[k] = ref_pos[r], len(r)
            s2[k], l2[k] = len(blob2), len(a); blob2 += a
        b1 = np.frombuffer(bytes(blob1) or b"\0", dtype=np.uint8); b2 = np.frombuffer(bytes(blob2) or b"\0", dtype=np.uint8)
        while True:
            cig = np.empty((max(n, 1), cigar_cap, 2), dtype=np.int32)
            ne = np.empty(max(n, 1), np.int32); off = np.empty(max(n, 1), np.int32); sc = np.empty(max(n, 1), np.int32)
            rc = self.lib.sw_align_batch(self.h, n, b1.ctypes.data, s1.ctypes.data, l1.ctypes.data, b2.ctypes.data, s2.ctypes.data,
                                         l2.ctypes.data, *weights, strategy, cigar_cap, cig.ctypes.data, ne.ctypes.data,
                                         off.ctypes.data, sc.ctypes.data)
            if rc != SW_OK:
                raise SwError(self.lib.sw_last_error(self.h).decode())
            if n == 0 or ne[:n].max() <= cigar_cap:
                break
            cigar_cap = int(ne[:n].max())             # a CIGAR did not fit: once more with room for the longest
        return [(int(off[k]), [(int(cig[k, e, 0]), int(cig[k, e, 1])) for e in range(ne[k])], int(sc[k])) for k in range(n)]

    def stats(self) -> dict:
        s = SwStats()
        self.lib.sw_get_stats(self.h, C.byref(s))
        return {f: getattr(s, f) for f, _ in s._fields_}


_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def haplotype_pairs(seed: int, n_pairs: int, ref_len: int | tuple[int, int] = (250, 500), per_ref: int = 16,
                    sub: float = 0.03, indel: float = 0.01, trim: float = 0.5):
    """(reference, alternate) pairs like GATK's haplotype-to-reference alignment: `per_ref` alternates per reference
    window, each the window with substitutions, short indels and sometimes trimmed ends (overhangs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    while len(out) < n_pairs:
        L = int(rng.integers(ref_len[0], ref_len[1] + 1)) if isinstance(ref_len, tuple) else int(ref_len)
        ref = _ACGT[rng.integers(0, 4, L)]
        for _ in range(min(per_ref, n_pairs - len(out))):
            alt, k = [], 0
            while k < L:
                u = rng.random()
                if u < sub:
                    alt.append(int(_ACGT[rng.integers(0, 4)])); k += 1
                elif u < sub + indel:
                    alt.extend(_ACGT[rng.integers(0, 4, int(rng.integers(1, 12)))].tolist())
                elif u < sub + 2 * indel:
                    k += int(rng.integers(1, 12))
                else:
                    alt.append(int(ref[k])); k += 1
            a = np.array(alt if alt else [int(ref[0])], dtype=np.uint8)
            lo = int(rng.integers(0, 9)) if rng.random() < trim else 0
            hi = len(a) - (int(rng.integers(0, 9)) if rng.random() < trim else 0)
            a = a[lo:max(hi, lo + 1)]
            if len(a) == 0:
                a = ref[:1]
            out.append((ref.tobytes(), a.tobytes()))
    return out
