"""Host-side batch container and the reference's wire format.

A batch is the cross product the reference's client hands to the accelerator: ``num_read`` reads (bases and the
four per-base quality tracks) against ``num_hap`` haplotypes (``read_t`` / ``hap_t``,
/root/reference/pairhmm/interface/PairHMMHostInterface.h:27-39).  Here it is stored flat: one byte array per
track plus an offsets array, which is also what the C ABI's ``pmm_*_flat`` entry points take.

``serialize_reads`` / ``serialize_haps`` produce exactly the bytes of the reference's ``serialize()``
(/root/reference/pairhmm/interface/PairHMMHostInterface.cpp:175-207): native-endian ``int32 num`` then, per
read, ``int32 len`` followed by ``len`` bytes of each of ``_b, _q, _i, _d, _c``; per haplotype ``int32 len`` and
``len`` bytes.  The C++ mirror in ``pairhmm/interface`` is checked against these in tests/test_wire_format.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Batch:
    read_off: np.ndarray  # int32 [num_read + 1]
    rs: np.ndarray        # uint8 bases (ASCII)
    q: np.ndarray         # uint8 base qualities
    i: np.ndarray         # uint8 insertion gap-open qualities
    d: np.ndarray         # uint8 deletion gap-open qualities
    c: np.ndarray         # uint8 gap-continuation qualities
    hap_off: np.ndarray   # int32 [num_hap + 1]
    hap: np.ndarray       # uint8 bases (ASCII)

    @property
    def num_read(self) -> int:
        return len(self.read_off) - 1

    @property
    def num_hap(self) -> int:
        return len(self.hap_off) - 1

    @property
    def read_lens(self) -> np.ndarray:
        return np.diff(self.read_off)

    @property
    def hap_lens(self) -> np.ndarray:
        return np.diff(self.hap_off)

    @property
    def num_pairs(self) -> int:
        return self.num_read * self.num_hap

    @property
    def num_cells(self) -> int:
        """The reference's cell count: sum(read_len) * sum(hap_len) (pairhmm/host/main.cpp:305-313)."""
        return int(self.read_lens.sum(dtype=np.int64)) * int(self.hap_lens.sum(dtype=np.int64))

    def read(self, k: int):
        a, b = int(self.read_off[k]), int(self.read_off[k + 1])
        return self.rs[a:b], self.q[a:b], self.i[a:b], self.d[a:b], self.c[a:b]

    def haplotype(self, k: int):
        return self.hap[int(self.hap_off[k]):int(self.hap_off[k + 1])]

    def slice_reads(self, lo: int, hi: int) -> "Batch":
        a, b = int(self.read_off[lo]), int(self.read_off[hi])
        return Batch((self.read_off[lo:hi + 1] - a).astype(np.int32), self.rs[a:b], self.q[a:b], self.i[a:b],
                     self.d[a:b], self.c[a:b], self.hap_off, self.hap)

    @staticmethod
    def from_lists(reads, haps) -> "Batch":
        """reads: iterable of (bases, q, i, d, c) byte-likes; haps: iterable of byte-likes."""
        reads = [tuple(np.frombuffer(bytes(t), dtype=np.uint8) if not isinstance(t, np.ndarray) else t.astype(np.uint8)
                       for t in r) for r in reads]
        haps = [np.frombuffer(bytes(h), dtype=np.uint8) if not isinstance(h, np.ndarray) else h.astype(np.uint8)
                for h in haps]
        ro = np.zeros(len(reads) + 1, dtype=np.int32)
        ro[1:] = np.cumsum([len(r[0]) for r in reads])
        ho = np.zeros(len(haps) + 1, dtype=np.int32)
        ho[1:] = np.cumsum([len(h) for h in haps])

        def cat(xs):
            return np.concatenate(xs).astype(np.uint8) if xs else np.zeros(0, dtype=np.uint8)
        return Batch(ro, cat([r[0] for r in reads]), cat([r[1] for r in reads]), cat([r[2] for r in reads]),
                     cat([r[3] for r in reads]), cat([r[4] for r in reads]), ho, cat(haps))


def serialize_reads(b: Batch) -> bytes:
    out = bytearray(np.int32(b.num_read).tobytes())
    for k in range(b.num_read):
        rs, q, i, d, c = b.read(k)
        out += np.int32(len(rs)).tobytes()
        out += rs.tobytes() + q.tobytes() + i.tobytes() + d.tobytes() + c.tobytes()
    return bytes(out)


def serialize_haps(b: Batch) -> bytes:
    out = bytearray(np.int32(b.num_hap).tobytes())
    for k in range(b.num_hap):
        h = b.haplotype(k)
        out += np.int32(len(h)).tobytes() + h.tobytes()
    return bytes(out)


def deserialize(read_blob: bytes, hap_blob: bytes) -> Batch:
    rb = np.frombuffer(read_blob, dtype=np.uint8)
    n = int(np.frombuffer(read_blob, dtype=np.int32, count=1)[0])
    pos, reads = 4, []
    for _ in range(n):
        ln = int(np.frombuffer(read_blob, dtype=np.int32, count=1, offset=pos)[0]); pos += 4
        reads.append(tuple(rb[pos + t * ln: pos + (t + 1) * ln] for t in range(5))); pos += 5 * ln
    hb = np.frombuffer(hap_blob, dtype=np.uint8)
    m = int(np.frombuffer(hap_blob, dtype=np.int32, count=1)[0])
    pos, haps = 4, []
    for _ in range(m):
        ln = int(np.frombuffer(hap_blob, dtype=np.int32, count=1, offset=pos)[0]); pos += 4
        haps.append(hb[pos:pos + ln]); pos += ln
    return Batch.from_lists(reads, haps)
