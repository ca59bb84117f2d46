"""The reference test bench's on-disk batch format: a folder of ``input<i>`` / ``output<i>`` text files
(/root/reference/pairhmm/host/main.cpp:67-159 reads them).  The C++ reader is
pairhmm/host/fixture_io.h; this module writes and reads the same files so that golden folders can be minted from
synthetic batches and oracle results, and real GATK dumps can be loaded into a Batch.

input<i>:  "readListSize <R> numHaplotypes <H>", then per read its length and five captioned lines of decimal byte
           values (bases, base quals, insertion quals, deletion quals, gap-continuation quals), one blank line, then
           per haplotype its length, a caption and the bases as characters.
output<i>: per pair "<log10 likelihood> <its IEEE-754 bits as a signed 64-bit integer>"; the bits are authoritative.
"""
from __future__ import annotations

import os

import numpy as np

from .batch import Batch

_CAPTIONS = ("readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP")


def write_input(path: str, b: Batch) -> None:
    with open(path, "w") as f:
        f.write(f"readListSize {b.num_read} numHaplotypes {b.num_hap}\n")
        for k in range(b.num_read):
            tracks = b.read(k)
            f.write(f"{len(tracks[0])}\n")
            for cap, t in zip(_CAPTIONS, tracks):
                f.write(cap + "\n")
                f.write(" ".join(str(int(np.int8(x))) for x in t.astype(np.uint8).view(np.int8)) + "\n")
        f.write("\n")
        for k in range(b.num_hap):
            h = b.haplotype(k)
            f.write(f"{len(h)}\nhaplotypeBases\n{h.tobytes().decode('latin-1')}\n")


def read_input(path: str) -> Batch:
    with open(path) as f:
        lines = f.read().split("\n")
    head = lines[0].split()
    if len(head) != 4:
        raise ValueError(f"{path}: bad header")
    nr, nh = int(head[1]), int(head[3])
    pos, reads, haps = 1, [], []
    for _ in range(nr):
        ln = int(lines[pos].split()[0]); pos += 1
        tracks = []
        for _t in range(5):
            vals = np.array(lines[pos + 1].split(), dtype=np.int64)
            if len(vals) != ln:
                raise ValueError(f"{path}: track length mismatch")
            tracks.append((vals & 0xFF).astype(np.uint8)); pos += 2
        reads.append(tuple(tracks))
    pos += 1
    for _ in range(nh):
        ln = int(lines[pos].split()[0])
        bases = lines[pos + 2]
        if len(bases) != ln:
            raise ValueError(f"{path}: haplotype length mismatch")
        haps.append(np.frombuffer(bases.encode("latin-1"), dtype=np.uint8)); pos += 3
    return Batch.from_lists(reads, haps)


def write_output(path: str, log10: np.ndarray) -> None:
    v = np.ascontiguousarray(log10, dtype=np.float64).ravel()
    bits = v.view(np.int64)
    with open(path, "w") as f:
        for x, b in zip(v, bits):
            f.write(f"{x:.17g} {int(b)}\n")


def read_output(path: str, size: int | None = None) -> np.ndarray:
    bits = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 2:
                bits.append(int(t[1]))
    a = np.array(bits, dtype=np.int64).view(np.float64)
    if size is not None and len(a) < size:
        raise ValueError(f"{path}: truncated")
    return a if size is None else a[:size]


def write_folder(folder: str, batches, log10s) -> None:
    os.makedirs(folder, exist_ok=True)
    for k, (b, out) in enumerate(zip(batches, log10s)):
        write_input(os.path.join(folder, f"input{k}"), b)
        write_output(os.path.join(folder, f"output{k}"), out)
