"""Seeded synthetic PairHMM batches for the five BASELINE.json configurations.

The reference's own generator (/root/reference/pairhmm/xlnx/pairhmm_test.cpp:21-82) re-seeds a default engine on
every call and therefore emits constant data; it is not reused.  The value distributions follow its intent
(:35-51, :72): base qualities ~ clamp(N(30,5), 6, 41), insertion/deletion gap-open qualities ~ clamp(N(40,1), 1,
60) with 5 % of bases at 20..30, gap-continuation 10 with 1 % other values.  One random reference window per
region (0.1 % N); haplotypes are the window with 1-4 SNPs / short indels; reads are substrings of a random
haplotype with per-base substitution errors at rate 10^(-Q/10).

numpy's PCG64 stream is stable across platforms, so a (config, seed) pair names the same bytes here and on the GPU
box.
"""
from __future__ import annotations

import numpy as np

from .batch import Batch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_N = ord("N")


def _window(rng, n):
    w = _ACGT[rng.integers(0, 4, size=n)]
    w = w.copy()
    w[rng.random(n) < 0.001] = _N
    return w


def _mutate_hap(rng, w, n_events):
    h = w
    for _ in range(n_events):
        pos = int(rng.integers(1, len(h) - 1))
        kind = rng.integers(0, 3)
        if kind == 0:      # SNP
            h = h.copy(); h[pos] = _ACGT[(np.searchsorted(_ACGT, h[pos]) + 1 + rng.integers(0, 3)) % 4] if h[pos] != _N else _ACGT[0]
        elif kind == 1:    # short insertion
            ins = _ACGT[rng.integers(0, 4, size=int(rng.integers(1, 6)))]
            h = np.concatenate([h[:pos], ins, h[pos:]])
        else:              # short deletion
            k = int(rng.integers(1, 6))
            h = np.concatenate([h[:pos], h[pos + k:]])
    return h


def _decoy(rng, h, frac):
    h = h.copy()
    m = rng.random(len(h)) < frac
    h[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
    return h


def _fit(rng, h, length):
    """Trim or extend a haplotype to exactly `length` bases."""
    if len(h) >= length:
        return h[:length]
    return np.concatenate([h, _ACGT[rng.integers(0, 4, size=length - len(h))]])


def _quals(rng, n, low=False):
    if low:
        q = rng.integers(2, 7, size=n)
    else:
        q = np.clip(np.rint(rng.normal(30, 5, size=n)), 6, 41)
    gi = np.clip(np.rint(rng.normal(40, 1, size=n)), 1, 60)
    gd = np.clip(np.rint(rng.normal(40, 1, size=n)), 1, 60)
    m = rng.random(n) < 0.05
    gi[m] = rng.integers(20, 31, size=int(m.sum()))
    m = rng.random(n) < 0.05
    gd[m] = rng.integers(20, 31, size=int(m.sum()))
    gc = np.full(n, 10)
    m = rng.random(n) < 0.01
    gc[m] = rng.integers(5, 31, size=int(m.sum()))
    return q.astype(np.uint8), gi.astype(np.uint8), gd.astype(np.uint8), gc.astype(np.uint8)


def _read_from(rng, h, length, low=False):
    length = min(length, len(h))
    off = int(rng.integers(0, len(h) - length + 1))
    b = h[off:off + length].copy()
    q, gi, gd, gc = _quals(rng, length, low)
    err = rng.random(length) < np.power(10.0, -q.astype(np.float64) / 10.0)
    b[err] = _ACGT[rng.integers(0, 4, size=int(err.sum()))]
    return b, q, gi, gd, gc


def region(rng, read_lens, hap_lens, decoy_frac=0.0, low_read_frac=0.0):
    """One active region: len(read_lens) reads against len(hap_lens) haplotypes."""
    hmax = int(max(hap_lens))
    w = _window(rng, hmax + 16)
    n_decoy = int(round(decoy_frac * len(hap_lens)))
    haps, sources = [], []
    for k, hl in enumerate(hap_lens):
        h = _fit(rng, _mutate_hap(rng, w, int(rng.integers(1, 5))), int(hl))
        if k < n_decoy:
            haps.append(_decoy(rng, h, 0.2))
        else:
            haps.append(h); sources.append(h)
    if not sources:
        sources = [w[:hmax]]
    reads = []
    for rl in read_lens:
        src = sources[int(rng.integers(0, len(sources)))]
        reads.append(_read_from(rng, src, int(rl), low=bool(rng.random() < low_read_frac)))
    return Batch.from_lists(reads, haps)


def config(idx: int, seed: int | None = None, scale: float = 1.0):
    """BASELINE.json configs[idx-1].  Returns a list of Batch (one per active region).

    scale < 1 shrinks the number of reads (configs 1-4) or regions (config 5) for quick parity cases.
    """
    seed = idx if seed is None else seed
    rng = np.random.Generator(np.random.PCG64(seed))

    def n(x):
        return max(1, int(round(x * scale)))
    if idx == 1:     # 128 reads x 101 bp, 32 haps x 400 bp
        return [region(rng, [101] * n(128), [400] * 32)]
    if idx == 2:     # 1000 reads x 151 bp, 64 haps of 300..600 bp
        return [region(rng, [151] * n(1000), rng.integers(300, 601, size=64))]
    if idx == 3:     # shape of config 2, underflow-heavy
        return [region(rng, [151] * n(1000), rng.integers(300, 601, size=64), decoy_frac=0.5, low_read_frac=0.25)]
    if idx == 4:     # 512 reads x 250 bp, 32 haps 1..2 kb skewed
        hl = np.where(rng.random(32) < 0.8, rng.integers(1000, 1201, size=32), rng.integers(1200, 2049, size=32))
        return [region(rng, [250] * n(512), hl)]
    if idx == 5:     # 2500 regions x (100 reads x 40 haps); 10 % of reads soft-clipped to 60..150
        out = []
        for _ in range(n(2500)):
            rl = np.where(rng.random(100) < 0.1, rng.integers(60, 151, size=100), 151)
            out.append(region(rng, rl, rng.integers(300, 601, size=40)))
        return out
    raise ValueError(f"unknown config {idx}")


CONFIG_NAMES = {
    1: "cfg1: 128 reads x 101 bp vs 32 haps x 400 bp",
    2: "cfg2: 1000 reads x 151 bp vs 64 haps x 300-600 bp (GATK-HaplotypeCaller-like active region)",
    3: "cfg3: cfg2 shape, underflow-heavy (decoy haplotypes, low-quality reads)",
    4: "cfg4: 512 reads x 250 bp vs 32 haps x 1-2 kb (skewed)",
    5: "cfg5: 2500 regions x (100 reads x 40 haps), 151 bp reads (10 % clipped) vs 300-600 bp haps",
}
