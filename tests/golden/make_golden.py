"""Regenerates tests/golden/pairhmm_golden.npz from the REFERENCE's own implementation.

Run in the build container (needs /root/reference to have been compiled into oracle/_ref by oracle/Makefile):
    python tests/golden/make_golden.py
The reference ships no golden vectors for this path (its fixtures live on S3/NFS), so these
are minted from its AVX code compiled with the pinned flags -O3 -mavx -ffp-contract=off: raw float likelihoods
(bit patterns), the float->double fallback mask, and the final log10 doubles (bit patterns), read-major, exactly
the contract of FalconPairHMM::computePairhmmAVX (/root/reference/pairhmm/xlnx/host/FalconPairHMM.cpp:69-95).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from acc_genomics_b200 import synth  # noqa: E402
from acc_genomics_b200.batch import Batch  # noqa: E402


def edge_cases():
    rng = np.random.Generator(np.random.PCG64(1234))
    cases = {}
    # every read length 1..40 against ragged haplotype lengths (stripe boundaries of 4/8 rows, lanes, 1-base inputs)
    cases["ragged"] = synth.region(rng, list(range(1, 41)), [1, 2, 3, 5, 8, 13, 31, 32, 33, 64, 65, 100])
    # odd bytes: N on both sides, lower case, gap characters, NUL, 0xFE; qualities with bit 7 set, 0 and 127
    b = synth.region(rng, [37, 64, 100, 151], [90, 200, 333])
    for arr in (b.rs, b.hap):
        m = rng.random(arr.size) < 0.08
        arr[m] = rng.choice(np.frombuffer(b"NNacgtn-*\x00\xfe", dtype=np.uint8), size=int(m.sum()))
    for arr in (b.q, b.i, b.d, b.c):
        m = rng.random(arr.size) < 0.1
        arr[m] = rng.choice(np.array([0, 1, 127, 128, 130, 200, 255], dtype=np.uint8), size=int(m.sum()))
    cases["odd_bytes"] = b
    # long reads: beyond one 32 x 16 row block (striped float kernel) and beyond 191 rows (striped double kernel)
    cases["long_reads"] = synth.region(rng, [192, 300, 511, 512, 513, 700], [600, 650, 800, 720], decoy_frac=0.5)
    # unrelated reads vs random haplotypes with expensive gaps: likelihoods around and below the double range
    # (2^-1022 / 2^1020 ~ 1e-615): results below 2^-800 take the flush-to-zero emulating kernel, some are exactly 0
    w = synth.region(rng, [120, 140, 160, 165, 170, 175, 180, 185, 190, 200, 250], [300, 500])
    w.hap[:] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=w.hap.size)]
    w.q[:] = 60
    w.i[:] = 90; w.d[:] = 90; w.c[:] = 90
    cases["deep_underflow"] = w
    return cases


def main():
    ref = oracle.reference()
    assert ref is not None, "oracle/_ref is not built (needs /root/reference)"
    cases = {
        "cfg1": synth.config(1, scale=0.1)[0],
        "cfg2": synth.config(2, scale=0.008)[0],
        "cfg3": synth.config(3, scale=0.008)[0],
        "cfg4": synth.config(4, scale=0.01)[0],
    }
    for k, b in enumerate(synth.config(5, scale=0.0008)):
        cases[f"cfg5_r{k}"] = b.slice_reads(0, 12)
    cases.update(edge_cases())
    out = {}
    for name, b in cases.items():
        raw, log10, mask = ref.batch(b, threads=8)
        for f in ("read_off", "rs", "q", "i", "d", "c", "hap_off", "hap"):
            out[f"{name}/{f}"] = getattr(b, f)
        out[f"{name}/raw_bits"] = raw.view(np.uint32)
        out[f"{name}/log10_bits"] = log10.view(np.uint64)
        out[f"{name}/mask"] = mask
        print(f"{name}: {b.num_read} x {b.num_hap}, fallback {mask.mean():.2f}, -inf {np.isinf(log10).sum()}, "
              f"min log10 {log10[np.isfinite(log10)].min():.1f}")
    for nm in ("ph2pr_f32", "ph2pr_f64", "m2m_f32", "m2m_f64"):
        out[f"tables/{nm}"] = ref.table(nm)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pairhmm_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
