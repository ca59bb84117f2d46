"""The oracle is pinned before it is trusted: against the committed golden vectors (minted from the reference's own
AVX code, tests/golden/make_golden.py), and -- where oracle/_ref exists -- against that code directly."""
import numpy as np
import pytest

from conftest import assert_bits_equal


def test_port_tables_match_golden(port, golden):
    _, tables = golden
    for name, ref in tables.items():
        assert_bits_equal(port.table(name), ref, name)


def test_port_matches_golden_vectors(port, golden):
    cases, _ = golden
    for name, (b, raw_bits, log10_bits, mask) in cases.items():
        raw, out, fb = port.batch(b, threads=4)
        assert_bits_equal(raw, raw_bits.view(np.float32), f"{name} raw")
        assert np.array_equal(fb, mask), f"{name} fallback mask"
        assert_bits_equal(out, log10_bits.view(np.float64), f"{name} log10")


def test_golden_covers_the_edge_cases(golden):
    cases, _ = golden
    assert cases["cfg3"][3].mean() > 0.3                      # underflow-heavy: fallback path exercised
    assert np.isinf(cases["deep_underflow"][2].view(np.float64)).any()     # results flushed to exactly zero
    lg = cases["deep_underflow"][2].view(np.float64)
    assert ((lg < -548) & np.isfinite(lg)).any()              # between 2^-800 and the flush limit
    assert cases["long_reads"][0].read_lens.max() > 511      # beyond one 32 x 16 row block
    assert cases["ragged"][0].read_lens.min() == 1 and cases["ragged"][0].hap_lens.min() == 1


def test_reference_matches_golden_vectors(reference, golden):
    if reference is None:
        pytest.skip("oracle/_ref not built on this machine")
    cases, tables = golden
    for name, ref in tables.items():
        assert_bits_equal(reference.table(name), ref, name)
    for name, (b, raw_bits, log10_bits, mask) in cases.items():
        raw, out, fb = reference.batch(b, threads=4)
        assert_bits_equal(raw, raw_bits.view(np.float32), f"{name} raw")
        assert np.array_equal(fb, mask)
        assert_bits_equal(out, log10_bits.view(np.float64), f"{name} log10")


def test_port_matches_reference_on_fresh_inputs(port, reference):
    if reference is None:
        pytest.skip("oracle/_ref not built on this machine")
    from acc_genomics_b200 import synth
    for cfg, scale, seed in ((1, 0.2, 11), (2, 0.012, 12), (3, 0.012, 13), (4, 0.012, 14)):
        b = synth.config(cfg, seed=seed, scale=scale)[0]
        r = reference.batch(b, threads=8)
        p = port.batch(b, threads=8)
        assert_bits_equal(p[0], r[0], f"cfg{cfg} raw")
        assert np.array_equal(p[2], r[2])
        assert_bits_equal(p[1], r[1], f"cfg{cfg} log10")


def test_single_pair_entry_points(port, golden):
    cases, _ = golden
    b, raw_bits, _, _ = cases["cfg2"]
    raw = raw_bits.view(np.float32)
    for i, j in ((0, 0), (3, 17), (7, 63)):
        v = port.f32(*b.read(i), b.haplotype(j))
        assert np.float32(v).view(np.uint32) == raw[i, j].view(np.uint32)
        d = port.f64(*b.read(i), b.haplotype(j))
        assert abs(np.log10(d) - 1020 * np.log10(2) - (np.log10(np.float64(v)) - 120 * np.log10(2))) < 1e-4 or v < 1e-28


def test_avx_and_scalar_baseline_agree_within_tolerance(reference, golden):
    """The reference's scalar spec (baseline_impl.cpp) differs from its AVX code only in low bits (reduction order,
    the stripe artefact): log10 within 1e-5 relative, as SURVEY.md section 2 row 2 reports."""
    if reference is None:
        pytest.skip("oracle/_ref not built on this machine")
    cases, _ = golden
    b, raw_bits, _, _ = cases["cfg1"]
    raw = raw_bits.view(np.float32)
    for i in range(0, b.num_read, 3):
        for j in range(0, b.num_hap, 5):
            s = reference.baseline_f32(*b.read(i), b.haplotype(j))
            a, c = np.log10(np.float64(raw[i, j])), np.log10(np.float64(s))
            assert abs(a - c) <= 1e-5 * abs(a)


def test_stripe_artefact_is_void(port, golden):
    """The AVX stripe initialisation hands M[r-1][1] to the first row of every later stripe as its "left M" in
    column 1 (avx-pairhmm-template.h:171-176).  M[r-1][1] is exactly 0 for r-1 >= 2, so a plain row-major recurrence
    with or without that term gives the bits of the oracle (which gives the bits of the reference, see above).
    Short haplotypes and cheap gaps are where column 1 would matter if it did."""
    from acc_genomics_b200.batch import Batch
    _, tables = golden
    rng = np.random.Generator(np.random.PCG64(5))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = [(acgt[rng.integers(0, 4, size=n)], np.full(n, 20, np.uint8), np.full(n, 3, np.uint8), np.full(n, 3, np.uint8),
              np.full(n, 2, np.uint8)) for n in (9, 13, 21)]
    b = Batch.from_lists(reads, [acgt[rng.integers(0, 4, size=n)] for n in (2, 3, 5, 9, 11, 17)])
    ph, m2m = tables["ph2pr_f64"], tables["m2m_f64"]
    cls = np.zeros(256, dtype=np.int64); cls[ord("C")] = 1; cls[ord("T")] = 2; cls[ord("G")] = 3; cls[ord("N")] = 4

    def plain(i, j, artefact):
        rs, q, gi, gd, gc = [x.astype(np.int64) for x in b.read(i)]
        hap = cls[b.haplotype(j)]
        R, C = len(rs), len(hap)
        pM = np.zeros(C + 1); pX = np.zeros(C + 1); pY = np.full(C + 1, 2.0 ** 1020 / C)
        for r in range(1, R + 1):
            a, d_, c_, q_ = gi[r - 1] & 127, gd[r - 1] & 127, gc[r - 1] & 127, q[r - 1] & 127
            mx, mn = max(a, d_), min(a, d_)
            pMM, pG, pMX, pXX, pMY = m2m[mx * (mx + 1) // 2 + mn], 1.0 - ph[c_], ph[a], ph[c_], ph[d_]
            wm, wx = 1.0 - ph[q_], ph[q_] / 3.0
            rc = cls[rs[r - 1]]
            cM = np.zeros(C + 1); cX = np.zeros(C + 1); cY = np.zeros(C + 1)
            for c in range(1, C + 1):
                w = wm if (rc == hap[c - 1] or rc == 4 or hap[c - 1] == 4) else wx
                cM[c] = ((pM[c - 1] * pMM + pX[c - 1] * pG) + pY[c - 1] * pG) * w
                cX[c] = pM[c] * pMX + pX[c] * pXX
                ml = pM[1] if (artefact and c == 1 and r > 1 and (r - 1) % 4 == 0) else cM[c - 1]
                cY[c] = ml * pMY + cY[c - 1] * pXX
            pM, pX, pY = cM, cX, cY
        sm = sx = 0.0
        for c in range(1, C + 1):
            sm += pM[c]; sx += pX[c]
        return sm + sx

    for i, j in ((0, 0), (1, 1), (2, 2), (0, 3), (1, 4), (2, 5)):
        want = port.f64(*b.read(i), b.haplotype(j))
        assert plain(i, j, True) == want
        assert plain(i, j, False) == want
