"""Smith-Waterman on the GPU against the oracle: alignment offset and CIGAR identical, every overhang strategy."""
import json
import os

import numpy as np
import pytest

import oracle
from acc_genomics_b200 import sw

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def aligner(built):
    a = sw.SmithWaterman(0)
    yield a
    a.close()


@pytest.fixture(scope="module")
def sw_checker(built):
    return oracle.sw_reference() or oracle.sw_port()


def check(aligner, chk, pairs, strategy, weights=sw.DEFAULT_WEIGHTS):
    got = aligner.align(pairs, strategy, weights)
    for (r, a), (off, cig, score) in zip(pairs, got):
        want = chk.align(r, a, strategy, weights)
        assert (off, cig) == want, f"strategy {strategy} ref {len(r)} alt {len(a)}: {sw.cigar_string(cig)} @ {off} vs {sw.cigar_string(want[1])} @ {want[0]}"


def test_golden_vectors(aligner):
    with open(os.path.join(ROOT, "tests", "golden", "sw_golden.json")) as f:
        cases = json.load(f)
    by = {}
    for c in cases:
        by.setdefault((c["strategy"], tuple(c.get("weights", sw.DEFAULT_WEIGHTS))), []).append(c)
    for (st, w), cs in by.items():
        got = aligner.align([(c["ref"].encode(), c["alt"].encode()) for c in cs], st, w)
        for c, (off, cig, _) in zip(cs, got):
            assert off == c["offset"] and [list(e) for e in cig] == c["cigar"], c


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_haplotype_batches(aligner, sw_checker, strategy):
    check(aligner, sw_checker, sw.haplotype_pairs(20 + strategy, 260, ref_len=(200, 520), per_ref=26), strategy)


@pytest.mark.parametrize("strategy", [0, 1, 2, 3])
def test_all_small_lengths(aligner, sw_checker, strategy):
    rng = np.random.Generator(np.random.PCG64(5))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    pairs = [(acgt[rng.integers(0, 4, a)].tobytes(), acgt[rng.integers(0, 2, b)].tobytes()) for a in range(1, 20) for b in range(1, 20)]
    check(aligner, sw_checker, pairs, strategy)


def test_block_boundaries_and_long_sequences(aligner, sw_checker):
    """Row counts around the row blocks (32 lanes x 4, 6, .. 14 rows per lane, chosen per pair: 128 .. 448 rows, then two blocks),
    column counts around the 8-cell backtrack word and the 64-column traceback tile, and the reference's maximum (1536)."""
    pairs = []
    for n1 in (127, 128, 129, 191, 192, 193, 255, 256, 257, 319, 320, 321, 447, 448, 449, 895, 896, 897):
        for n2 in (7, 8, 9, 63, 64, 65, 130):
            pairs += sw.haplotype_pairs(n1 * 131 + n2, 1, ref_len=n1, per_ref=1, trim=0.0)
            r, a = pairs[-1]
            pairs[-1] = (r, (a * 3)[:n2] if len(a) < n2 else a[:n2])
    pairs += [(r[:1536], a[:1536]) for r, a in sw.haplotype_pairs(99, 3, ref_len=(1500, 1536), per_ref=3)]
    for st in range(4):
        check(aligner, sw_checker, pairs, st)
    # beyond the reference's fixed buffers (MAX_SEQ_LEN = 1536) only the C restatement can check
    long = sw.haplotype_pairs(98, 2, ref_len=(2500, 3000), per_ref=2)
    check(aligner, oracle.sw_port(), long, 0)
    check(aligner, oracle.sw_port(), long, 3)
    # the longest the C ABI accepts: four warps' shared memory no longer fits a CTA, the launch uses fewer warps per CTA
    longest = [(r[:4095], a[:4095]) for r, a in sw.haplotype_pairs(99, 3, ref_len=(4095, 4200), per_ref=3)]
    assert max(len(r) for r, _ in longest) == 4095
    check(aligner, oracle.sw_port(), longest, 0)


def test_repeats_and_ties(aligner, sw_checker):
    """Homopolymers and tandem repeats make many equal scores: every tie-break of the reference has to be reproduced."""
    pairs = [(b"A" * 60, b"A" * 37), (b"AC" * 40, b"AC" * 31 + b"A"), (b"ACG" * 30, b"ACG" * 12 + b"T" + b"ACG" * 10),
             (b"T" * 100, b"T" * 100), (b"GATTACA" * 12, b"GATTACA" * 5 + b"GATACA" + b"GATTACA" * 4), (b"C" * 9, b"G" * 9),
             (b"ACGT" * 25, b"TGCA" * 25), (b"A" * 200 + b"C" * 50, b"C" * 50 + b"A" * 200)]
    for st in range(4):
        check(aligner, sw_checker, pairs, st)
    check(aligner, sw_checker, pairs, 0, (100, -50, -100, -20))


def test_shared_reference_and_cigar_growth(aligner, sw_checker):
    ref = sw.haplotype_pairs(3, 1, ref_len=400, per_ref=1)[0][0]
    alts = [a for _, a in sw.haplotype_pairs(4, 40, ref_len=400, per_ref=40, sub=0.02, indel=0.06)]
    pairs = [(ref, a) for a in alts]
    got = aligner.align(pairs, 0, cigar_cap=4)                     # forces the retry with a larger capacity
    for (r, a), (off, cig, _) in zip(pairs, got):
        assert (off, cig) == sw_checker.align(r, a, 0)
    assert max(len(c) for _, c, _ in got) > 4
    st = aligner.stats()
    assert st["pairs"] == 40 and st["cells"] == sum(len(r) * len(a) for r, a in pairs)


def test_invalid_inputs(aligner):
    with pytest.raises(sw.SwError):
        aligner.align([(b"ACGT", b"")], 0)
    with pytest.raises(sw.SwError):
        aligner.align([(b"ACGT", b"ACGT")], 7)
    with pytest.raises(sw.SwError):
        aligner.align([(b"A" * 5000, b"ACGT")], 0)
    assert aligner.align([], 0) == []


def test_cpp_host_layer(built, sw_checker, tmp_path):
    """htc-sw/: the reference's host entry points (SWPairwiseAlignmentMultiBatch, FalconSWFPGA_run, single pair) on the GPU --
    the self-checking bench, and a pair file whose output is compared with the oracle."""
    import subprocess
    root = ROOT
    exe = os.path.join(root, "htc-sw", "bin", "sw_host")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "htc-sw")], check=True, stdout=subprocess.DEVNULL)
    p = subprocess.run([exe, "cuda:0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "0 failures" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    pairs = sw.haplotype_pairs(31, 40, ref_len=(50, 400), per_ref=5)
    lines, want = [], []
    for k, (r, a) in enumerate(pairs):
        st = k % 4
        lines.append(f"{st} {r.decode()} {a.decode()}")
        off, cig = sw_checker.align(r, a, st)
        want.append(f"{off} {sw.cigar_string(cig)}")
    (tmp_path / "pairs.txt").write_text("\n".join(lines) + "\n")
    p = subprocess.run([exe, "cuda:0", str(tmp_path / "pairs.txt")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    assert p.stdout.split("\n")[:len(want)] == want


def test_fragmented_alignments_overflow_the_cigar_estimate(built, sw_checker):
    """The compact CIGAR array is sized for 16 elements per pair; noisy long alternates need several times that.  The engine
    notices, repeats the batch with room for everything, and remembers (no worst-case n_pairs x cigar_cap buffers)."""
    a = sw.SmithWaterman(0)
    pairs = sw.haplotype_pairs(77, 48, ref_len=(900, 1400), per_ref=6, sub=0.05, indel=0.04)
    got = a.align(pairs, 0, cigar_cap=512)
    assert np.mean([len(c) for _, c, _ in got]) > 20, "not fragmented enough to overflow the estimate"
    for (r, alt), (off, cig, _) in zip(pairs, got):
        assert (off, cig) == sw_checker.align(r, alt, 0)
    # a second, ordinary batch on the same context
    check(a, sw_checker, sw.haplotype_pairs(78, 64, ref_len=(200, 400), per_ref=8), 0)
    a.close()
