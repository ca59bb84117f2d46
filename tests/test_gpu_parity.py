"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on identical inputs.

Bar (BASELINE.json north_star): raw float likelihoods BIT-EXACT with the reference's AVX code built without FMA
contraction, float->double fallback decision identical on every pair, final log10 within 1e-5 relative -- in fact
bit-exact too, because the double re-run reproduces the reference's arithmetic and log10 is taken by the host libm."""
import numpy as np
import pytest

from acc_genomics_b200 import synth
from acc_genomics_b200.batch import Batch, serialize_haps, serialize_reads
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5      # tolerance north_star states for log10 likelihoods; asserted in addition to bit equality


def check(engine, checker, b, what):
    raw, out, mask = engine.forward(b)
    raw_r, out_r, fb_r = checker.batch(b, threads=8)
    assert_bits_equal(raw, raw_r, f"{what}: raw float likelihood")
    assert np.array_equal(mask, fb_r), f"{what}: fallback decision"
    fin = np.isfinite(out_r)
    assert np.array_equal(np.isinf(out), ~fin), f"{what}: -inf pattern"
    assert (np.abs(out[fin] - out_r[fin]) <= REL_TOL * np.abs(out_r[fin])).all(), f"{what}: log10 beyond 1e-5 relative"
    assert_bits_equal(out, out_r, f"{what}: final log10")
    return raw, out, mask


def test_golden_vectors(engine, golden, widening):
    """Committed vectors minted from the reference's own AVX code (tests/golden/make_golden.py)."""
    cases, _ = golden
    for name, (b, raw_bits, log10_bits, mask) in cases.items():
        raw, out, m = engine.forward(b)
        assert_bits_equal(raw, raw_bits.view(np.float32), f"{name} raw")
        assert np.array_equal(m, mask), f"{name} mask"
        assert_bits_equal(out, log10_bits.view(np.float64), f"{name} log10")
    st = engine.stats()
    assert st["kernel_launches"] >= 3          # row parameters, float pass, double pass


def test_flush_to_zero_path_taken(engine, golden):
    """Results below 2^-800 are recomputed with x86 flush-to-zero emulated; some golden values are exactly 0."""
    cases, _ = golden
    b, _, log10_bits, _ = cases["deep_underflow"]
    engine.forward(b)
    st = engine.stats()
    assert st["fallback_pairs"] == b.num_pairs and st["flush_pairs"] > 0
    assert np.isinf(log10_bits.view(np.float64)).sum() > 0


@pytest.mark.parametrize("cfg,scale,seed", [(1, 1.0, 1), (2, 0.06, 2), (3, 0.06, 3), (4, 0.05, 4), (2, 0.04, 77), (3, 0.04, 78)])
def test_baseline_configs_vs_oracle(engine, checker, cfg, scale, seed):
    b = synth.config(cfg, seed=seed, scale=scale)[0]
    raw, out, mask = check(engine, checker, b, f"cfg{cfg}")
    if cfg == 3:
        assert 0.3 < mask.mean() < 0.7          # underflow-heavy: the fallback path carries real weight


def test_config5_multi_region_job(engine, checker):
    """Many regions in one job (the whole-genome stream), mixed read lengths -> several kernel variants."""
    regs = synth.config(5, scale=0.006)
    j = engine.stage(regs); engine.launch()
    raw = engine.fetch_raw(); out, nfb = engine.fetch_log10(); mask = engine.fetch_fallback_mask()
    pos = 0
    for b in regs:
        raw_r, out_r, fb_r = checker.batch(b, threads=8)
        n = b.num_pairs
        assert_bits_equal(raw[pos:pos + n].reshape(raw_r.shape), raw_r, "cfg5 raw")
        assert np.array_equal(mask[pos:pos + n].reshape(fb_r.shape), fb_r)
        assert_bits_equal(out[pos:pos + n].reshape(out_r.shape), out_r, "cfg5 log10")
        pos += n
    assert pos == j["pairs"] and nfb == int(mask.sum())


def test_every_entry_point_agrees(engine, checker):
    b = synth.config(3, seed=5, scale=0.02)[0]
    raw_r, out_r, fb_r = checker.batch(b, threads=8)
    rs, hs = serialize_reads(b), serialize_haps(b)
    assert_bits_equal(engine.forward_raw_serialized(rs, hs, b.num_pairs), raw_r, "pmm_forward_raw_serialized")
    out, nfb = engine.forward_log10_serialized(rs, hs, b.num_pairs)
    assert_bits_equal(out, out_r, "pmm_forward_log10_serialized"); assert nfb == int(fb_r.sum())
    out, nfb = engine.forward_log10_structs(b)
    assert_bits_equal(out, out_r, "pmm_forward_log10"); assert nfb == int(fb_r.sum())


@pytest.fixture(params=["on", "off"])
def widening(request, engine):
    """Small jobs get 16 or 32 lanes per read by default; "off" keeps the variants a large job of the same reads would use
    (so the narrow variants see the small edge cases too)."""
    engine.set_option("small_job_widening", request.param)
    yield request.param
    engine.set_option("small_job_widening", "on")


def test_all_small_lengths(engine, checker, widening):
    """Read lengths 1..70 x haplotype lengths 1..40: every boundary-row count, lane count and short-haplotype
    (shorter than the wavefront) combination of the small variants."""
    rng = np.random.Generator(np.random.PCG64(21))
    b = synth.region(rng, list(range(1, 71)), list(range(1, 41)))
    check(engine, checker, b, "small lengths")


def test_lengths_around_every_variant_boundary(engine, checker, widening):
    rng = np.random.Generator(np.random.PCG64(22))
    lens = sorted({k * w + d for w in (8, 16, 32) for k in range(4, 17) for d in (-2, -1, 0, 1)})
    lens = [x for x in lens if x <= 520]
    b = synth.region(rng, lens, [37, 150, 301], decoy_frac=0.34)
    check(engine, checker, b, "variant boundaries")


def test_long_reads_and_long_haplotypes(engine, checker):
    """Beyond the reference FPGA's limits (MAX_READ_LEN 192, MAX_HAP_LEN 1024, PairHMMFpgaInterface.h:16-17):
    multi-stripe float kernel (> 511 bases), multi-stripe double kernel (> 191 bases), 3 kb haplotypes."""
    rng = np.random.Generator(np.random.PCG64(23))
    b = synth.region(rng, [190, 191, 192, 193, 383, 384, 511, 512, 513, 1023, 1024, 1500], [2000, 3000, 1800], decoy_frac=0.34)
    raw, out, mask = check(engine, checker, b, "long reads")
    assert mask.any() and not mask.all()


def test_odd_bytes(engine, checker):
    """N matches everything, any non-ACGTN byte is 'A' (ConvertChar), qualities are used modulo 128."""
    rng = np.random.Generator(np.random.PCG64(24))
    b = synth.region(rng, [30, 77, 151, 151, 200], [80, 180, 400, 401])
    for arr in (b.rs, b.hap):
        m = rng.random(arr.size) < 0.1
        arr[m] = rng.choice(np.frombuffer(b"NNNacgtnRYKM-*. \x00\x01\x7f\x80\xfe", dtype=np.uint8), size=int(m.sum()))
    for arr in (b.q, b.i, b.d, b.c):
        m = rng.random(arr.size) < 0.15
        arr[m] = rng.integers(0, 256, size=int(m.sum()), dtype=np.uint8)
    check(engine, checker, b, "odd bytes")
    allN = Batch.from_lists([(b"N" * 40, bytes([30]) * 40, bytes([40]) * 40, bytes([40]) * 40, bytes([10]) * 40)], [b"N" * 60, b"ACGT" * 15])
    check(engine, checker, allN, "all N")


def test_single_pair_and_single_row(engine, checker):
    one = Batch.from_lists([(b"A", b"\x1e", b"\x28", b"\x28", b"\x0a")], [b"C"])
    check(engine, checker, one, "1x1 bases")
    rng = np.random.Generator(np.random.PCG64(25))
    check(engine, checker, synth.region(rng, [151], [450]), "one read, one haplotype")
    check(engine, checker, synth.region(rng, [151], [450] * 70), "one read, many haplotypes")
    check(engine, checker, synth.region(rng, [151] * 70, [450]), "many reads, one haplotype")


def test_invalid_inputs_are_rejected(engine):
    from acc_genomics_b200.engine import PmmError, PMM_ERR_INVALID, PMM_ERR_STATE
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    bad = Batch(np.array([0, 0, 4], dtype=np.int32), acgt, acgt, acgt, acgt, acgt, np.array([0, 4], dtype=np.int32), acgt)
    with pytest.raises(PmmError) as e:
        engine.stage([bad])
    assert e.value.code == PMM_ERR_INVALID
    with pytest.raises(PmmError) as e:
        engine.launch()                       # nothing staged after the failed stage
    assert e.value.code == PMM_ERR_STATE
    with pytest.raises(PmmError):
        engine.forward_raw_serialized(b"\x05\x00\x00\x00\x10", b"\x01\x00\x00\x00", 16)   # truncated wire format
    for key, value in (("run_tiers", "x"), ("run_tiers", "2,0,4"), ("run_tiers", "99"), ("small_job_widening", "maybe"),
                       ("tasks_per_warp", "0"), ("no_such_option", "1")):
        with pytest.raises(PmmError):
            engine.set_option(key, value)
    engine.set_option("run_tiers", "2,40,4"); engine.set_option("small_job_widening", "on")     # the defaults, accepted


def test_repeat_launch_is_idempotent(engine):
    b = synth.config(3, seed=9, scale=0.03)[0]
    engine.stage([b]); engine.launch()
    a = engine.fetch_raw(); oa, _ = engine.fetch_log10()
    engine.launch(); engine.launch()
    c = engine.fetch_raw(); oc, _ = engine.fetch_log10()
    assert_bits_equal(a, c, "raw after relaunch"); assert_bits_equal(oa, oc, "log10 after relaunch")


@pytest.mark.parametrize("tasks_per_warp,max_run", [(1, 64), (64, 1), (2, 3), (6, 6)])
def test_double_rerun_task_shapes(engine, checker, golden, tasks_per_warp, max_run):
    """The double re-run groups the failing haplotypes of a read into tasks (a list of haplotypes, not a run of
    neighbours).  However the list is cut -- everything in one task, one pair per task, odd runs -- the results are the
    reference's: haplotypes shorter than the warp (a lane crosses two separators in one window), results deep enough for
    the flush-to-zero repeat inside a multi-haplotype task, reads of more than one stripe, failing and passing
    haplotypes interleaved."""
    engine.set_option("f64_tasks_per_warp", tasks_per_warp)
    engine.set_option("f64_max_run", max_run)
    try:
        rng = np.random.Generator(np.random.PCG64(31))
        # unrelated reads against random haplotypes with expensive gaps (deep underflow, some results exactly 0) ...
        hl = list(range(1, 41, 3)) + [64, 65, 300, 500, 7, 2, 450, 33]
        w = synth.region(rng, [20, 60, 120, 140, 151, 151, 159, 160, 175, 200, 250, 300], hl)
        w.hap[:] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=w.hap.size)]
        w.q[:] = 60
        w.i[:] = 90; w.d[:] = 90; w.c[:] = 90
        raw, out, mask = check(engine, checker, w, "deep underflow, ragged haplotypes")
        assert mask.sum() > w.num_pairs // 2 and engine.stats()["flush_pairs"] > 0
        # ... and a region where about half of the haplotypes of every read fail, interleaved with the ones that pass
        b = synth.config(3, seed=41, scale=0.05)[0]
        hp = rng.permutation(b.num_hap)
        b = Batch.from_lists([b.read(k) for k in range(b.num_read)], [b.haplotype(int(k)) for k in hp])
        raw, out, mask = check(engine, checker, b, "interleaved failing haplotypes")
        assert 0.3 < mask.mean() < 0.7
        cases, _ = golden
        gb, raw_bits, log10_bits, gmask = cases["deep_underflow"]
        raw, out, m = engine.forward(gb)
        assert_bits_equal(out, log10_bits.view(np.float64), "golden deep_underflow log10")
    finally:
        engine.set_option("f64_tasks_per_warp", 6)
        engine.set_option("f64_max_run", 6)


def test_gkl_testcase_batch(engine, checker):
    """pmm_forward_log10_testcases: GKL's per-pair `testcase` array (host_type.h:69-73).  A full cross product, two
    regions back to back, a ragged tail and a lone pair in one call: out[k] belongs to tc[k]."""
    a = synth.config(3, seed=51, scale=0.012)[0]            # 12 reads x 64 haplotypes, about half of them fall back
    b = synth.config(5, seed=52, scale=0.0004)[0].slice_reads(0, 9)
    want, pairs, nfb_want = [], [], 0
    for reg in (a, b):
        reads = [tuple(bytes(x) for x in reg.read(k)) for k in range(reg.num_read)]
        haps = [bytes(reg.haplotype(k)) for k in range(reg.num_hap)]
        _, out_r, fb_r = checker.batch(reg, threads=8)
        for i, r in enumerate(reads):
            for j, h in enumerate(haps):
                pairs.append((r, h)); want.append(out_r[i, j]); nfb_want += int(fb_r[i, j])
    # ragged tail: one read of region a against three of its haplotypes only, then a lone pair of region b
    reads_a = [tuple(bytes(x) for x in a.read(k)) for k in range(2)]
    haps_a = [bytes(a.haplotype(k)) for k in (5, 1, 40)]
    _, out_a, fb_a = checker.batch(a, threads=8)
    for j, hj in zip((5, 1, 40), haps_a):
        pairs.append((reads_a[1], hj)); want.append(out_a[1, j]); nfb_want += int(fb_a[1, j])
    rb = tuple(bytes(x) for x in b.read(3)); hb = bytes(b.haplotype(7))
    _, out_b, fb_b = checker.batch(b, threads=8)
    pairs.append((rb, hb)); want.append(out_b[3, 7]); nfb_want += int(fb_b[3, 7])
    out, nfb = engine.forward_log10_testcases(pairs)
    assert_bits_equal(out, np.array(want), "testcase batch")
    assert nfb == nfb_want
    st = engine.stats()
    assert st["pairs"] == len(pairs)
    from acc_genomics_b200.engine import PmmError
    with pytest.raises(PmmError):
        engine.forward_log10_testcases([((b"", b"", b"", b"", b""), b"ACGT")])      # read of length 0


def test_stage_launch_stage_without_fetch(engine, checker):
    """The pinned input arena is reused by the next stage: it must wait for the previous job's copy."""
    b1 = synth.config(2, seed=61, scale=0.05)[0]
    b2 = synth.config(1, seed=62, scale=0.5)[0]
    engine.stage([b1]); engine.launch()
    engine.stage([b2]); engine.launch()
    out, _ = engine.fetch_log10()
    assert_bits_equal(out.reshape(b2.num_read, b2.num_hap), checker.batch(b2, threads=8)[1], "second job")
    # fetches of one launch share their copies: any order, any repetition
    raw1 = engine.fetch_raw(); m = engine.fetch_fallback_mask(); out2, _ = engine.fetch_log10(); raw2 = engine.fetch_raw()
    assert_bits_equal(raw1, raw2, "raw twice"); assert_bits_equal(out, out2, "log10 twice")
    assert np.array_equal(m, raw1 < np.float32(1e-28))


@pytest.mark.parametrize("key,value", [("priority", "0"), ("priority", "3"), ("sync", "auto"), ("sync", "hybrid"), ("sync", "block")])
def test_stream_priority_and_wait_modes_do_not_change_results(built, checker, key, value):
    from acc_genomics_b200.engine import PairHMMEngine
    e = PairHMMEngine(0)
    e.set_option(key, value)
    b = synth.config(3, seed=91, scale=0.04)[0]
    raw, out, mask = e.forward(b)
    raw_r, out_r, fb_r = checker.batch(b, threads=8)
    assert_bits_equal(raw, raw_r, f"{key}={value} raw"); assert_bits_equal(out, out_r, f"{key}={value} log10")
    e.close()
