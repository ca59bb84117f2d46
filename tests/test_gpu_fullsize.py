"""BASELINE.json's full sizes through size-independent properties (the oracle would take minutes there), plus an
oracle check on a random sample of pairs."""
import numpy as np
import pytest

from acc_genomics_b200 import synth
from acc_genomics_b200.batch import Batch
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def permuted(b: Batch, rperm, hperm) -> Batch:
    return Batch.from_lists([b.read(int(k)) for k in rperm], [b.haplotype(int(k)) for k in hperm])


@pytest.mark.parametrize("cfg", [2, 3, 4])
def test_full_config_properties(engine, checker, cfg):
    b = synth.config(cfg)[0]
    raw, out, mask = engine.forward(b)
    assert np.isfinite(raw).all() and (raw >= 0).all() and not np.isnan(out).any()
    assert np.array_equal(mask, raw < np.float32(1e-28))
    assert (out <= 0).all()                                   # log10 of a probability

    # (1) pairs are independent: permuting reads and haplotypes permutes the result bit for bit -- different
    #     warps, lanes, task cuts and neighbours for every pair
    rng = np.random.Generator(np.random.PCG64(100 + cfg))
    rp, hp = rng.permutation(b.num_read), rng.permutation(b.num_hap)
    raw2, out2, mask2 = engine.forward(permuted(b, rp, hp))
    assert_bits_equal(raw2, raw[np.ix_(rp, hp)], "permutation invariance (raw)")
    assert_bits_equal(out2, out[np.ix_(rp, hp)], "permutation invariance (log10)")

    # (2) splitting the region into two regions of one job changes nothing
    half = b.num_read // 2
    parts = [b.slice_reads(0, half), b.slice_reads(half, b.num_read)]
    engine.stage(parts); engine.launch()
    raw3 = engine.fetch_raw().reshape(b.num_read, b.num_hap)
    assert_bits_equal(raw3, raw, "region split invariance")

    # (3) a random sample of pairs against the oracle, bit for bit
    for _ in range(48):
        i, j = int(rng.integers(b.num_read)), int(rng.integers(b.num_hap))
        f = checker.f32(*b.read(i), b.haplotype(j))
        assert np.float32(f).view(np.uint32) == raw[i, j].view(np.uint32)
        if mask[i, j]:
            d = checker.f64(*b.read(i), b.haplotype(j))
            lic = checker.log10_ic()[1]
            assert np.float64(np.log10(d) - lic).view(np.uint64) == out[i, j].view(np.uint64) or (d == 0 and np.isinf(out[i, j]))


def test_config5_stream_sample(engine, checker):
    """A 2 % slice of the 10^7-pair stream (50 regions) as one job; three regions checked in full against the oracle."""
    regs = synth.config(5, scale=0.02)
    j = engine.stage(regs); engine.launch()
    raw = engine.fetch_raw(); out, nfb = engine.fetch_log10()
    assert j["pairs"] == 50 * 4000 and np.isfinite(raw).all()
    offs = np.cumsum([0] + [r.num_pairs for r in regs])
    for k in (0, 17, 49):
        raw_r, out_r, fb_r = checker.batch(regs[k], threads=8)
        assert_bits_equal(raw[offs[k]:offs[k + 1]].reshape(raw_r.shape), raw_r, f"region {k} raw")
        assert_bits_equal(out[offs[k]:offs[k + 1]].reshape(out_r.shape), out_r, f"region {k} log10")
