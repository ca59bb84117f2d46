"""BASELINE.json's full sizes: every pair of configs 2, 3 and 4 and a tenth of config 5 against the oracle, bit for bit
(the checker runs the whole batch on all host cores in about a second), plus size-independent properties."""
import os

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

    # (3) every pair of the configuration against the oracle, bit for bit: raw floats, the float-versus-double decision
    #     and the final log10 (the contract of FalconPairHMM::computePairhmmAVX, xlnx/host/FalconPairHMM.cpp:69-95)
    raw_r, out_r, fb_r = checker.batch(b, threads=os.cpu_count() or 1)
    assert_bits_equal(raw, raw_r, f"cfg{cfg} raw, all {b.num_pairs} pairs")
    assert np.array_equal(mask, fb_r), f"cfg{cfg} fallback decision"
    assert_bits_equal(out, out_r, f"cfg{cfg} log10, all {b.num_pairs} pairs")


def test_config5_stream_tenth(engine, checker):
    """A tenth of the 10^7-pair stream (250 regions, 10^6 pairs) as jobs of 25 regions; every region checked in full
    against the oracle."""
    regs = synth.config(5, scale=0.1)
    assert len(regs) == 250
    threads = os.cpu_count() or 1
    nfb_total = 0
    for j0 in range(0, len(regs), 25):
        job = regs[j0:j0 + 25]
        j = engine.stage(job); engine.launch()
        raw = engine.fetch_raw(); out, nfb = engine.fetch_log10()
        mask = engine.fetch_fallback_mask()
        assert j["pairs"] == 25 * 4000
        offs = np.cumsum([0] + [r.num_pairs for r in job])
        for k, r in enumerate(job):
            raw_r, out_r, fb_r = checker.batch(r, threads=threads)
            assert_bits_equal(raw[offs[k]:offs[k + 1]].reshape(raw_r.shape), raw_r, f"region {j0 + k} raw")
            assert np.array_equal(mask[offs[k]:offs[k + 1]].reshape(fb_r.shape), fb_r), f"region {j0 + k} decision"
            assert_bits_equal(out[offs[k]:offs[k + 1]].reshape(out_r.shape), out_r, f"region {j0 + k} log10")
        nfb_total += nfb
    assert nfb_total > 0
