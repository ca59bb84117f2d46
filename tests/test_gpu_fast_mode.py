"""Fast mode ("mode" = "fast"): the float cell update contracted to 4 FMUL + 4 FFMA.  north_star's bar for it:
log10 likelihoods within 1e-5 relative of the reference's AVX implementation, and the float-versus-double fallback
decision bit-identical.  The decision is protected by an exact re-check of every pair whose fast result lies in a
guard band around 1e-28f; widening the band to almost everything must reproduce the exact kernel bit for bit."""
import numpy as np
import pytest

from acc_genomics_b200 import synth
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5      # |log10 - ref| <= REL_TOL * |ref|   (BASELINE.json north_star)


@pytest.fixture(scope="module")
def fast_engine(built):
    from acc_genomics_b200.engine import PairHMMEngine
    e = PairHMMEngine(0)
    e.set_option("mode", "fast")
    yield e
    e.close()


@pytest.mark.parametrize("cfg,scale", [(1, 1.0), (2, 0.05), (3, 0.05), (4, 0.06), (5, 0.002)])
def test_fast_mode_within_tolerance_and_same_decision(fast_engine, checker, cfg, scale):
    for b in synth.config(cfg, scale=scale):
        raw_r, out_r, fb_r = checker.batch(b, threads=8)
        raw, out, mask = fast_engine.forward(b)
        assert np.array_equal(mask, fb_r), "float->double decision differs from the reference"
        # pairs that fell back were computed by the (exact) double kernel
        assert_bits_equal(out[fb_r], out_r[fb_r], "double re-run")
        fin = np.isfinite(out_r)
        assert np.array_equal(np.isfinite(out), fin)
        err = np.abs(out[fin] - out_r[fin]) / np.abs(out_r[fin])
        assert err.max() <= REL_TOL, f"log10 off by {err.max():.3g} relative"
        keep = ~fb_r
        rel = np.abs(raw[keep].astype(np.float64) - raw_r[keep]) / raw_r[keep]
        assert rel.max() < 2e-5, f"raw float off by {rel.max():.3g}"          # a few hundred ulp at most


def test_wide_guard_band_reproduces_exact_results(built, checker):
    """guard = 0.999: every result in [1e-31, 2e-28) is re-run by the exact kernel -> bit-identical there."""
    from acc_genomics_b200.engine import PairHMMEngine
    e = PairHMMEngine(0)
    e.set_option("mode", "fast"); e.set_option("guard", 0.999)
    b = synth.config(3, scale=0.06)[0]                                    # half of the pairs are around / below the threshold
    raw_r, out_r, fb_r = checker.batch(b, threads=8)
    raw, out, mask = e.forward(b)
    st = e.stats()
    band = (raw_r >= np.float32(1e-31)) & (raw_r < np.float32(1.9e-28))
    assert band.sum() > 0 and st["recheck_pairs"] >= band.sum() - 2        # band edges are fuzzy by construction
    assert np.array_equal(mask, fb_r)
    assert_bits_equal(raw[band], raw_r[band], "re-checked raw results")
    assert_bits_equal(out[band | fb_r], out_r[band | fb_r], "re-checked / double log10")
    e.close()


def test_exact_mode_has_no_recheck(engine):
    b = synth.config(3, scale=0.03)[0]
    engine.forward(b)
    assert engine.stats()["recheck_pairs"] == 0


def test_mode_option_validation(built):
    from acc_genomics_b200.engine import PairHMMEngine, PmmError
    e = PairHMMEngine(0)
    with pytest.raises(PmmError):
        e.set_option("mode", "sloppy")
    with pytest.raises(PmmError):
        e.set_option("guard", 1.5)
    e.set_option("mode", "exact")
    e.close()
