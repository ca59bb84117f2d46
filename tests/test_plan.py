"""Host logic: cutting a job into warp-tasks (pmm_plan_flat, no GPU needed)."""
import numpy as np
import pytest

from acc_genomics_b200 import synth
from acc_genomics_b200.batch import Batch


def coverage(batches, tasks):
    """Every (read, hap) pair of every region is produced exactly once, at the right output index."""
    total = sum(b.num_pairs for b in batches)
    seen = np.zeros(total, dtype=np.int32)
    # global read/hap index -> (region, local index)
    rfirst = np.cumsum([0] + [b.num_read for b in batches]); hfirst = np.cumsum([0] + [b.num_hap for b in batches])
    ofirst = np.cumsum([0] + [b.num_pairs for b in batches])
    for t in tasks:
        for g, r in enumerate(t["read"]):
            reg = int(np.searchsorted(rfirst, r, side="right") - 1)
            b = batches[reg]
            assert hfirst[reg] <= t["hap_first"] and t["hap_first"] + t["num_hap"] <= hfirst[reg + 1]
            for n in range(t["num_hap"]):
                want = ofirst[reg] + (r - rfirst[reg]) * b.num_hap + (t["hap_first"] + n - hfirst[reg])
                assert t["out_base"][g] + n == want
                seen[want] += 1
    assert (seen == 1).all()


@pytest.mark.parametrize("cfg,scale", [(1, 1.0), (2, 0.2), (4, 0.25), (5, 0.004)])
def test_every_pair_covered_once(built, cfg, scale):
    from acc_genomics_b200 import engine
    batches = synth.config(cfg, scale=scale)
    tasks = engine.plan(batches)
    coverage(batches, tasks)
    lens = np.concatenate([b.read_lens for b in batches])
    for t in tasks:
        assert 1 <= len(t["read"]) <= 32 // t["W"]
        for r in t["read"]:
            assert t["striped"] or lens[r] + 1 <= t["K"] * t["W"], "read does not fit its row block"


def test_variant_choice(built):
    from acc_genomics_b200 import engine

    def variant(length):
        # a job large enough to fill the GPU: small ones get wider lanes, see test_small_jobs_get_wider_lanes
        b = synth.region(np.random.Generator(np.random.PCG64(1)), [length] * 400, [length + 50] * 24)
        t = engine.plan(b)[0]
        return t["K"], t["W"], t["striped"]
    assert variant(151) == (19, 8, False)        # 4 reads per warp, 151 + 1 boundary row = all 152 rows used
    k, w, s = variant(250)
    assert not s and 251 <= k * w <= 256
    k, w, s = variant(101)
    assert not s and k * w >= 102 and k * w <= 112
    assert variant(511)[:2] == (16, 32) and not variant(511)[2]
    assert variant(512)[2] and variant(3000)[2]  # longer than 32 x 16 - 1: multi-stripe kernel
    k, w, s = variant(1)
    assert not s and w == 8


def test_queue_depth_follows_gpu_size(built):
    from acc_genomics_b200 import engine
    b = synth.config(2, scale=0.5)
    small = engine.plan(b, sm_count=16, tasks_per_warp=2)
    big = engine.plan(b, sm_count=148, tasks_per_warp=8)
    assert len(big) > len(small)
    coverage(b, small); coverage(b, big)


def test_rejects_malformed_jobs(built):
    from acc_genomics_b200 import engine
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    good = Batch.from_lists([(acgt, acgt, acgt, acgt, acgt)], [acgt])
    empty_read = Batch(np.array([0, 0, 4], dtype=np.int32), acgt, acgt, acgt, acgt, acgt, good.hap_off, good.hap)
    with pytest.raises(engine.PmmError):
        engine.plan(empty_read)
    empty_hap = Batch(good.read_off, acgt, acgt, acgt, acgt, acgt, np.array([0, 4, 4], dtype=np.int32), acgt)
    with pytest.raises(engine.PmmError):
        engine.plan(empty_hap)
    assert len(engine.plan(good)) == 1


def test_rare_variants_join_the_next_larger_launch(built):
    """A few short reads do not get a kernel launch of their own; a large share of short reads does."""
    from acc_genomics_b200 import engine
    rng = np.random.Generator(np.random.PCG64(7))
    haps = [450] * 40
    few = [synth.region(rng, [151] * 90 + [70, 80, 90, 100, 101, 110, 120, 130, 140, 150], haps) for _ in range(4)]
    tasks = engine.plan(few)
    assert {(t["K"], t["W"]) for t in tasks} == {(19, 8)}
    coverage(few, tasks)
    half = [synth.region(rng, [151] * 52 + [101] * 48, haps) for _ in range(40)]
    tasks = engine.plan(half)
    assert {(t["K"], t["W"]) for t in tasks} == {(19, 8), (13, 8)}
    coverage(half, tasks)


def test_small_jobs_get_wider_lanes(built):
    """Fewer tasks than the GPU has SMSPs: 16 or 32 lanes per read (shorter, more numerous tasks) instead of 8."""
    from acc_genomics_b200 import engine
    rng = np.random.Generator(np.random.PCG64(11))
    tiny = synth.region(rng, [151] * 10, [400] * 5)              # 3 groups x 5 haplotypes at 19 x 8
    tasks = engine.plan(tiny)
    assert {(t["K"], t["W"]) for t in tasks} == {(5, 32)} and len(tasks) == 50
    coverage([tiny], tasks)
    small = synth.region(rng, [151] * 100, [400] * 20)           # 25 groups x 20 = 500 tasks: 16 lanes
    tasks = engine.plan(small)
    assert {(t["K"], t["W"]) for t in tasks} == {(10, 16)}
    coverage([small], tasks)
    medium = synth.region(rng, [151] * 100, [400] * 40)          # 1000 tasks: the default
    assert {(t["K"], t["W"]) for t in engine.plan(medium)} == {(19, 8)}
    # the same job on a GPU an eighth the size is not small
    assert {(t["K"], t["W"]) for t in engine.plan(small, sm_count=16)} == {(19, 8)}


def test_graded_runs_cover_every_pair_and_put_long_runs_first(built):
    """Large jobs: a region's haplotypes are cut into runs of decreasing size (what a task pays once is paid rarely, the
    launch's tail is made of one-haplotype tasks).  Every pair still appears exactly once; within a launch the tasks are
    ordered longest first; a job with few tasks per warp keeps one haplotype per task."""
    from acc_genomics_b200 import engine
    big = synth.config(2)                                        # 250 groups x 64 haplotypes: 13.5 per resident warp
    tasks = engine.plan(big)
    coverage(big, tasks)
    sizes = sorted({t["num_hap"] for t in tasks})
    assert sizes[0] == 1 and sizes[-1] >= 3 and len(tasks) < 12000          # 16 000 without grading
    ones = sum(t["num_hap"] == 1 for t in tasks)
    assert 2000 <= ones <= 5000                                   # about two per resident warp (1 184), not all of them
    hap_len = np.diff(big[0].hap_off)
    cost = [int(hap_len[t["hap_first"]:t["hap_first"] + t["num_hap"]].sum()) + t["num_hap"] for t in tasks]
    cmax = max(cost)
    cls = [63 - c * 63 // cmax for c in cost]                     # the planner's 64 cost classes
    assert cls == sorted(cls), "tasks of a launch are queued longest first (by cost class)"
    few = synth.config(4)                                         # 256 groups x 32 long haplotypes: 7 per warp
    assert {t["num_hap"] for t in engine.plan(few)} == {1}
    job = synth.config(5, scale=0.01)                             # 25 regions of 100 x 40
    tj = engine.plan(job)
    coverage(job, tj)
    assert max(t["num_hap"] for t in tj) == 4 and min(t["num_hap"] for t in tj) == 1
