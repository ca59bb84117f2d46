"""GPU box: randomized check of the planner's task shapes on jobs large enough for graded runs.  Random multi-region jobs
(mixed read lengths, so several kernel variants and rare-variant merging; haplotype counts from 3 to 80) are run under
random `run_tiers`, `tasks_per_warp` and `small_job_widening` settings; every result must equal, bit for bit, the one the
plainest plan gives (one haplotype per task, no widening), and that one is checked against the oracle.
    python tools/fuzz_plan_gpu.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.Generator(np.random.PCG64(seed))
chk = oracle.reference() or oracle.port()
eng = PairHMMEngine(0)
t0 = time.time(); it = 0; pairs = 0; shapes = set()


def rand_job():
    regs = []
    for _ in range(int(rng.integers(1, 7))):
        nr, nh = int(rng.integers(20, 400)), int(rng.integers(3, 80))
        style = int(rng.integers(0, 4))
        if style == 0:   rl = np.full(nr, 151)
        elif style == 1: rl = np.where(rng.random(nr) < 0.15, rng.integers(40, 150, nr), 151)
        elif style == 2: rl = rng.integers(90, 260, nr)
        else:            rl = np.full(nr, int(rng.integers(60, 300)))
        hl = rng.integers(int(rng.integers(20, 300)), int(rng.integers(320, 900)), nh)
        regs.append(synth.region(rng, [int(x) for x in rl], [int(x) for x in hl], decoy_frac=float(rng.choice([0, 0, 0.4])),
                                 low_read_frac=float(rng.choice([0, 0.2]))))
    return regs


def run(regs):
    eng.stage(regs); eng.launch()
    raw = eng.fetch_raw().view(np.uint32).copy(); out, nfb = eng.fetch_log10()
    return raw, out.view(np.uint64).copy(), nfb, eng.stats()["f32_tasks"]


while time.time() - t0 < budget:
    it += 1
    regs = rand_job()
    pairs += sum(b.num_pairs for b in regs)
    eng.set_option("run_tiers", "0"); eng.set_option("small_job_widening", "off"); eng.set_option("tasks_per_warp", 16)
    raw0, out0, nfb0, n0 = run(regs)
    if it % 4 == 1:                                               # the plain plan against the oracle
        want = [chk.batch(b, threads=os.cpu_count()) for b in regs]
        assert np.array_equal(raw0, np.concatenate([w[0].ravel() for w in want]).view(np.uint32)), ("oracle raw", it)
        assert np.array_equal(out0, np.concatenate([w[1].ravel() for w in want]).view(np.uint64)), ("oracle log10", it)
    for _ in range(3):
        tiers = f"{int(rng.integers(1, 5))},{int(rng.choice([20, 40, 60, 100]))},{int(rng.choice([2, 3, 4, 8]))}"
        eng.set_option("run_tiers", tiers)
        eng.set_option("small_job_widening", str(rng.choice(["on", "off"])))
        eng.set_option("tasks_per_warp", int(rng.choice([1, 4, 16, 64])))
        raw, out, nfb, n = run(regs)
        shapes.add((tiers, n < n0))
        assert np.array_equal(raw, raw0) and np.array_equal(out, out0) and nfb == nfb0, ("plan changed the result", it, tiers)
eng.set_option("run_tiers", "2,40,4"); eng.set_option("small_job_widening", "on")
print(f"plan fuzz ok: {it} jobs, {pairs} pairs, {sum(1 for s in shapes if s[1])} graded settings took effect, {time.time() - t0:.0f} s, "
      f"seed {seed}, checker {chk.kind}")
