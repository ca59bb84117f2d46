"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every kernel variant
family once -- W = 8, 16, 32, striped float, double 5 and 6 rows, flush, fast mode + re-check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine
rng = np.random.Generator(np.random.PCG64(7))
eng = PairHMMEngine(0)
jobs = [synth.config(3, scale=0.008),                                             # 151-base reads, fallback heavy
        [synth.region(rng, [40, 70, 101, 130, 200, 250, 300], [90, 333, 64, 700])],   # many variants in one job
        [synth.region(rng, [600, 700], [1500, 900], decoy_frac=0.5)]]             # striped float + multi-stripe double
for mode, guard in (("exact", None), ("fast", None), ("fast", 0.999)):
    eng.set_option("mode", mode)
    if guard is not None:
        eng.set_option("guard", guard)
    for j in jobs:
        eng.stage(j); eng.launch()
        raw = eng.fetch_raw(); out, nfb = eng.fetch_log10()
        print(mode, guard, len(raw), nfb, eng.stats()["recheck_pairs"], eng.stats()["flush_pairs"])
print("done")
