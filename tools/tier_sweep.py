"""Tuning sweep on a GPU box: the float pass of configs 2, 3, 4 and a config-5 job under graded runs
(pmm_set_option "run_tiers" = "depth,share,top"; depth 0 = runs of one size).
Every result is compared bit for bit with the ungraded plan.  Writes gpurun_out/tier_sweep.jsonl."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine

eng = PairHMMEngine(0)
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/tier_sweep.jsonl", "a")
depths = sys.argv[1].split() if len(sys.argv) > 1 else ["0", "2,40,4", "1,40,4", "3,40,4", "2,33,4", "2,50,4", "2,40,8", "2,60,8"]
for cfg, scale in ((2, 1.0), (3, 1.0), (4, 1.0), (5, 0.01), (1, 1.0)):
    bs = synth.config(cfg, scale=scale)
    cells = sum(b.num_cells for b in bs)
    ref = None
    for d in depths:
        eng.set_option("run_tiers", d)
        eng.stage(bs)
        for _ in range(3):
            eng.launch()
        eng.sync()
        ms, msd = [], []
        for _ in range(9):
            eng.launch(); eng.sync(); st = eng.stats(); ms.append(st["ms_f32"]); msd.append(st["ms_fallback"])
        raw = eng.fetch_raw().view(np.uint32)
        if ref is None:
            ref = raw.copy()
        rec = dict(cfg=cfg, scale=scale, run_tiers=d, ms_f32=round(float(np.median(ms)), 4), ms_f32_min=round(min(ms), 4),
                   ms_fallback=round(float(np.median(msd)), 4), gcups_f32=round(cells / (np.median(ms) * 1e-3) * 1e-9, 1),
                   tasks=eng.stats()["f32_tasks"], bit_equal=bool(np.array_equal(raw, ref)))
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n"); out.flush()
eng.set_option("run_tiers", "2,40,4")
print("fp32 peak", eng.measure_fp32_peak())
