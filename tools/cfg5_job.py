"""GPU box: one 25-region job of config 5 -- launches, float-pass time, which variants the planner chose."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acc_genomics_b200 import synth, engine
from acc_genomics_b200.engine import PairHMMEngine
regs = synth.config(5, scale=0.01)
cells = sum(b.num_cells for b in regs)
plan = engine.plan(regs)
c = collections.Counter((t["K"], t["W"]) for t in plan)
print("tasks per variant:", sorted(c.items()))
eng = PairHMMEngine(0)
for opt in ("0,0", "-1,0", "19,8"):   # planner's choice (rare variants consolidated), every variant its own launch, one variant
    eng.set_option("force_variant", opt)
    eng.stage(regs)
    for _ in range(3): eng.launch()
    eng.sync()
    ts = []
    for _ in range(7):
        eng.launch(); eng.sync(); st = eng.stats(); ts.append((st["ms_f32"], st["ms_fallback"]))
    ts.sort()
    print("force", opt, "launches", st["kernel_launches"], "ms_f32 %.3f ms_fallback %.3f" % ts[0], "f32 GCUPS %.0f step GCUPS %.0f" % (cells / ts[0][0] * 1e-6, cells / sum(ts[0]) * 1e-6), "fallback", st["fallback_pairs"])
