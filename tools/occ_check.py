"""GPU box: float-pass time of configs 1, 2, 4 and a 25-region job of config 5 with the planner's own variant choice
(used to compare builds with different occupancy thresholds, PAIRHMM_B200_LIB=...)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine
eng = PairHMMEngine(0)
for cfg, scale in ((1, 1.0), (2, 1.0), (4, 1.0), (5, 0.01)):
    regs = synth.config(cfg, scale=scale)
    cells = sum(b.num_cells for b in regs)
    eng.stage(regs)
    for _ in range(3): eng.launch()
    eng.sync()
    ts = []
    for _ in range(9):
        eng.launch(); eng.sync(); st = eng.stats(); ts.append((st["ms_f32"], st["ms_fallback"]))
    ts.sort()
    print(f"cfg{cfg} f32 {ts[0][0]:.3f} ms {cells / ts[0][0] * 1e-6:.0f} GCUPS  fallback {ts[0][1]:.3f} ms  step {cells / sum(ts[0]) * 1e-6:.0f} GCUPS")
