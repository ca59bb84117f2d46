"""GPU box: exact vs fast mode on the four single-region configs -- device time per step, recheck counts, max error."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine
res = {}
for mode in ("exact", "fast"):
    eng = PairHMMEngine(0); eng.set_option("mode", mode)
    for cfg in (1, 2, 3, 4):
        bs = synth.config(cfg); cells = sum(b.num_cells for b in bs)
        eng.stage(bs)
        for _ in range(3): eng.launch()
        eng.sync()
        t = []
        for _ in range(7):
            eng.launch(); eng.sync(); st = eng.stats(); t.append((st["ms_f32"] + st["ms_fallback"], st["ms_f32"], st["ms_fallback"]))
        t.sort()
        out, nfb = eng.fetch_log10(); st = eng.stats()
        res[(mode, cfg)] = out
        print(json.dumps(dict(mode=mode, cfg=cfg, ms_step=t[0][0], ms_f32=t[0][1], ms_fallback=t[0][2], gcups_step=cells / t[0][0] * 1e-6,
                              gcups_f32=cells / t[0][1] * 1e-6, fallback=nfb, recheck=st["recheck_pairs"])), flush=True)
for cfg in (1, 2, 3, 4):
    a, b = res[("exact", cfg)], res[("fast", cfg)]
    fin = np.isfinite(a)
    print("cfg", cfg, "max |dlog10|/|log10| fast vs exact:", float(np.max(np.abs(a[fin] - b[fin]) / np.abs(a[fin]))), "identical:", int((a.view(np.uint64) == b.view(np.uint64)).sum()), "of", a.size)
