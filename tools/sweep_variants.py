"""Tuning sweep on a GPU box: time the float pass of configs 1, 2 and 4 with every (K rows per lane, W lanes per
read) variant that fits the read length, and check each against the planner's default choice bit for bit.
Writes gpurun_out/sweep_variants.jsonl.  Feeds the cost model of csrc/pmm_plan.cpp (step_cost)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine, PmmError

VARIANTS = [(k, 8) for k in range(4, 21)] + [(k, 16) for k in (4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16)] + \
           [(k, 32) for k in (4, 5, 6, 7, 8, 9, 10, 12, 14, 16)]
eng = PairHMMEngine(0)
for kv in os.environ.get("PMM_OPTS", "").split():
    eng.set_option(*kv.split("="))
os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/sweep_variants.jsonl", "a")
cfgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [2, 1, 4]
tpw = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6]
for cfg in cfgs:
    bs = synth.config(cfg)
    R = int(max(b.read_lens.max() for b in bs))
    cells = sum(b.num_cells for b in bs)
    eng.set_option("force_variant", "0,0")
    eng.stage(bs); eng.launch()
    ref = eng.fetch_raw().view(np.uint32)
    for (K, W) in [(0, 0)] + [v for v in VARIANTS if v[0] * v[1] >= R + 1 and v[0] * v[1] < 2 * (R + 1)]:
        for t in tpw:
            try:
                eng.set_option("force_variant", f"{K},{W}")
                eng.set_option("tasks_per_warp", t)
            except PmmError as e:
                print("skip", K, W, e); continue
            eng.stage(bs)
            for _ in range(3):
                eng.launch()
            eng.sync()
            ms = []
            for _ in range(7):
                eng.launch(); eng.sync(); ms.append(eng.stats()["ms_f32"])
            raw = eng.fetch_raw().view(np.uint32)
            ok = bool(np.array_equal(raw, ref))
            rec = dict(cfg=cfg, K=K, W=W, tasks_per_warp=t, ms_f32=min(ms), gcups=cells / (min(ms) * 1e-3) * 1e-9,
                       tasks=eng.stats()["f32_tasks"], bit_equal_to_default=ok)
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n"); out.flush()
print("fp32 peak", eng.measure_fp32_peak())
