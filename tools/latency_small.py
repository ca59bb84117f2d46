"""GPU box: latency of single small jobs through the one-shot C ABI call (what a per-region caller such as GATK sees):
one region of config 5 (100 reads x 40 haplotypes), config 1 (128 x 32) and a 10 x 5 toy, with the engine's own breakdown."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine

eng = PairHMMEngine(0)
for kv in os.environ.get("PMM_OPTS", "").split():
    eng.set_option(*kv.split("="))
cases = [("cfg5 region 100x40", synth.config(5, scale=0.0004)[0]), ("cfg1 128x32", synth.config(1)[0])]
rng = np.random.Generator(np.random.PCG64(3))
cases.append(("toy 10x5", synth.region(rng, [151] * 10, [400] * 5)))
cases.append(("region 30x8", synth.region(rng, [151] * 30, [350, 400, 450, 500] * 2)))
cases.append(("region 100x20", synth.region(rng, [151] * 100, [350, 400, 450, 500] * 5)))
cases.append(("region 300x6", synth.region(rng, [151] * 300, [350, 400, 450, 500, 420, 380])))
for name, b in cases:
    out = np.empty(b.num_pairs)
    for _ in range(5):
        eng.stage(b); eng.launch(); eng.fetch_log10(out)
    t, parts = [], []
    for _ in range(200):
        t0 = time.perf_counter(); eng.stage(b); t1 = time.perf_counter(); eng.launch(); t2 = time.perf_counter(); eng.fetch_log10(out); t3 = time.perf_counter()
        t.append(t3 - t0); parts.append((t1 - t0, t2 - t1, t3 - t2))
    st = eng.stats()
    p = np.median(np.array(parts), axis=0) * 1e6
    print(json.dumps(dict(case=name, pairs=int(b.num_pairs), cells=int(b.num_cells), us_total=round(float(np.median(t)) * 1e6, 1),
                          us_p90=round(float(np.percentile(t, 90)) * 1e6, 1), us_stage=round(float(p[0]), 1), us_launch_call=round(float(p[1]), 1),
                          us_fetch=round(float(p[2]), 1), gpu_us_f32=round(st["ms_f32"] * 1e3, 1), gpu_us_fallback=round(st["ms_fallback"] * 1e3, 1),
                          launches=st["kernel_launches"], fallback_pairs=st["fallback_pairs"], tasks=st["f32_tasks"],
                          gcups=round(b.num_cells / float(np.median(t)) * 1e-9, 1))))
