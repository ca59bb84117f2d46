"""GPU box: back-to-back pmm_launch of one staged job -- the sum of the per-launch event brackets (what bench.py's `value`
uses) against the wall time of the whole sequence, with and without an L2 flush between launches."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine
torch.cuda.set_device(0)
eng = PairHMMEngine(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side); eng.set_option("stream", side.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cfg in [int(x) for x in os.environ.get("CFGS", "2 3").split()]:
    b = synth.config(cfg); cells = sum(x.num_cells for x in b)
    eng.stage(b)
    for _ in range(5): eng.launch()
    eng.sync()
    for do_flush in (True, False):
        n = 30
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        torch.cuda.synchronize(); eng.sync()
        t0 = time.perf_counter()
        for a, bb in ev:
            if do_flush: flush.fill_(1)
            a.record(); eng.launch(); bb.record()
        eng.sync(); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        br = [a.elapsed_time(bb) for a, bb in ev]
        span = ev[0][0].elapsed_time(ev[-1][1])
        st = eng.stats()
        print(json.dumps({"cfg": cfg, "flush": do_flush, "sum_brackets_ms_per_step": round(sum(br) / n, 4), "first_a_to_last_b_ms_per_step": round(span / n, 4),
                          "wall_ms_per_step": round(wall * 1e3 / n, 4), "brackets_first5": [round(x, 3) for x in br[:5]], "brackets_last3": [round(x, 3) for x in br[-3:]],
                          "gcups_by_brackets": round(cells * n / sum(br) * 1e-6), "gcups_by_wall": round(cells * n / wall * 1e-9),
                          "ms_f32_last": round(st["ms_f32"], 3), "ms_fallback_last": round(st["ms_fallback"], 3)}))
eng.close()
