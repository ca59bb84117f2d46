"""GPU box: the double re-run (build_fallback_kernel + pmm_forward_kernel<double,...>) on configs 2, 3, 4 and a config 5
job under different task shapes.  Prints the pass time (CUDA events of the engine), its share of the measured FP64 issue
peak (12 DP instructions per cell of the re-run pairs) and checks the log10 results against the oracle."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine

eng = PairHMMEngine(0)
peak64 = eng.measure_fp64_peak()
print(json.dumps({"fp64_peak_lane_instr_per_s": peak64}))
chk = oracle.reference() or oracle.port()
shapes = [tuple(int(x) for x in s.split(",")) for s in os.environ.get("SHAPES", "64,1 8,4 4,8 4,16 2,16 1,32").split()]
for cfg in [int(c) for c in os.environ.get("CFGS", "2 3 4 5").split()]:
    regs = synth.config(cfg, scale=0.01 if cfg == 5 else 1.0)
    ref = [chk.batch(b, threads=os.cpu_count() or 1) for b in regs]
    out_r = np.concatenate([r[1].ravel() for r in ref]); fb_r = np.concatenate([r[2].ravel() for r in ref])
    # cells of the pairs that take the re-run
    fcells = 0
    for b, r in zip(regs, ref):
        fcells += int((b.read_lens[:, None].astype(np.int64) * b.hap_lens[None, :].astype(np.int64))[r[2]].sum())
    for tpw, mr in shapes:
        eng.set_option("f64_tasks_per_warp", tpw); eng.set_option("f64_max_run", mr)
        eng.stage(regs)
        for _ in range(3):
            eng.launch()
        ms = []
        for _ in range(10):
            eng.launch(); eng.sync(); ms.append(eng.stats()["ms_fallback"])
        out, nfb = eng.fetch_log10()
        ok = bool(np.array_equal(out.view(np.uint64), out_r.view(np.uint64))) and nfb == int(fb_r.sum())
        t = float(np.median(ms))
        print(json.dumps({"cfg": cfg, "tasks_per_warp": tpw, "max_run": mr, "fallback_pairs": nfb, "ms_fallback": round(t, 4),
                          "ms_f32": round(eng.stats()["ms_f32"], 4), "dp_frac_of_peak": round(fcells * 12 / (t * 1e-3) / peak64, 4),
                          "bit_equal": ok}), flush=True)
eng.close()
