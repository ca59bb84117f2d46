"""Developer smoke check on a GPU box: CUDA path vs oracle on small slices of configs 1-4, and a first timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from acc_genomics_b200 import synth
from acc_genomics_b200.engine import PairHMMEngine

eng = PairHMMEngine(0)
ref = oracle.reference() or oracle.port()
print("checker:", ref.kind)
ok = True
for cfg, scale in ((1, 0.25), (2, 0.03), (3, 0.03), (4, 0.03)):
    b = synth.config(cfg, scale=scale)[0]
    raw_r, out_r, fb_r = ref.batch(b, threads=8)
    raw, out, mask = eng.forward(b)
    eq_raw = np.array_equal(raw.view(np.uint32), raw_r.view(np.uint32))
    eq_mask = np.array_equal(mask, fb_r)
    eq_out = np.array_equal(out.view(np.uint64), out_r.view(np.uint64))
    nbad = int((raw.view(np.uint32) != raw_r.view(np.uint32)).sum())
    print(f"cfg{cfg} reads={b.num_read} haps={b.num_hap} raw_bit_equal={eq_raw} (bad {nbad}/{raw.size}) mask_equal={eq_mask} "
          f"log10_bit_equal={eq_out} fb={fb_r.mean():.3f} stats={eng.stats()}")
    if not eq_raw:
        bad = np.argwhere(raw.view(np.uint32) != raw_r.view(np.uint32))[:5]
        for i, j in bad: print("   ", i, j, raw[i, j], raw_r[i, j])
    if not eq_out and eq_raw:
        bad = np.argwhere(out.view(np.uint64) != out_r.view(np.uint64))[:5]
        for i, j in bad: print("   out", i, j, out[i, j], out_r[i, j], mask[i, j])
    ok &= eq_raw and eq_mask and eq_out

# timing on full config 2
b = synth.config(2)[0]
eng.stage([b])
for _ in range(3): eng.launch()
eng.sync()
t = []
for _ in range(5):
    eng.launch(); eng.sync(); t.append(eng.stats()["ms_f32"])
st = eng.stats()
print("cfg2 full: cells=%d ms_f32=%s ms_fallback=%.3f GCUPS(f32 pass)=%.1f tasks=%d" % (b.num_cells, [round(x, 3) for x in t], st["ms_fallback"], b.num_cells / (min(t) * 1e-3) * 1e-9, st["f32_tasks"]))
print("fp32 peak:", eng.measure_fp32_peak())
print("ALL OK" if ok else "MISMATCH")
