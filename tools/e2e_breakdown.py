"""GPU box: where the serial end-to-end time of one cfg2 job goes (host stage / kernels / fetch), C ABI and C++ layer."""
import sys, os, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from acc_genomics_b200 import synth, fixtures
from acc_genomics_b200.engine import PairHMMEngine
import acc_genomics_b200.batch as B
eng = PairHMMEngine(0)
b = synth.config(2)
eng.stage(b); eng.launch(); eng.fetch_log10()
res = np.empty(b[0].num_pairs)
T = {"stage": [], "launch": [], "fetch": []}
for _ in range(30):
    t0 = time.perf_counter(); eng.restage(); t1 = time.perf_counter(); eng.launch(); t2 = time.perf_counter(); eng.fetch_log10(res); t3 = time.perf_counter()
    T["stage"].append(t1 - t0); T["launch"].append(t2 - t1); T["fetch"].append(t3 - t2)
st = eng.stats()
print({k: round(float(np.median(v)) * 1e3, 3) for k, v in T.items()}, "ms; engine stats ms_stage %.3f ms_f32 %.3f ms_fallback %.3f ms_fetch %.3f" % (st["ms_stage"], st["ms_f32"], st["ms_fallback"], st["ms_fetch"]))
rs, hs = B.serialize_reads(b[0]), B.serialize_haps(b[0])
t = []
for _ in range(20):
    t0 = time.perf_counter(); eng.forward_log10_serialized(rs, hs, b[0].num_pairs); t.append(time.perf_counter() - t0)
print("pmm_forward_log10_serialized: %.3f ms -> %.0f GCUPS" % (np.median(t) * 1e3, b[0].num_cells / np.median(t) * 1e-9))
# C++ layer: test bench on a cfg2-sized fixture, direct and client mode
d = tempfile.mkdtemp()
out, _ = eng.fetch_log10()
fixtures.write_folder(d, [b[0]] * 6, [out] * 6)
for mode in ([], ["--client"]):
    p = subprocess.run([os.path.join("pairhmm", "bin", "pairhmm_host_tb")] + mode + ["-", d], capture_output=True, text=True)
    print(" ".join(mode) or "direct", "|", " | ".join(l for l in p.stdout.splitlines() if l[:1].isdigit() or "GCUPS" in l or "failed" in l or "bit-id" in l))
