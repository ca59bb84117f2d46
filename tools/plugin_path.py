"""GPU box: config 2 through the client path (pairhmm_worker_forward = PairHMMClient + PairHMMWorker over libPairHMMTask.so)
for different tile counts, next to the serial C-ABI path; one traced call per tile count on stderr."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    tiles = sys.argv[1]
    os.environ["PAIRHMM_WORKER_TILES"] = tiles
    os.environ.setdefault("PAIRHMM_DEVICES", "0"); os.environ.setdefault("PAIRHMM_SLOTS", "4")
    from acc_genomics_b200 import hostlayer, synth
    b = synth.config(int(os.environ.get("CFG", "2")))[0]
    job = hostlayer.WorkerJob(b)
    for _ in range(5):
        hostlayer.worker_forward(job)
    n = 40
    t0 = time.perf_counter()
    for _ in range(n):
        hostlayer.worker_forward(job)
    dt = (time.perf_counter() - t0) / n
    print(json.dumps({"tiles": int(tiles), "ms_per_batch": round(dt * 1e3, 4), "gcups": round(b.num_cells / dt * 1e-9, 1)}), flush=True)
    if os.environ.get("TRACE_ONE"):
        os.environ["PAIRHMM_TRACE"] = "1"
    hostlayer.shutdown()
else:
    for t in os.environ.get("TILES", "1 2 3 4 6").split():
        subprocess.run([sys.executable, __file__, t], check=False)
