"""GPU box: config 2 through the client path (pairhmm_worker_forward) from several caller threads, one client each, under
different tile policies: the worker's own choice (one tile per batch while other workers run, else 3-6), and forced tile
counts (PAIRHMM_WORKER_TILES).  One subprocess per setting."""
import json, os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 2:
    tiles, threads = sys.argv[1], int(sys.argv[2])
    if tiles != "auto":
        os.environ["PAIRHMM_WORKER_TILES"] = tiles
    os.environ.setdefault("PAIRHMM_DEVICES", "0"); os.environ.setdefault("PAIRHMM_SLOTS", "4")
    from acc_genomics_b200 import hostlayer, synth
    b = synth.config(int(os.environ.get("CFG", "2")))[0]
    jobs = [hostlayer.WorkerJob(b) for _ in range(threads)]
    per = int(os.environ.get("PER", "24"))
    go = threading.Barrier(threads + 1)

    def work(j):
        for _ in range(3):
            hostlayer.worker_forward(j)
        go.wait()
        for _ in range(per):
            hostlayer.worker_forward(j)
    th = [threading.Thread(target=work, args=(j,)) for j in jobs]
    [t.start() for t in th]
    go.wait()
    t0 = time.perf_counter()
    [t.join() for t in th]
    dt = (time.perf_counter() - t0) / (per * threads)
    print(json.dumps({"tiles": tiles, "threads": threads, "slots": os.environ["PAIRHMM_SLOTS"], "ms_per_batch": round(dt * 1e3, 4),
                      "gcups": round(b.num_cells / dt * 1e-9, 1)}), flush=True)
    hostlayer.shutdown()
else:
    for threads in (1, 2, 3, 4):
        for t in os.environ.get("TILES", "auto 1 2 4").split():
            subprocess.run([sys.executable, __file__, t, str(threads)], check=False)
