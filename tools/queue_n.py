"""GPU box: bench.py's queue block alone (the cfg5 stream through ONE pmm_pool over N GPUs) under different wait modes.
Usage: queue_n.py N [jobs_per_gpu]; PMM_POOL_SYNC=spin|block|hybrid selects the feeders' wait."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]); jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 96
ctx = int(os.environ.get("CTX", "4"))
q, _, _ = bench.measure_queue(n, jobs, os.environ.get("TIMELINE", ""), ctx)
q["contexts"] = ctx
q["sync"] = os.environ.get("PMM_POOL_SYNC", "spin (default)")
print(json.dumps({k: q[k] for k in ("sync", "contexts", "n_gpus", "jobs", "gcups", "gcups_1gpu", "efficiency", "gpu_idle_frac", "mean_per_job")}), flush=True)
