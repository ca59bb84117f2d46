"""The N > 1 path of bench.py on CPU: one process per GPU, no data-path collective -- the ranks only agree on the config,
meet at barriers and take the MAX of their timings.  Two gloo ranks run that control plane here; the data path itself
(pmm_pool over several GPUs) is one process and is covered by tests/test_gpu_pool.py on the GPU box."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from acc_genomics_b200 import synth
    # every rank builds the same batch (same seed): weak scaling of independent batches, nothing to exchange
    b = bench.workload(1, 1, 0.25)
    cfg = bench.config_dict(1, 0.25, b)
    # this rank's timings: rank 1 is the slow one on the first entry, rank 0 on the second
    mine = [1.0 + rank, 5.0 - rank, 2.5]
    worst = bench.max_over_ranks(mine, world, "cpu")
    cpu_group = dist.new_group(backend="gloo")
    dist.barrier(group=cpu_group)                       # what keeps ranks 1.. idle while rank 0 drives the queue block
    q.put((rank, worst, cfg["cells_per_step_per_gpu"], cfg["workload"], int(b[0].rs[:64].sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_on_config_and_take_the_slowest_timing():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert got[0][1] == got[1][1] == [2.0, 5.0, 2.5]              # MAX over ranks, identical on both
    assert got[0][2:] == got[1][2:]                               # same cells, same workload name, same bytes


def test_reference_arm_and_our_arm_describe_the_same_config():
    """--impl reference prints the same `config` object as the GPU arm would (the driver's same_config check), and only rank 0
    prints at all."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    import bench
    assert line["impl"] == "reference" and line["config"] == bench.config_dict(1, 1.0, bench.workload(1, 1, 1.0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port") and line["value"] > 0
    env["RANK"] = "1"; env["WORLD_SIZE"] = "2"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0 and p.stdout.strip() == ""
