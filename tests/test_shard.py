"""Multi-GPU host logic on CPU: regions are partitioned over ranks with no data-path collective.  Two gloo ranks
each take their share, compute a stand-in result for their regions, and the gathered shares must tile the job."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from acc_genomics_b200 import shard, synth


def test_assignment_is_a_balanced_partition():
    costs = [5, 1, 9, 3, 3, 7, 2, 8, 4, 6]
    for world in (1, 2, 3, 4, 8):
        parts = shard.assign_regions(costs, world)
        assert sorted(k for p in parts for k in p) == list(range(len(costs)))
        loads = [sum(costs[k] for k in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batches = synth.config(5, scale=0.004)            # 10 regions, same bytes on every rank (seeded)
    mine = shard.my_regions(batches, rank, world)
    # stand-in for the per-region result: one checksum per region, computed only by the owner
    res = torch.zeros(len(batches), dtype=torch.float64)
    for k in mine:
        res[k] = float(batches[k].num_cells)
    owned = torch.zeros(len(batches), dtype=torch.int64); owned[mine] = 1
    # control-plane gather only (what bench.py does with timings); the data path itself has no collective
    gathered = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(gathered, res)
    owners = [torch.zeros_like(owned) for _ in range(world)]
    dist.all_gather(owners, owned)
    if rank == 0:
        q.put((torch.stack(gathered).sum(0).numpy(), torch.stack(owners).sum(0).numpy(), [b.num_cells for b in batches]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_tile_the_job():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, owners, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (owners == 1).all()                       # every region owned by exactly one rank
    assert np.array_equal(total, np.array(want, dtype=np.float64))
