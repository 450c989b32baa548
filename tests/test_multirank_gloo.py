"""world_size-2 gloo check of the N>1 path's host logic: every rank derives the same
longest-first partition without communicating, the shards are disjoint and cover the job, and
the max-over-ranks reduction bench.py uses behaves."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from viloca_b200 import synthetic
    # strong scaling: the windows of ONE genome over the ranks, longest first; the cost uses the
    # committed iteration table (amplicon: the counts vary 10..17178 between windows)
    spec = synthetic.WORKLOADS["sars_amplicon"]
    parts, how = bench.shard_windows(spec, spec.n_windows, world)
    costs, _ = bench.predicted_costs(spec, spec.n_windows)
    mine = [(w, costs[w]) for w in parts[rank]]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = torch.tensor([float(rank + 1), float(len(mine))], dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((gathered, tmax.tolist(), tsum.tolist(), spec.n_windows, how))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_is_consistent():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax, tsum, n_items, how = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    flat = [x[0] for shard in gathered for x in shard]
    assert len(flat) == n_items == 98 and len(set(flat)) == n_items        # disjoint cover of the genome
    assert how.startswith("profile-guided")
    loads = [sum(c for _, c in shard) for shard in gathered]
    # longest-first: no rank carries more than the mean load plus the largest single window
    biggest = max(c for shard in gathered for _, c in shard)
    assert max(loads) <= sum(loads) / world + biggest
    assert tmax[0] == 2.0 and tsum[1] == 98.0
