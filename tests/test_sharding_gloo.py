"""CPU, world_size 2 over gloo: the N > 1 path (frame sharding + stats gather) without GPUs.
Each rank runs the oracle on its shard; the union of shards must equal the single-process result."""
import hashlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frame_hash(idx):
    import bevgen_b200 as bg
    from oracle import c_oracle as co
    f = bg.synthetic.frame_c2(idx, n=300, S=2)
    bev = co.lidar_bev_fused(f.points4, f.sweep_offsets, f.rel_tf, 2)
    return int.from_bytes(hashlib.sha256(bev.tobytes()).digest()[:6], "little")


def _worker(rank, world, port, nframes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bevgen_b200 import sharding
    mine = list(sharding.shard_range(nframes, rank, world))
    hashes = {f"h{i}": float(_frame_hash(i)) for i in mine}
    # every rank contributes the same keys: pad with -1 for frames it does not own
    stats = {f"h{i}": hashes.get(f"h{i}", -1.0) for i in range(nframes)}
    stats["ms"] = 10.0 + rank
    g = sharding.gather_stats(stats)
    mx = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((g, mx))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_cover_the_batch():
    world, nframes = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nframes, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, mx = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == 11.0 and g["ms"] == [10.0, 11.0]
    for i in range(nframes):
        owners = [v for v in g[f"h{i}"] if v >= 0]
        assert len(owners) == 1, "every frame is owned by exactly one rank"
        assert owners[0] == float(_frame_hash(i))
