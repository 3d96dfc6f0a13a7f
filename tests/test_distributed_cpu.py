"""N>1 host logic on CPU: world_size-2 gloo processes exercise the seed sharding and the statistics all-reduce
that bench.py / trainers use across GPUs (the data path itself has no collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mocca_envs_b200.distributed import allreduce_stats, global_env_ids, max_over_ranks, shard_seed

    n = 8
    stats = {"episodes": 10 + rank, "return_sum": -5.0 * (rank + 1), "length_sum": 100.0 + rank, "nonfinite": 0,
             "overflow": rank}
    out = allreduce_stats(stats)
    tmax = max_over_ranks(1.0 + rank)
    q.put((rank, shard_seed(1234, rank, n), list(global_env_ids(rank, n)), out, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_stats_and_sharding():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, ids0, out0, t0), (r1, s1, ids1, out1, t1) = res
    assert (s0, s1) == (1234, 1242)
    assert ids0 + ids1 == list(range(16))  # shards tile the global env index space
    assert out0 == out1
    assert out0["episodes"] == 21 and out0["return_sum"] == -15.0 and out0["length_sum"] == 201.0
    assert out0["overflow"] == 1 and abs(out0["mean_return"] - (-15.0 / 21)) < 1e-12
    assert t0 == t1 == 2.0


def test_shard_seeds_match_single_process_seeding():
    """Seeds of a 2-rank run are the same set as a 1-rank run with twice the envs."""
    from mocca_envs_b200.distributed import shard_seed
    from mocca_envs_b200.seeding import mt_state_rows

    base, n = 77, 4
    single = mt_state_rows([base + i for i in range(2 * n)])
    sharded = np.concatenate([mt_state_rows([shard_seed(base, r, n) + i for i in range(n)]) for r in range(2)])
    assert np.array_equal(single, sharded)
