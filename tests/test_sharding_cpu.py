"""world_size-2 gloo test of the scene sharding + single metrics all-reduce (host logic, CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uni3detr_b200.sharding import reduce_metrics, scene_indices
    mine = scene_indices(total, rank, world)
    checksum = float(sum(i * i for i in mine))
    n, t, c = reduce_metrics(len(mine), 1.0 + rank, checksum)
    q.put((rank, mine, n, t, c))
    dist.destroy_process_group()


def test_scene_sharding_and_metrics_allreduce():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seen = sorted(i for _, mine, *_ in res for i in mine)
    assert seen == list(range(total))                      # every scene exactly once
    for rank, mine, n, t, c in res:
        assert all(i % world == rank for i in mine)
        assert n == total and t == 2.0 and c == float(sum(i * i for i in range(total)))


def test_single_process_is_identity():
    from uni3detr_b200.sharding import reduce_metrics, scene_indices
    assert scene_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert reduce_metrics(5, 0.25, 7.0) == (5.0, 0.25, 7.0)
