"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: rendezvous on 127.0.0.1, barrier, max-over-ranks
timing reduction, per-rank batch split and the flat-bucket gradient mean all-reduce used by the training step."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from ayolov2_b200 import dist_utils as du

    assert du.init("gloo") == world and du.world_size() == world
    du.barrier()
    mx = du.max_over_ranks(10.0 + rank)
    grads = [torch.full((7,), float(rank + 1)), torch.full((3, 5), 2.0 * (rank + 1)), torch.ones(4, dtype=torch.float64) * rank]
    nb = du.allreduce_mean_(grads, bucket_bytes=64)
    q.put((rank, mx, [g.clone() for g in grads], nb, du.shard_batch(128, rank, world)))
    du.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_gloo_world2_helpers():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, mx, grads, nb, shard in res:
        assert mx == 11.0  # max over ranks of 10 + rank
        assert shard == 64
        assert nb >= 2  # small bucket size forces several buckets
        assert torch.allclose(grads[0], torch.full((7,), 1.5))
        assert torch.allclose(grads[1], torch.full((3, 5), 3.0))
        assert torch.allclose(grads[2], torch.full((4,), 0.5, dtype=torch.float64))
