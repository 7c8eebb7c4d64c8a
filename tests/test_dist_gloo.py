"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py (stream partition, max-over-ranks timing,
summed frames).  The data path itself has no collective (streams are independent)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rd_vio_b200.parallel import aggregate_throughput, partition_streams
    ids = partition_streams(rank, world, 4)
    ms = 10.0 if rank == 0 else 20.0               # rank 1 is the slow one
    fps, ms_max = aggregate_throughput(len(ids) * 5, ms, dist)
    out[rank] = (ids, fps, ms_max)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_aggregation():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_run, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == [0, 1, 2, 3] and out[1][0] == [4, 5, 6, 7]
    for r in range(world):
        assert out[r][2] == 20.0                                  # max over ranks
        assert abs(out[r][1] - (2 * 4 * 5) / 20e-3) < 1e-6        # whole-job frames / slowest rank's time


def test_round_robin_layout():
    from rd_vio_b200.parallel import partition_round_robin
    parts = partition_round_robin(list(range(10)), 4)
    assert parts[0] == [0, 4, 8] and parts[3] == [3, 7] and sorted(sum(parts, [])) == list(range(10))
