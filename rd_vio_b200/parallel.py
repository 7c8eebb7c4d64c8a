"""Multi-GPU layout of the front-end: camera streams are independent, so ranks own
disjoint blocks of streams and the only cross-rank traffic is the reduction of timing scalars
(no data-path collective, no NCCL on the pixel path)."""
from __future__ import annotations


def partition_streams(rank: int, world: int, streams_per_gpu: int):
    """Weak-scaling layout used by bench.py: rank r owns streams [r*S, (r+1)*S)."""
    if not (0 <= rank < world) or streams_per_gpu < 1:
        raise ValueError("bad rank/world/streams_per_gpu")
    return list(range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu))


def partition_round_robin(stream_ids, world: int):
    """Stream s -> GPU s mod world (deployment layout: a stream's pyramids never move)."""
    return [[s for s in stream_ids if s % world == r] for r in range(world)]


def aggregate_throughput(frames_local: int, ms_local: float, dist=None, device=None):
    """Whole-job frames/s: sum of frames over ranks / MAX of the per-rank device times."""
    if dist is None or not dist.is_initialized():
        return frames_local / (ms_local * 1e-3), ms_local
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    f = torch.tensor([float(frames_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(f.item()) / (float(t.item()) * 1e-3), float(t.item())
