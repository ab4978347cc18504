"""Frame sharding for multi-GPU batch runs: stereo frames are independent units (the reference's StereoProcessor holds
only the ORB config and the camera, stereo.rs:31-34), so rank r simply owns its own contiguous range of frames and no
collective touches the data path.  torch.distributed is used only to (a) barrier around the timed region, (b) take the
max device time over ranks and (c) sum the frames processed."""
from __future__ import annotations


def frame_range(rank: int, world: int, total_frames: int) -> range:
    """Contiguous, disjoint, complete split of `total_frames` frames (strong scaling: fixed total work)."""
    base, rem = divmod(total_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def weak_seeds(rank: int, frames_per_rank: int, unique: int = 0) -> list[int]:
    """Weak scaling: every rank processes `frames_per_rank` frames of its own; seeds are disjoint across ranks."""
    nuniq = unique or frames_per_rank
    return [rank * frames_per_rank + (i % nuniq) for i in range(frames_per_rank)]


def strong_seeds(rank: int, world: int, total_frames: int, unique: int = 0) -> list[int]:
    """Strong scaling (BASELINE configs[4]): the fixed frame set 0 .. total_frames-1 split into contiguous per-rank ranges;
    seed == global frame index (`unique` > 0 cycles that many distinct frames inside the rank's range)."""
    mine = frame_range(rank, world, total_frames)
    nuniq = unique or len(mine)
    return [mine.start + (i % nuniq) for i in range(len(mine))]


def reduce_max(x: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch

    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(frames_all_ranks: float, max_ms_over_ranks: float) -> float:
    """Whole-job frames/s: all frames of all ranks divided by the slowest rank's time."""
    return frames_all_ranks / (max_ms_over_ranks * 1e-3)
