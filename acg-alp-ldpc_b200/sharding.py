"""Frame sharding over ranks and the path's only collective.

Frames are independent: rank r of W owns the GLOBAL frame indices
[begin, end) given by shard_range(); the Philox channel is addressed by global
index, so the union of the shards is the same set of frames for every W.  The
counter blocks (exp()'s correct / pseudo / Hamming sums, experiment.h:70-78) are
summed over ranks with one all-reduce at the end -- NCCL over NVLink on the GPU
box, gloo in the CPU tests.
"""
import numpy as np


def shard_range(total_frames, rank, world):
    """Contiguous, balanced split of [0, total_frames): sizes differ by at most one."""
    base, extra = divmod(int(total_frames), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_counters(counters, device=None):
    """Sum a dict (or array) of integer counters over all ranks of the default process group."""
    import torch
    import torch.distributed as dist
    keys = None
    if isinstance(counters, dict):
        keys = sorted(counters)
        vec = np.array([counters[k] for k in keys], dtype=np.int64)
    else:
        vec = np.asarray(counters, dtype=np.int64)
    t = torch.from_numpy(vec.copy())
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy()
    return dict(zip(keys, (int(x) for x in out))) if keys is not None else out
