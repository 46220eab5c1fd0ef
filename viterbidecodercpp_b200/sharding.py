"""Multi-GPU plumbing: one process per GPU, frames sharded in contiguous ranges, no data-path collective (frames are independent:
SURVEY.md section 8e).  torch.distributed is only used for the barrier and for max-over-ranks timing."""


def frame_range(rank, world, n_frames):
    """contiguous range [f0, f1) of frames owned by `rank` (same split as vitb_decode_batch_multi in csrc/vitb_api.cu)"""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
