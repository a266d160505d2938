"""Multi-GPU plumbing: one process per GPU, frames sharded by rank, no data-path collective.
The only exchange is the all-reduce of the two simulation counters (bit errors, block errors) -- NCCL over
NVLink on GPUs, gloo in the CPU tests -- and the max-over-ranks of device timings."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend, device=None):
    """Idempotent init from the torchrun environment; a single process needs no group."""
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_range(total_frames, rank, world):
    """Contiguous split of a (B,N) batch: the first B % world ranks take one extra frame."""
    base, rem = divmod(int(total_frames), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_counters(counters):
    """counters: int64 tensor [bit_errors, block_errors(, frames)] on the rank's device; summed in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
