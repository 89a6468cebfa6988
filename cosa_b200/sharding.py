"""Batch sharding across the GPUs of one box.

Images are independent units: rank r of G processes images [r*B/G, (r+1)*B/G) and nothing on the data path
crosses ranks.  ``torch.distributed`` (NCCL on the GPUs, gloo in the CPU tests) carries only the scalar
loss and the timing/count reductions, exactly like the reference under DDP (main.py:49, :385).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [start, stop) of ``total`` images for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(int(total), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _device_for_backend():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_reduce_sum(value):
    """Sum of a Python float / 1-element tensor over all ranks (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def all_reduce_max(value):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mean_loss_over_ranks(local_loss, local_images):
    """Image-weighted mean of the per-rank losses (each rank's loss is already divided by its local N,
    seg_helper.py:893), i.e. the loss of the un-sharded batch."""
    num = all_reduce_sum(float(local_loss) * local_images)
    den = all_reduce_sum(local_images)
    return num / den


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
