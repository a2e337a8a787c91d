"""Data-parallel plumbing of the hot path: keyframes are sharded over the ranks of one
box, every rank samples rays from its own shard, and ONE all-reduce of the flat gradient per step is
the only collective (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests).

Replicas stay bit-identical because every rank applies the same averaged gradient with the same
AdamW state; rank 0's initial parameters are broadcast once."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def keyframe_owner(frame_index, world_size):
    """Round-robin by insertion order: keyframe k lives on rank k mod world."""
    return int(frame_index) % int(world_size)


def shard_keyframes(frame_indices, rank, world_size):
    return [f for f in frame_indices if keyframe_owner(f, world_size) == rank]


def allreduce_sum_(flat, group=None):
    """In-place sum over ranks of a flat buffer (the packed gradient).  No-op for a single rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def average_gradients_(flat, group=None):
    """sum over ranks / world: each rank's gradient is the mean over ITS samples, so the result is the
    mean over all samples when the per-rank valid-sample counts agree (they do for all-valid depth;
    otherwise it is the mean of per-rank means, the usual data-parallel convention)."""
    n = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    allreduce_sum_(flat, group)
    if n > 1:
        flat.mul_(1.0 / n)
    return flat


def broadcast_parameters_(flat, src=0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src, group=group)
    return flat
