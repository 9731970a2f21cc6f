"""Batch sharding of the MSDeformAttn path across the GPUs of one box.

The op has no cross-sample dependency (the batch index only selects the value slab:
ms_deform_im2col_cuda.cuh:263,269,345), so the path shards by batch with NO data-path collective —
which is what the reference's DDP launch (segmentation/dist_train.sh:8-9) does implicitly. The only
collective here is the MAX-reduction of the per-rank device time used for reporting.
"""
import torch
import torch.distributed as dist


def batch_shard(total, world, rank):
    """Contiguous [begin, end) slice of `total` samples owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad world/rank: %d/%d' % (world, rank))
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_op_inputs(value, sampling_locations, attention_weights, world, rank, grad_output=None):
    """Slice the batch-indexed operator inputs for `rank`. spatial_shapes / level_start_index are replicated."""
    b, e = batch_shard(value.shape[0], world, rank)
    out = [value[b:e].contiguous(), sampling_locations[b:e].contiguous(), attention_weights[b:e].contiguous()]
    if grad_output is not None:
        out.append(grad_output[b:e].contiguous())
    return out


def max_over_ranks(x, device=None):
    """MAX over ranks of a python float (device time in ms); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
