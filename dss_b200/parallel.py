"""View-sharded data parallelism for the splatting path.

Views are independent units: the N views of a step are split contiguously over the G ranks (one process per
GPU), every rank holds the full cloud, renders and back-propagates its own slice, and ONE all-reduce(sum) of
the packed point-gradient buffer (d position | d normal | d colour, (P0, 9) fp32) makes the gradients
identical everywhere, so identical optimizer steps need no parameter broadcast.  No other collective is on
the path.  Works with any torch.distributed backend ("nccl" over NVLink on the B200 box, "gloo" in the CPU
tests); there is no collective inside the kernels because the path has no exchange step other than this one.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_views", "pack_point_grads", "unpack_point_grads", "allreduce_point_grads",
           "allreduce_visibility"]


def shard_views(n_views: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of the view batch owned by ``rank`` (sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, extra = divmod(n_views, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_point_grads(grad_points: Optional[torch.Tensor], grad_normals: Optional[torch.Tensor],
                     grad_colours: Optional[torch.Tensor], P0: int, device=None) -> torch.Tensor:
    """(P0, 9) fp32 buffer [d pos | d normal | d colour]; missing parts are zero."""
    ref = next(t for t in (grad_points, grad_normals, grad_colours) if t is not None)
    buf = torch.zeros((P0, 9), dtype=torch.float32, device=device if device is not None else ref.device)
    for i, t in enumerate((grad_points, grad_normals, grad_colours)):
        if t is not None:
            buf[:, 3 * i: 3 * i + 3] = t
    return buf


def unpack_point_grads(buf: torch.Tensor):
    return buf[:, 0:3], buf[:, 3:6], buf[:, 6:9]


def allreduce_point_grads(buf: torch.Tensor, group=None, async_op: bool = False):
    """Sum the packed gradient buffer over all ranks, in place.  Returns the work handle when async."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def allreduce_visibility(visible: torch.Tensor, group=None):
    """any-over-views visibility (point_modeling.py:172-173) across ranks: one tiny all-reduce(max)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return visible
    v = visible.to(torch.uint8)
    dist.all_reduce(v, op=dist.ReduceOp.MAX, group=group)
    return v.to(visible.dtype)
