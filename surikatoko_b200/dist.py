"""torch.distributed plumbing for multi-GPU runs: one process per GPU, NCCL over NVLink.

The engine asks the host for an in-place sum all-reduce of `count` doubles at a device pointer, ordered on its CUDA stream
(srk_ba_set_allreduce).  Here that is torch.distributed.all_reduce on a tensor view of the engine's own buffer, issued with the
engine's stream as torch's current stream, so NCCL's stream waits for the producing kernels and the consumers wait for NCCL.
"""
import torch
import torch.distributed as dist


class _DevView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def attach_allreduce(engine, stream, device, group=None):
    """Registers the all-reduce callback on `engine` (which must already run on `stream`)."""
    views = {}
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce(ptr, count, _stream):
        key = (ptr, count)
        t = views.get(key)
        if t is None:
            t = torch.as_tensor(_DevView(ptr, count), device=device)
            views[key] = t
        with torch.cuda.stream(stream):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    engine.set_allreduce(allreduce, rank, world)
    return views


def attach_nccl(engine, group=None):
    """The library's own NCCL exchange (srk_ba_nccl_init, C++ host path, no Python callback per all-reduce): torch.distributed only
    carries the 128-byte ncclUniqueId from rank 0 to the other ranks once."""
    from .capi import nccl_unique_id
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    engine.nccl_init(box[0], rank, world)
