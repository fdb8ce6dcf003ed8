"""
Multi-GPU plumbing: configurations are the shard unit.

The reference already sums per-batch histograms (calculators/radial_distribution_function.py:
884-885) and splits the sampled configurations into batches with ``np.array_split`` (:594); here
the same primitive assigns rank g block g, and ONE collective — an allreduce(sum) of the
[n_pairs][nbins] 64-bit accumulator, <= 120 KB — combines the partial histograms over
NVLink/NVSwitch.  There is no other exchange on this path.
"""
from __future__ import annotations

import numpy as np


def shard_configurations(sample_configurations, world_size: int, rank: int) -> np.ndarray:
    """Block ``rank`` of ``np.array_split(sample_configurations, world_size)`` (may be empty)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    return np.array_split(np.asarray(sample_configurations), world_size)[rank]


def counts_as_tensor(handle):
    """The handle's device accumulator viewed as an int64 torch tensor (no copy).  uint64 and
    int64 sums are the same bits; NCCL has no uint64 reduction in torch."""
    import torch
    ptr, _stream = handle.device_counts()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (handle.n_pairs, handle.nbins), "typestr": "<i8",
                                    "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_Wrap(), device=f"cuda:{handle.device}")


_RAW_COMMS = {}  # device -> mdsuite_b200._cuda.nccl.Communicator over the default group


def raw_communicator(device: int):
    """The process's raw NCCL communicator for ``device`` over the ranks of the default
    torch.distributed group (created on first use; every rank must ask for it at the same point)."""
    from mdsuite_b200._cuda import nccl
    comm = _RAW_COMMS.get(device)
    if comm is None:
        comm = nccl.Communicator.from_torch(device=device)
        _RAW_COMMS[device] = comm
    return comm


def close_raw_communicators():
    for comm in _RAW_COMMS.values():
        comm.close()
    _RAW_COMMS.clear()


def allreduce_counts(handle, group=None, raw: bool = True):
    """In-place sum of the per-GPU histograms over the process group — ONE ncclAllReduce.

    raw=True (default): through the C ABI, ``mds_rdf_allreduce(handle, ncclComm_t)``, on the
    library's own stream; ordering with the kernels is the library's business.  raw=False: torch's
    process group reduces a tensor view of the accumulator, bracketed by join / wait."""
    import torch
    import torch.distributed as dist
    if raw and group is None and hasattr(handle, "allreduce"):
        with torch.cuda.device(handle.device):
            handle.allreduce(raw_communicator(handle.device))
        return None
    t = counts_as_tensor(handle)
    with torch.cuda.device(handle.device):
        handle.join()  # torch's stream waits for the kernels
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        handle.wait()  # the library's stream waits for the collective
    return t


def allreduce_host_counts(counts: np.ndarray, group=None) -> np.ndarray:
    """Same reduction for host-resident histograms (gloo); used by CPU tests of the sharding
    logic and as the path for process groups without a CUDA backend."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(counts).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy().view(np.uint64)


def distributed_context():
    """(rank, world, dist module or None) under an initialised torch.distributed group of more than
    one rank, else (0, 1, None)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


def allreduce_host_sums(sums: np.ndarray, device=None, group=None) -> np.ndarray:
    """Sum of a host float64 array over the process group — the ADF's one collective: the weight
    sums [batch][combination][bin] of the configurations every rank took.  Goes through the GPU
    when the group's backend is NCCL."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(sums, dtype=np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy()
