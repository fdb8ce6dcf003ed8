"""A raw NCCL communicator for callers of the C ABI that do not go through torch's process group
(include/mds_rdf.h: mds_rdf_allreduce takes an ``ncclComm_t``).  ctypes over libnccl.so.2 — the
copy torch ships if torch is importable, else whatever the loader finds.  The unique id is created
by rank 0 and handed to the other ranks by the caller's own channel (``exchange``): under
torch.distributed that is ``broadcast_object_list``; an MPI program would use MPI_Bcast."""
from __future__ import annotations

import ctypes
import os

NCCL_UNIQUE_ID_BYTES = 128


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * NCCL_UNIQUE_ID_BYTES)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    candidates = []
    try:
        import nvidia.nccl  # the wheel torch depends on
        candidates.append(os.path.join(os.path.dirname(nvidia.nccl.__path__[0] + "/"), "lib",
                                       "libnccl.so.2"))
    except Exception:  # noqa: BLE001
        pass
    candidates.append("libnccl.so.2")
    last = None
    for path in candidates:
        try:
            lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
            break
        except OSError as exc:
            last = exc
    else:
        raise RuntimeError(f"libnccl.so.2 not found: {last}")
    lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
    lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId,
                                     ctypes.c_int]
    lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    lib.ncclGetErrorString.restype = ctypes.c_char_p
    lib.ncclGetErrorString.argtypes = [ctypes.c_int]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what}: {load().ncclGetErrorString(rc).decode()}")


def unique_id() -> bytes:
    uid = _UniqueId()
    _check(load().ncclGetUniqueId(ctypes.byref(uid)), "ncclGetUniqueId")
    return ctypes.string_at(ctypes.byref(uid), NCCL_UNIQUE_ID_BYTES)  # (c_char arrays stop at NUL)


class Communicator:
    """``ncclCommInitRank`` on the current CUDA device.  ``handle`` is the ``ncclComm_t``."""

    def __init__(self, uid: bytes, world_size: int, rank: int):
        u = _UniqueId()
        ctypes.memmove(ctypes.byref(u), uid, NCCL_UNIQUE_ID_BYTES)
        comm = ctypes.c_void_p()
        _check(load().ncclCommInitRank(ctypes.byref(comm), world_size, u, rank), "ncclCommInitRank")
        self.handle = comm.value
        self.world_size, self.rank = world_size, rank

    @classmethod
    def from_torch(cls, device=None, group=None) -> "Communicator":
        """A communicator over the ranks of an initialised torch.distributed group: rank 0 makes
        the unique id, ``broadcast_object_list`` carries it."""
        import torch
        import torch.distributed as dist
        if device is not None:
            torch.cuda.set_device(device)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group else 0,
                                   group=group)
        return cls(box[0], world, rank)

    def close(self):
        if self.handle:
            load().ncclCommDestroy(ctypes.c_void_p(self.handle))
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
