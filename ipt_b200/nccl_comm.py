"""A raw ncclComm_t for `ipt_plane_allreduce` (harness glue, one process per GPU).

The C ABI takes the communicator of the host application (include/ipt_b200.h). A Python host has none to hand over —
torch.distributed keeps its communicators private — so this creates one with the NCCL already loaded in the process:
rank 0 draws the ncclUniqueId, `torch.distributed` (any backend) ships it to the other ranks, every rank calls
ncclCommInitRank. The data path then runs through the product's own collective, not through torch's.
"""
from __future__ import annotations

import ctypes as C


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class NcclComm:
    def __init__(self, dist, rank: int, world: int):
        self.nccl = C.CDLL("libnccl.so.2")  # the copy torch already loaded
        uid = _UniqueId()
        if rank == 0:
            rc = self.nccl.ncclGetUniqueId(C.byref(uid))
            if rc:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        raw = [bytes(uid)] if rank == 0 else [None]
        dist.broadcast_object_list(raw, src=0)
        C.memmove(C.byref(uid), raw[0], 128)
        self.comm = C.c_void_p()
        self.nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        rc = self.nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        if rc:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")

    def close(self):
        if self.comm:
            self.nccl.ncclCommDestroy(self.comm)
            self.comm = C.c_void_p()
