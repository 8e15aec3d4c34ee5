"""2+ GPU check of ipt_plane_allreduce (the C-ABI collective) with a raw ncclComm_t, without torch's collectives on
the data path: `torchrun --nproc-per-node 2 tools/multi_gpu_nccl_abi.py`. torch.distributed (gloo) only ships the
ncclUniqueId to the other ranks. Each rank renders its pass range; after the all-reduce every rank must hold the
single-GPU render of all passes."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import torch.distributed as dist

from ipt_b200 import capi, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
torch.cuda.set_device(local)
nccl = C.CDLL("libnccl.so.2")  # the copy torch already loaded


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
raw = [bytes(uid)] if rank == 0 else [None]
dist.broadcast_object_list(raw, src=0)
C.memmove(C.byref(uid), raw[0], 128)
comm = C.c_void_p()
nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

W = H = 96
total = 6
sd = capi.SceneDescription("cornell")
sc = capi.Scene(sd, local)
plane = capi.Plane(sc, W, H)
begin, count = sharding.shard_passes(total, world, rank)
plane.render(capi.default_params(width=W, height=H, pass_begin=begin, pass_count=count, seed=5))
plane.allreduce(comm)
s, q, c = plane.download()
one_s, one_q, one_c, _ = sc.render_host(capi.default_params(width=W, height=H, pass_begin=0, pass_count=total, seed=5))
ok = np.array_equal(c, one_c) and np.allclose(s, one_s, rtol=1e-5, atol=1e-6) and np.allclose(q, one_q, rtol=1e-5, atol=1e-6)
print(f"rank {rank}/{world}: passes [{begin},{begin + count}) allreduce == single-GPU render: {ok}", flush=True)
nccl.ncclCommDestroy(comm)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
