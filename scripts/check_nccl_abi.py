"""torchrun --nproc-per-node 2 scripts/check_nccl_abi.py

The C-ABI collectives (b200_nccl_*) on real multi-GPU NCCL, without torch.distributed in the data
path: rank 0 creates the unique id (handed over through the torchrun store), every rank builds its
communicator, averages a gradient bucket and sums a confusion matrix; results are checked exactly."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dasemanticsegmentationaml_b200 import _lib, build

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")            # only to pass the 128-byte id around
if rank == 0:
    build.build()
dist.barrier()
lib = _lib.lib()
uid = ctypes.create_string_buffer(128)
if rank == 0:
    assert lib.b200_nccl_unique_id(uid) == 0, _lib.last_error()
t = torch.tensor(list(uid.raw), dtype=torch.uint8)
dist.broadcast(t, 0)
uid = ctypes.create_string_buffer(bytes(t.tolist()), 128)
comm = ctypes.c_void_p()
assert lib.b200_nccl_init(uid, world, rank, ctypes.byref(comm)) == 0, _lib.last_error()
g = torch.full((1 << 20,), float(rank + 1), device="cuda")
h = torch.arange(361, dtype=torch.int64, device="cuda") * (rank + 1) * 4000000007
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
assert lib.b200_nccl_allreduce_grads(comm, ctypes.c_void_p(g.data_ptr()), ctypes.c_int64(g.numel()), 1, s) == 0
assert lib.b200_nccl_allreduce_hist(comm, ctypes.c_void_p(h.data_ptr()), 361, s) == 0
torch.cuda.synchronize()
ok = bool((g == (world + 1) / 2.0).all()) and torch.equal(
    h, torch.arange(361, dtype=torch.int64, device="cuda") * 4000000007 * (world * (world + 1) // 2))
print("rank %d: NCCL C-ABI %s" % (rank, "PASSED" if ok else "FAILED"), flush=True)
lib.b200_nccl_destroy(comm)
dist.barrier()
os._exit(0 if ok else 1)
