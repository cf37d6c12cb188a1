"""torchrun --nproc-per-node 2 scripts/check_global_ce.py

losses.GLOBAL_BATCH_MEAN on real NCCL: with a different number of labelled pixels on every rank, the fused
cross-entropy must return the mean over the GLOBAL batch on all ranks (what nn.DataParallel + one
CrossEntropyLoss on the gathered logits computes, reference train.py:145-152,214-217,497), and a rank's
gradient divided by the world size must equal that rank's slice of the gradient of the single-process loss
on the concatenated batch."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dasemanticsegmentationaml_b200 import build, losses

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
if rank == 0:
    build.build()
dist.barrier()
dev = torch.device("cuda", local)
n, h, w, H, W = 2, 16, 32, 128, 256
g = torch.Generator().manual_seed(5 + rank)
lr = torch.zeros(n, h, w, 32)
lr[..., :19] = 3 * torch.randn(n, h, w, 19, generator=g)
lab = torch.randint(0, 19, (n, H, W), generator=g)
lab[:, : 16 + 48 * rank] = 255            # rank 0 ignores 16 rows, rank 1 64 rows, ...
lr, lab = lr.to(dev), lab.to(dev)

losses.GLOBAL_BATCH_MEAN = True
a = lr.clone().requires_grad_(True)
loss = losses.upsample_cross_entropy(a, lab)
loss.backward()

# single-process statement of the same thing: gather everything, normalise locally over the concatenation
lrs = [torch.empty_like(lr) for _ in range(world)]
labs = [torch.empty_like(lab) for _ in range(world)]
dist.all_gather(lrs, lr)
dist.all_gather(labs, lab)
losses.GLOBAL_BATCH_MEAN = False
b = torch.cat(lrs).clone().requires_grad_(True)
ref = losses.upsample_cross_entropy(b, torch.cat(labs))
ref.backward()
mine = b.grad[rank * n:(rank + 1) * n]
err_l = abs(loss.item() - ref.item()) / abs(ref.item())
err_g = float((a.grad / world - mine).norm() / mine.norm())
local_mean = losses.upsample_cross_entropy(lr, lab).item()      # what a rank-local mean would have returned
print("rank %d: global-mean loss %.6f (single process %.6f, rank-local mean %.6f), loss err %.2e, gradient err %.2e"
      % (rank, loss.item(), ref.item(), local_mean, err_l, err_g), flush=True)
ok = err_l < 1e-6 and err_g < 1e-5 and abs(local_mean - ref.item()) > 1e-4 * abs(ref.item())
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("GLOBAL CE CHECK", "PASSED" if int(flag) else "FAILED", flush=True)
dist.barrier()
os._exit(0 if int(flag) else 1)
