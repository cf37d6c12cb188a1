"""torchrun --nproc-per-node 2 scripts/check_reducer.py

Checks the overlapped in-place gradient exchange (ops.REDUCER) on real multi-GPU NCCL: every
exchanged span is snapshotted just before its all-reduce; afterwards (1) the span must equal the mean
of the ranks' snapshots -- nothing wrote into it after it was exchanged, nothing was exchanged twice;
(2) every parameter gradient must lie inside an exchanged span from the second iteration on;
(3) the gradients must be bit-identical on all ranks.  (Comparing against a separately computed flat
all-reduce is meaningless here: train-mode gradients of the random-init net differ by ~50 % between
two runs of the same backward because of fp32 atomic ordering.)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dasemanticsegmentationaml_b200 import build, losses, ops, train as T
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
if rank == 0:
    build.build()
dist.barrier()
dev = torch.device("cuda", local)
torch.manual_seed(0)
m, d = BiSeNet("STDCNet813", 19).to(dev).train(), FCDiscriminator(19).to(dev).train()
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(4, 3, 256, 512, generator=g).to(dev)
lab = torch.randint(0, 19, (4, 256, 512), generator=g).to(dev)
params = list(m.parameters()) + list(d.parameters())
ok = True
for it in range(4):
    for p in params:
        p.grad = None
    ops.REDUCER.snapshots = []
    loss, lr = T.supervised_loss(m, x, lab)
    dout = d(losses.upsample_softmax(lr[0], 256, 512))
    (loss + losses.bce_with_logits_const(dout, 0.0)).backward()
    spans = [(lo, hi) for lo, hi, _ in ops.REDUCER.spans]
    grads = [p.grad for p in params if p.grad is not None]
    inside = sum(1 for t in grads if any(lo <= t.data_ptr() and t.data_ptr() + 4 * t.numel() <= hi for lo, hi in spans))
    T.allreduce_grads(params)
    torch.cuda.synchronize()
    worst = 0.0
    for view, snap in ops.REDUCER.snapshots:
        parts = [torch.empty_like(snap) for _ in range(world)]
        dist.all_gather(parts, snap)
        mean = sum(parts) / world
        scale = float(mean.abs().max()) + 1e-20
        worst = max(worst, float((view - mean).abs().max()) / scale)
    same = True
    for t in grads:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        same = same and all(torch.equal(parts[0], q) for q in parts[1:])
    print("rank %d iter %d: %d spans, %d of %d gradients inside them, span-vs-mean worst %.2e, identical on all ranks: %s"
          % (rank, it, len(spans), inside, len(grads), worst, same), flush=True)
    ok = ok and worst < 1e-5 and same and (it == 0 or inside == len(grads)) and len(spans) >= (1 if it == 0 else 3)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("REDUCER CHECK", "PASSED" if int(flag) else "FAILED", flush=True)
dist.barrier()
os._exit(0 if int(flag) else 1)
