"""world_size-2 gloo test of the data-parallel host logic: flat-bucket gradient averaging that skips
grad-less parameters, and the int64 confusion-matrix sum."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dasemanticsegmentationaml_b200 import train as T
    params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4))]
    params[0].grad = torch.full((5,), float(rank + 1))
    params[1].grad = torch.arange(6.0).reshape(3, 2) * (rank + 1)
    # params[2] has no gradient on any rank (dead classifier head)
    T.allreduce_grads(params)
    hist = torch.arange(361, dtype=torch.int64) * (rank + 1)
    dist.all_reduce(hist)
    q.put((rank, params[0].grad.tolist(), params[1].grad.tolist(), params[2].grad is None, hist.sum().item()))
    dist.destroy_process_group()


def test_two_rank_gradient_average_and_hist_sum():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g0, g1, dead, hsum in results:
        assert g0 == [1.5] * 5
        assert g1 == (torch.arange(6.0).reshape(3, 2) * 1.5).tolist()
        assert dead
        assert hsum == sum(range(361)) * 3
