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


def _fake_val_batch(model, images, labels, n_classes, hist):
    """CPU stand-in for train._val_batch (the CUDA forward + kernels): the 'prediction' is a fixed
    function of the image, the counting is the oracle's.  What the test exercises is val() itself:
    accumulation over batches, the all-reduces, the host arithmetic."""
    import numpy as np
    from oracle import segnet_oracle as O
    pred = (images[:, 0] * 7 + images[:, 1]).long() % n_classes
    if hist is None:
        hist = torch.zeros(n_classes * n_classes, dtype=torch.int64)
    h = O.fast_hist(labels.numpy().reshape(-1), pred.numpy().reshape(-1), n_classes)
    hist += torch.from_numpy(np.asarray(h, dtype=np.int64).reshape(-1))
    correct = (pred == labels).flatten(1).sum(1)
    return hist, correct


def _dataset(seed=3, n_batches=5, nb=2, h=12, w=20):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        images = torch.randint(0, 19, (nb, 3, h, w), generator=g)
        labels = torch.randint(0, 20, (nb, h, w), generator=g)
        labels[labels == 19] = 255
        out.append((images, labels))
    return out


def _val_worker(rank, world, port, q, shards):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dasemanticsegmentationaml_b200 import train as T
    T._val_batch = _fake_val_batch
    data = _dataset()
    mine = [data[i] for i in shards[rank]]
    precision, miou = T.val(None, mine, 19, device=torch.device("cpu"))
    q.put((rank, precision, miou, T.val.last_hist.tolist()))
    dist.destroy_process_group()


def _run_val(shards):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + len(shards[0])
    procs = [ctx.Process(target=_val_worker, args=(r, 2, port, q, shards)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_val_two_ranks_matches_single_process_bit_for_bit():
    """train.val over a sharded evaluation set (BASELINE config 5; reference train.py:24-61): the
    all-reduced int64 confusion matrix equals the single-process one bit for bit, every rank returns
    the same (precision, mIoU), also with an uneven and with an EMPTY shard."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from dasemanticsegmentationaml_b200 import train as T
    from oracle import segnet_oracle as O
    saved = T._val_batch
    T._val_batch = _fake_val_batch
    try:
        single_p, single_miou = T.val(None, _dataset(), 19, device=torch.device("cpu"))
        single_hist = T.val.last_hist.tolist()
    finally:
        T._val_batch = saved
    # the single-process numbers themselves against the oracle's restatement of the reference loop
    hist = np.zeros((19, 19))
    ratios = []
    for images, labels in _dataset():
        pred = (images[:, 0] * 7 + images[:, 1]).long() % 19
        for i in range(images.shape[0]):
            ratios.append(O.compute_global_accuracy(pred[i].numpy(), labels[i].numpy()))
            hist += O.fast_hist(labels[i].numpy().flatten(), pred[i].numpy().flatten(), 19)
    assert single_hist == hist.astype(np.int64).reshape(-1).tolist()
    assert single_p == float(np.mean(ratios))
    assert single_miou == float(np.mean(O.per_class_iu(hist)))
    for shards in (([0, 1, 2], [3, 4]), ([0, 1, 2, 3, 4], [])):
        for rank, p, miou, h in _run_val(shards):
            assert h == single_hist, "summed confusion matrix differs from the single-process one"
            assert miou == single_miou
            assert abs(p - single_p) < 1e-15
    empty_p, empty_miou = T.val(None, [], 19, device=torch.device("cpu"))
    assert np.isnan(empty_p) and empty_miou == 0.0
