"""fast_hist on the device must be bit-exact with numpy's bincount formulation (utils.py:161-167)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def np_fast_hist(a, b, n):
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)


@pytest.mark.parametrize("count", [0, 1, 31, 4097, 512 * 1024, 1024 * 2048])
def test_fast_hist_bit_exact(cuda_lib, count):
    from dasemanticsegmentationaml_b200 import kernels as K
    rng = np.random.default_rng(count)
    label = rng.integers(0, 20, size=count).astype(np.int64)
    label[label == 19] = 255
    pred = rng.integers(0, 19, size=count).astype(np.int64)
    hist = torch.zeros(361, dtype=torch.int64, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    K.fast_hist_accumulate(torch.from_numpy(label).cuda(), torch.from_numpy(pred).cuda(), 19, hist, bad)
    K.fast_hist_accumulate(torch.from_numpy(label).cuda(), torch.from_numpy(pred).cuda(), 19, hist, bad)
    ref = 2 * np_fast_hist(label, pred, 19)
    assert bad.item() == 0
    assert np.array_equal(hist.cpu().numpy().reshape(19, 19), ref)


def test_count_equal_batched_bit_exact(cuda_lib):
    """Per-image hit counts of the evaluation loop (utils.py:151-159 per image, train.py:50) for the
    label / prediction dtypes eval produces."""
    from dasemanticsegmentationaml_b200 import kernels as K
    rng = np.random.default_rng(5)
    n, h, w = 5, 37, 61
    label = rng.integers(0, 20, size=(n, h, w)).astype(np.int64)
    label[label == 19] = 255
    pred = rng.integers(0, 19, size=(n, h, w)).astype(np.int64)
    want = (label == pred).reshape(n, -1).sum(1)
    for pdt in (torch.int64, torch.uint8):
        out = torch.zeros(n, dtype=torch.int64, device="cuda")
        K.count_equal_batched(torch.from_numpy(label).cuda(), torch.from_numpy(pred).cuda().to(pdt), out)
        K.count_equal_batched(torch.from_numpy(label).cuda(), torch.from_numpy(pred).cuda().to(pdt), out)
        assert np.array_equal(out.cpu().numpy(), 2 * want)


def test_val_loop_matches_oracle(cuda_lib):
    """train.val (reference train.py:24-61) on the device: the confusion matrix, mIoU and the mean
    per-image precision equal the oracle's restatement of the reference loop applied to the SAME
    predictions, bit for bit (integer counts; float64 host arithmetic in the reference's order)."""
    from oracle import segnet_oracle as O
    from dasemanticsegmentationaml_b200 import train as T
    from dasemanticsegmentationaml_b200.model import BiSeNet
    torch.manual_seed(3)
    m = BiSeNet("STDCNet813", 19).cuda()
    g = torch.Generator().manual_seed(4)
    batches = []
    for nb in (2, 3):
        x = torch.randn(nb, 3, 128, 256, generator=g).cuda()
        lab = torch.randint(0, 20, (nb, 128, 256), generator=g)
        lab[lab == 19] = 255
        batches.append((x, lab.cuda()))
    precision, miou = T.val(m, batches, 19)
    hist = np.zeros((19, 19))
    ratios = []
    for x, lab in batches:
        _, pred = T.eval_batch(m, x, lab, 19, None, pred_dtype=torch.int64)
        for i in range(x.shape[0]):
            p_i, l_i = pred[i].cpu().numpy(), lab[i].cpu().numpy()
            ratios.append(O.compute_global_accuracy(p_i, l_i))
            hist += O.fast_hist(l_i.flatten(), p_i.flatten(), 19)
    assert np.array_equal(T.val.last_hist.cpu().numpy().reshape(19, 19), hist.astype(np.int64))
    assert precision == float(np.mean(ratios))
    assert miou == float(np.mean(O.per_class_iu(hist)))


def test_nccl_c_abi_single_rank(cuda_lib):
    """b200_nccl_* (SURVEY 8b): communicator from a 128-byte unique id, fp32 gradient average and
    int64 confusion-matrix sum; with one rank both are the identity (2-rank run: scripts/check_nccl_abi.py)."""
    import ctypes
    lib = cuda_lib
    uid = ctypes.create_string_buffer(128)
    assert lib.b200_nccl_unique_id(uid) == 0, lib.b200_last_error()
    comm = ctypes.c_void_p()
    assert lib.b200_nccl_init(uid, 1, 0, ctypes.byref(comm)) == 0, lib.b200_last_error()
    g = torch.arange(1000, dtype=torch.float32, device="cuda")
    h = torch.arange(361, dtype=torch.int64, device="cuda") * 12345678901
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.b200_nccl_allreduce_grads(comm, ctypes.c_void_p(g.data_ptr()), ctypes.c_int64(g.numel()), 1, s) == 0
    assert lib.b200_nccl_allreduce_hist(comm, ctypes.c_void_p(h.data_ptr()), 361, s) == 0
    torch.cuda.synchronize()
    assert torch.equal(g, torch.arange(1000, dtype=torch.float32, device="cuda"))
    assert torch.equal(h, torch.arange(361, dtype=torch.int64, device="cuda") * 12345678901)
    assert lib.b200_nccl_destroy(comm) == 0
