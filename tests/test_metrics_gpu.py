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
