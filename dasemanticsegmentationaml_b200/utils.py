"""Metrics, losses and schedule helpers — drop-in for the hot-path part of the reference's ``utils.py``.

``fast_hist`` keeps the reference signature ``fast_hist(a, b, n)`` (a = label, b = prediction, called
as at train.py:47) and returns the same int64 ``[n, n]`` matrix, bit for bit, but counts on the GPU
(per-warp shared-memory histograms, int64 atomics) when given CUDA tensors.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import kernels as K
from . import losses


def poly_lr_scheduler(optimizer, init_lr, iter, lr_decay_iter=1, max_iter=300, power=0.9):
    """Polynomial learning-rate decay (reference utils.py:11-26)."""
    lr = init_lr * (1 - iter / max_iter) ** power
    optimizer.param_groups[0]['lr'] = lr
    # optim.FusedSGD / FusedAdam keep their hyper-parameters in device memory (that is what makes a
    # step replayable inside a CUDA graph): push the new rate now, so that a GraphedStep replay --
    # which never runs optimizer.step() on the host again -- trains at the decayed rate
    refresh = getattr(optimizer, "refresh_hyperparameters", None)
    if refresh is not None:
        refresh()
    return lr


def reverse_one_hot(image):
    """[C, H, W] scores -> [H, W] int64 class map (reference utils.py:98-122)."""
    return torch.argmax(image.permute(1, 2, 0), dim=-1)


def fast_hist_device(a, b, n, hist=None):
    """Accumulates the confusion matrix of CUDA tensors into a device int64 ``[n*n]`` tensor and
    returns it without synchronising (use this inside evaluation loops; sum over ranks with NCCL)."""
    if not (torch.is_tensor(a) and a.is_cuda):
        raise _lib.B200Error("fast_hist_device needs CUDA tensors; there is no CPU fallback")
    _lib.ensure_device(a.device.index or 0)
    if hist is None:
        hist = torch.zeros(n * n, dtype=torch.int64, device=a.device)
    if b.dtype not in (torch.int64, torch.int32, torch.uint8):
        b = b.long()
    if a.dtype != torch.int64 and not (a.dtype == b.dtype and a.dtype in (torch.int32, torch.uint8)):
        a = a.long()
    if a.dtype == torch.int64 and b.dtype == torch.int32:
        b = b.long()
    bad = torch.zeros(1, dtype=torch.int32, device=a.device)
    K.fast_hist_accumulate(a.reshape(-1), b.reshape(-1), n, hist, bad)
    hist._b200_bad = bad
    return hist


def fast_hist(a, b, n):
    """Confusion matrix ``bincount(n * a[k] + b[k], minlength=n**2).reshape(n, n)`` over
    ``k = (a >= 0) & (a < n)`` (reference utils.py:161-167); a = label, b = prediction.
    Accepts CUDA tensors (counted on the GPU) and returns a numpy int64 array like the reference."""
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        raise _lib.B200Error("fast_hist: pass CUDA tensors (the numpy path of the reference is the "
                             "CPU baseline, not part of this package)")
    hist = fast_hist_device(a, b, n)
    out = hist.cpu().numpy().reshape(n, n)
    if int(hist._b200_bad.item()) != 0:
        raise ValueError("fast_hist: a prediction outside [0, n) produced an index beyond n*n "
                         "(numpy would fail to reshape)")
    return out


def per_class_iu(hist):
    """Per-class IoU in float64 on the host, from the exact integer matrix (reference utils.py:170-172)."""
    epsilon = 1e-5
    hist = np.asarray(hist)
    return (np.diag(hist)) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + epsilon)


def compute_global_accuracy(pred, label):
    """Fraction of equal positions (reference utils.py:151-159, a Python loop there)."""
    if torch.is_tensor(pred) and pred.is_cuda:
        out = torch.zeros(1, dtype=torch.int64, device=pred.device)
        lab = label if label.dtype == torch.int64 else label.long()
        pr = pred if pred.dtype in (torch.int64, torch.uint8) else pred.long()
        K.count_equal(lab.reshape(-1).contiguous(), pr.reshape(-1).contiguous(), out)
        return float(out.item()) / float(lab.numel())
    raise _lib.B200Error("compute_global_accuracy needs CUDA tensors; there is no CPU fallback")


class OHEM_CrossEntroy_Loss(nn.Module):
    """Online hard example mining CE (reference utils.py:256-271).  ``forward(output, target)``
    accepts either full-size logits ``[N, C, H, W]`` like the reference, or the low-resolution fp32
    NHWC logits of ``BiSeNet.forward_lowres`` (then the bilinear up-sampling is fused in)."""

    def __init__(self, threshold, keep_num):
        super(OHEM_CrossEntroy_Loss, self).__init__()
        self.threshold = threshold
        self.keep_num = keep_num

    def forward(self, output, target):
        if output.dim() == 4 and output.shape[-1] == 32 and output.dtype == torch.float32 and \
                output.shape[1:3] != target.shape[-2:]:
            return losses.upsample_ohem_cross_entropy(output, target, self.threshold, self.keep_num)
        n, c, h, w = output.shape
        lr = torch.zeros((n, h, w, 32), dtype=torch.float32, device=output.device)
        lr[..., :c] = output.permute(0, 2, 3, 1)
        return losses.upsample_ohem_cross_entropy(lr, target, self.threshold, self.keep_num, n_classes=c)
