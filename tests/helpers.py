import numpy as np
import torch


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def cosine(a, b):
    a, b = a.detach().float().cpu().reshape(-1), b.detach().float().cpu().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def load_oracle_state(module, sd, prefix=""):
    """Copies an oracle state dict (no alias keys) into a product module."""
    own = module.state_dict()
    for k, v in sd.items():
        if not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        assert kk in own and own[kk].shape == v.shape, kk
        own[kk] = v.detach().clone()
    module.load_state_dict(own)
    return module


def sub_state(sd, prefix):
    return {k: v for k, v in sd.items() if k.startswith(prefix)}


def to_device(sd, device, requires_grad=False):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone().to(device)
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def gate(name, value, threshold):
    """Assert value > threshold, printing the measured number so that a GPU run's log records it."""
    print("GATE %-60s %.6f (> %.4f)" % (name, value, threshold))
    assert value > threshold, (name, value, threshold)


class quantised_oracle(object):
    """Context manager: oracle stores conv outputs / activations as bf16 like the CUDA path
    (SURVEY 8c quantisation-matched oracle)."""

    def __enter__(self):
        from oracle import segnet_oracle as O
        self._old, O.QUANT = O.QUANT, True

    def __exit__(self, *a):
        from oracle import segnet_oracle as O
        O.QUANT = self._old
