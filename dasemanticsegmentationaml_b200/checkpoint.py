"""Checkpoint I/O of the train loops (train.py:106-118, 280-291; train_nni.py same; BiSeNet.load_weight
model_stages.py:252-258).

The reference saves ``model.module.state_dict()`` (seg net, DataParallel unwrapped) but
``model_D1.state_dict()`` (discriminator, keys carry the ``module.`` prefix — that is what the
shipped ``GTA5_model/GTA5_10_D1.pth`` looks like).  Both layouts load here, in either direction.
"""
import os

import torch


def unwrap(model):
    """The module behind a DataParallel / DistributedDataParallel wrapper (or the module itself)."""
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def save_state(model, path, keep_wrapper_prefix=False):
    """``torch.save(state_dict)``; ``keep_wrapper_prefix`` reproduces the reference's discriminator
    files (``module.``-prefixed keys, train.py:282-283)."""
    sd = unwrap(model).state_dict()
    if keep_wrapper_prefix:
        sd = {"module." + k: v for k, v in sd.items()}
    d = os.path.dirname(os.path.abspath(path))
    os.makedirs(d, exist_ok=True)
    torch.save(sd, path)
    return path


def strip_prefix(state_dict, prefix="module."):
    if state_dict and all(k.startswith(prefix) for k in state_dict):
        return {k[len(prefix):]: v for k, v in state_dict.items()}
    return dict(state_dict)


def load_state(model, source, strict=True, map_location="cpu"):
    """Load a checkpoint file or a state_dict into ``model``; ``module.`` prefixes are tolerated.
    Non-strict mode overlays the matching keys on the model's own state (``BiSeNet.load_weight``
    semantics, model_stages.py:252-258) and returns the keys that were not used."""
    sd = torch.load(source, map_location=map_location) if isinstance(source, (str, os.PathLike)) else source
    sd = strip_prefix(sd)
    target = unwrap(model)
    if strict:
        target.load_state_dict(sd)
        return []
    own = target.state_dict()
    unused = [k for k in sd if k not in own or own[k].shape != sd[k].shape]
    own.update({k: v for k, v in sd.items() if k not in unused})
    target.load_state_dict(own)
    return unused
