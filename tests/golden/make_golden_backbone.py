"""Golden vectors pinning the oracle's AddBottleneck / ``use_conv_last`` restatements
(oracle.segnet_oracle.add_bottleneck, stdcnet813(block=, use_conv_last=)) to the UNMODIFIED
reference ``model/stdcnet.py`` (STDCNet813(type="add"), STDCNet813(use_conv_last=True)).

Run in the build container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_backbone.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from oracle import segnet_oracle as O  # noqa: E402
from model.stdcnet import STDCNet813  # noqa: E402  (reference)


def sample(t, step=53):
    a = t.detach().double().reshape(-1).numpy()
    return np.concatenate([a[::step], [a.sum(), np.abs(a).sum(), float(a.size)]])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {}
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 64, 128, generator=g)
    out["x"] = x.numpy()
    for tag, block, last in (("add", "add", False), ("cat_last", "cat", True)):
        sd = O.make_backbone_state(seed=17, block=block, use_conv_last=last, prefix="")
        net = STDCNet813(type=block, use_conv_last=last)
        own = net.state_dict()
        for k, v in sd.items():
            assert k in own and own[k].shape == v.shape, k
            own[k] = v.clone()
        net.load_state_dict(own)
        net.train()
        feats = net(x)
        gw = torch.Generator().manual_seed(22)
        loss = sum((f * torch.randn(f.shape, generator=gw)).sum() for f in feats)
        loss.backward()
        for i, f in enumerate(feats):
            out["%s_feat%d" % (tag, i)] = sample(f)
        names = ["features.2.conv_list.0.conv.weight", "features.4.conv_list.2.conv.weight", "features.6.avd_layer.0.weight",
                 "features.7.conv_list.3.bn.weight"]
        if block == "add":
            names += ["features.2.skip.0.weight", "features.4.skip.2.weight", "features.6.skip.3.bias"]
        if last:
            names += ["conv_last.conv.weight", "conv_last.bn.bias"]
        params = dict(net.named_parameters())
        for n in names:
            out["%s_grad:%s" % (tag, n)] = sample(params[n].grad, 29)
        net.eval()
        with torch.no_grad():
            out["%s_eval_feat4" % tag] = sample(net(x)[4])
    np.savez_compressed(os.path.join(HERE, "reference_backbone.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
