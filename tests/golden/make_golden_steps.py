"""Golden vectors for the two adversarial train-loop bodies, produced with the UNMODIFIED reference
modules (build container only; ``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_steps.py

``train.py`` / ``train_nni.py`` themselves need CUDA + AMP, so the loop bodies (train.py:192-262,
train_nni.py:105-163) are replayed here statement by statement on the CPU in fp32 with the
reference's ``BiSeNet`` / ``FCDiscriminator`` and torch's own losses / optimizers.  Output:
``tests/golden/reference_steps.npz`` (losses and samples of updated parameters).
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
from tests.golden.make_golden import load_into, sample  # noqa: E402

from model.model_stages import BiSeNet  # noqa: E402  (reference)
from model.discriminator import FCDiscriminator  # noqa: E402

KEYS = ("conv_out.conv_out.weight", "conv_out32.conv_out.weight", "cp.backbone.features.0.conv.weight",
        "cp.arm16.conv.bn.running_var")
D_KEYS = ("conv1.weight", "classifier.bias")


def fresh():
    torch.manual_seed(0)
    net, d = BiSeNet("STDCNet813", 19), FCDiscriminator(19)
    load_into(net, O.make_bisenet_state(seed=31))
    load_into(d, O.make_discriminator_state("dense", seed=6))
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4)
    opt_d = torch.optim.Adam(d.parameters(), lr=1e-3, betas=(0.9, 0.99))
    return net, d, opt, opt_d


def inputs():
    g = torch.Generator().manual_seed(17)
    x = torch.randn(2, 3, 64, 128, generator=g)
    xt = torch.randn(2, 3, 64, 128, generator=g)
    labels = torch.randint(0, 20, (2, 1, 64, 128), generator=g)
    labels[labels == 19] = 255
    return x, labels, xt


def main():
    torch.set_num_threads(8)
    out = {}
    x, labels, xt = inputs()
    out["x"], out["xt"], out["labels"] = x.numpy(), xt.numpy(), labels.numpy().astype(np.int16)
    loss_func = torch.nn.CrossEntropyLoss(ignore_index=255)
    bce = torch.nn.BCEWithLogitsLoss()
    lam = 0.01

    # ---- train.py:192-262 ---------------------------------------------------------------------
    model, model_D1, optimizer, optimizer_D1 = fresh()
    model.train(), model_D1.train()
    optimizer.zero_grad(), optimizer_D1.zero_grad()
    for p in model_D1.parameters():
        p.requires_grad = False
    output, out16, out32 = model(x)
    loss = loss_func(output, labels.squeeze(1)) + loss_func(out16, labels.squeeze(1)) + loss_func(out32, labels.squeeze(1))
    loss.backward()
    optimizer.step()
    output_t, _, _ = model(xt)
    optimizer.zero_grad()
    D_out1 = model_D1(torch.softmax(output_t, dim=1))
    loss_adv = bce(D_out1, torch.zeros_like(D_out1))
    (loss_adv * lam).backward()
    optimizer.step()
    for p in model_D1.parameters():
        p.requires_grad = True
    output, output_t = output.detach(), output_t.detach()
    D_out1 = model_D1(torch.softmax(output, dim=1))
    loss_d_s = bce(D_out1, torch.zeros_like(D_out1))
    loss_d_s.backward()
    optimizer_D1.step()
    D_out1 = model_D1(torch.softmax(output_t, dim=1))
    loss_d_t = bce(D_out1, torch.ones_like(D_out1))
    optimizer_D1.zero_grad()
    loss_d_t.backward()
    optimizer_D1.step()
    out["da_losses"] = np.array([loss.item(), loss_adv.item(), loss_d_s.item(), loss_d_t.item()])
    for k in KEYS:
        out["da:" + k] = sample(model.state_dict()[k], 7)
    for k in D_KEYS:
        out["da_d:" + k] = sample(model_D1.state_dict()[k], 7)

    # ---- train_nni.py:105-163 -----------------------------------------------------------------
    model, model_D1, optimizer, optimizer_D1 = fresh()
    model.train(), model_D1.train()
    for p in model_D1.parameters():
        p.requires_grad = False
    optimizer.zero_grad(), optimizer_D1.zero_grad()
    output, out16, out32 = model(x)
    loss = loss_func(output, labels.squeeze(1)) + loss_func(out16, labels.squeeze(1)) + loss_func(out32, labels.squeeze(1))
    loss.backward()
    output_t, out16_t, out32_t = model(xt)
    D_out1 = model_D1(torch.softmax(out32_t, dim=1))
    loss_D1 = bce(D_out1, torch.zeros_like(D_out1)) * lam
    loss_D1.backward()
    for p in model_D1.parameters():
        p.requires_grad = True
    out32, out32_t = out32.detach(), out32_t.detach()
    D_out1 = model_D1(torch.softmax(out32, dim=1))
    loss_s = bce(D_out1, torch.zeros_like(D_out1))
    loss_s.backward()
    D_out1 = model_D1(torch.softmax(out32_t, dim=1))
    loss_t = bce(D_out1, torch.ones_like(D_out1))
    loss_t.backward()
    optimizer.step()
    optimizer_D1.step()
    out["nni_losses"] = np.array([loss.item(), loss_D1.item(), loss_s.item(), loss_t.item()])
    for k in KEYS:
        out["nni:" + k] = sample(model.state_dict()[k], 7)
    for k in D_KEYS:
        out["nni_d:" + k] = sample(model_D1.state_dict()[k], 7)
    np.savez_compressed(os.path.join(HERE, "reference_steps.npz"), **out)
    print("wrote", len(out), "arrays", out["da_losses"], out["nni_losses"])


if __name__ == "__main__":
    main()
