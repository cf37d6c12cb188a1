"""Depthwise stride-2 convolution kernels (csrc/depthwise.cu: avd_layer 3x3 of the STDC bottlenecks with the fused
AvgPool2d(3, 2, 1) skip, stdcnet.py:24-28,73-78; 4x4 with bias of the depthwise-separable discriminators,
discriminator.py:35-45) against torch's fp32 grouped convolution on the same bf16-rounded activations.
fp32 weights and accumulation on both sides: bf16 outputs within one rounding (rel-L2 4e-3), fp32 gradients 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


CASES = [
    # n, c, h, w, k
    (2, 32, 37, 61, 4),      # discriminator conv1 (19 -> 32 padded channels), ragged odd map
    (2, 64, 34, 66, 4),      # maps grown by the k=1, p=1 pointwise convs
    (1, 256, 19, 35, 4),
    (1, 512, 11, 19, 4),
    (2, 128, 32, 64, 3),     # CatBottleneck stride-2 block (with the pooled skip)
    (1, 256, 23, 41, 3),
    (8, 32, 128, 256, 4),    # several CTAs per SM, many tiles
]


@pytest.mark.parametrize("case", CASES)
def test_dwconv_forward(cuda_lib, case):
    from dasemanticsegmentationaml_b200 import kernels as K
    n, c, h, w, k = case
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, h, w, c, device="cuda", generator=g).to(BF)
    wt = torch.randn(c, 1, k, k, device="cuda", generator=g) * 0.3
    bias = torch.randn(c, device="cuda", generator=g) if k == 4 else None
    ho, wo = (h + 2 - k) // 2 + 1, (w + 2 - k) // 2 + 1
    xf = x.float().permute(0, 3, 1, 2).contiguous()
    ref = F.conv2d(xf, wt, bias=bias, stride=2, padding=1, groups=c)
    z = torch.full((n, ho, wo, c + 16), 5.0, device="cuda", dtype=BF)[..., 8:8 + c]   # a channel slice of a wider buffer
    pool = torch.empty(n, ho, wo, c, device="cuda", dtype=BF) if k == 3 else None
    stats = torch.zeros(2, c, device="cuda")
    act, slope = (2, 0.2) if k == 4 else (0, 0.0)
    K.dwconv_s2_fwd(x, k, wt.view(-1), bias, z, pool, act, slope, stats if k == 3 else None)
    torch.cuda.synchronize()
    want = (F.leaky_relu(ref, 0.2) if k == 4 else ref).permute(0, 2, 3, 1)
    assert rel_l2(z, want) < 4e-3
    if k == 3:
        assert rel_l2(pool, F.avg_pool2d(xf, 3, 2, 1).permute(0, 2, 3, 1)) < 4e-3
        flat = z.float().reshape(-1, c)
        assert rel_l2(stats[0], flat.sum(0)) < 1e-3 and rel_l2(stats[1], (flat * flat).sum(0)) < 1e-3


@pytest.mark.parametrize("case", CASES)
def test_dwconv_backward(cuda_lib, case):
    from dasemanticsegmentationaml_b200 import kernels as K
    n, c, h, w, k = case
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.randn(n, h, w, c, device="cuda", generator=g).to(BF)
    wt = torch.randn(c, 1, k, k, device="cuda", generator=g) * 0.3
    ho, wo = (h + 2 - k) // 2 + 1, (w + 2 - k) // 2 + 1
    dz = torch.randn(n, ho, wo, c, device="cuda", generator=g).to(BF)
    dpool = torch.randn(n, ho, wo, c, device="cuda", generator=g).to(BF) if k == 3 else None
    # (contiguous NCHW on the torch side: torch 2.11's CUDA avg_pool2d backward returns wrong values when both the
    #  input and the incoming gradient are channels-last strided views)
    xf = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    bias = torch.zeros(c, device="cuda", requires_grad=True)
    y = F.conv2d(xf, wr, bias=bias, stride=2, padding=1, groups=c)
    loss = (y * dz.float().permute(0, 3, 1, 2).contiguous()).sum()
    if k == 3:
        loss = loss + (F.avg_pool2d(xf, 3, 2, 1) * dpool.float().permute(0, 3, 1, 2).contiguous()).sum()
    loss.backward()
    dx = torch.full((n, h, w, c), 9.0, device="cuda", dtype=BF)
    K.dwconv_s2_dgrad(dz, dpool, k, wt.view(-1), dx)
    dw, db = torch.zeros(c * k * k, device="cuda"), torch.zeros(c, device="cuda")
    K.dwconv_s2_wgrad(dz, x, k, dw, db)
    torch.cuda.synchronize()
    assert rel_l2(dx, xf.grad.permute(0, 2, 3, 1)) < 4e-3
    assert rel_l2(dw.view(c, 1, k, k), wr.grad) < 1e-3
    assert rel_l2(db, bias.grad) < 1e-3
