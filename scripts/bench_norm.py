"""Times the BatchNorm / activation passes of norm.cu at the layer shapes of the DA step and
reports them against the HBM roofline (algorithmic bytes: every tensor read once, written once)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K

build.build()
dev = "cuda"
BF = torch.bfloat16
PEAK = 6548.2
SHAPES = [(1048576, 32), (262144, 64), (262144, 128), (65536, 256), (65536, 128), (65536, 64), (65536, 32),
          (16384, 512), (16384, 256), (16384, 128), (16384, 64), (4096, 512), (4096, 256), (4096, 128)]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = {"fwd": 0.0, "bwd": 0.0, "bwd2": 0.0}
for npix, c in SHAPES:
    z = torch.randn(1, 1, npix, c, device=dev).to(BF)
    a = torch.empty_like(z)
    dy = torch.randn(1, 1, npix, c, device=dev).to(BF)
    dy2 = torch.randn(1, 1, npix, c, device=dev).to(BF)
    dz = torch.empty_like(z)
    stats = torch.stack([z.float().sum((0, 1, 2)), z.float().square().sum((0, 1, 2))]).contiguous()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    buf = torch.empty(4, c, device=dev)
    red = torch.zeros(2, c, device=dev)
    f = timeit(lambda: K.bn_norm_act(z, a, stats, float(npix), gamma, beta, rm, rv, True, 1, 0.0, buf[0], buf[1], buf[2], buf[3]))
    b = timeit(lambda: K.bn_act_bwd_fused(dy, None, z, dz, buf[0], buf[1], buf[2], buf[3], red, 1, 0.0))
    b2 = timeit(lambda: K.bn_act_bwd_fused(dy, dy2, z, dz, buf[0], buf[1], buf[2], buf[3], red, 1, 0.0))
    e = npix * c
    print("px%-8d C%-4d  fwd %6.1f us %5.0f GB/s (%.2f) | bwd %6.1f us %5.0f GB/s (%.2f) | bwd+dy2 %6.1f us %5.0f GB/s (%.2f)"
          % (npix, c, f, 4 * e / f / 1e3, 4 * e / f / 1e3 / PEAK, b, 6 * e / b / 1e3, 6 * e / b / 1e3 / PEAK,
             b2, 8 * e / b2 / 1e3, 8 * e / b2 / 1e3 / PEAK))
    tot["fwd"] += f; tot["bwd"] += b; tot["bwd2"] += b2
print("sum over shapes (us):", tot, "(back-to-back launches on warm L2: small shapes are L2-resident)")
