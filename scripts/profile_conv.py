"""Runs a few representative implicit-GEMM launches (for `ncu --set full -k regex:igemm`)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K

build.build()
dev = "cuda"
CASES = [  # n, cin, cout, h, w, r, stride, pad
    (8, 256, 256, 64, 128, 3, 1, 1),    # conv_out 3x3: 27 % of the seg forward
    (8, 128, 256, 128, 256, 4, 2, 1),   # discriminator conv3
    (8, 128, 64, 64, 128, 3, 1, 1),     # bottleneck tail
    (8, 32, 64, 512, 1024, 4, 2, 1),    # discriminator conv1 (19 -> padded 32 channels)
    (8, 64, 128, 128, 256, 1, 1, 0),    # block-1 1x1: store/epilogue bound (100 MB in+out)
]
if os.environ.get("CASE"):
    CASES = [CASES[int(os.environ["CASE"])]]
reps = int(os.environ.get("REPS", "1"))
for (n, cin, cout, h, w, r, stride, pad) in CASES:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wgt = torch.randn(cout, cin, r, r, device=dev) * 0.05
    filt = K.pack_filter(wgt)
    geom = K.fwd_geometry(h, w, r, r, stride, pad)
    out = torch.empty((n, geom.Hout, geom.Wout, filt.shape[0]), device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, cout, device=dev)
    for _ in range(reps):
        K.conv_igemm(x, filt, out, geom, stats=stats)
    dz = torch.randn(n, geom.Hout, geom.Wout, cout, device=dev).to(torch.bfloat16)
    dw = torch.zeros_like(wgt)
    if not os.environ.get("NO_WGRAD"):
        for _ in range(reps):
            K.conv_wgrad(dz, x, dw, r, r, stride, pad)
torch.cuda.synchronize()
print("done")
