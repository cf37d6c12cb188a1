"""A/B launches of the CTA-pair kernel against the one-CTA kernels on a few shapes (CUDA-event
timing; with NCU=1 one launch per configuration for `ncu --set full -k regex:conv_igemm`).
TUNES = comma-separated tune words, CASE = index into CASES."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K

build.build()
dev = "cuda"
CASES = [  # n, cin, cout, h, w, r, stride, pad, stats
    (8, 256, 256, 64, 128, 3, 1, 1, True),     # conv_out 3x3
    (8, 128, 64, 64, 128, 3, 1, 1, True),      # bottleneck tail, BN = 64
    (8, 128, 256, 128, 256, 4, 2, 1, False),   # discriminator conv3
    (8, 64, 128, 256, 512, 4, 2, 1, False),    # discriminator conv2
    (8, 1024, 128, 16, 32, 3, 1, 1, True),     # arm32: small M, deep K
    (8, 384, 256, 64, 128, 1, 1, 0, True),     # ffm 1x1
    (8, 64, 128, 128, 256, 1, 1, 0, True),     # stage-3 1x1, memory bound (M262144 N128 K64)
    # shapes where the launcher's heuristic (tune = 0) was >= 1.5x off the tuner's best (profiles/r2_tile_sweep_epilogue.txt)
    (8, 1024, 128, 16, 32, 3, 1, 1, True),     # 7: tuned 28.3 us
    (8, 512, 256, 16, 32, 3, 1, 1, True),      # 8: tuned 19.5 us
    (8, 256, 128, 32, 64, 3, 1, 1, True),      # 9: tuned 16.7 us
    (8, 32, 32, 128, 256, 3, 1, 1, False),     # 10: tuned 21.0 us
    (8, 32, 32, 90, 160, 3, 1, 1, True),       # 11: tuned 17.0 us
    (8, 1024, 128, 23, 40, 3, 1, 1, True),     # 12: tuned 35.4 us
]
if os.environ.get("CASE"):
    CASES = [CASES[int(c)] for c in os.environ["CASE"].split(",")]
reps = int(os.environ.get("REPS", "1"))
for (n, cin, cout, h, w, r, stride, pad, st) in CASES:
    x = torch.randn(n, h, w, cin, device=dev).to(torch.bfloat16)
    wgt = torch.randn(cout, cin, r, r, device=dev) * 0.05
    filt = K.pack_filter(wgt)
    geom = K.fwd_geometry(h, w, r, r, stride, pad)
    out = torch.empty((n, geom.Hout, geom.Wout, filt.shape[0]), device=dev, dtype=torch.bfloat16)
    stats = torch.zeros(2, cout, device=dev) if st else None
    key = K.conv_key(n, h, w, filt.shape[2], filt.shape[0], geom, 0, st)
    bn = min(256, cout)
    P = 1 << 22
    if os.environ.get("TUNES"):
        tunes = [int(t) for t in os.environ["TUNES"].split(",")]
    else:
        tunes = [0, K.TUNED.get(key, 0), bn | P, bn | P | (1 << 24), bn | P | (2 << 24), bn | P | (4 << 24)]
        if bn > 64:
            tunes += [(bn // 2) | P, (bn // 2) | P | (2 << 24)]
    flops = 2.0 * n * geom.Hout * geom.Wout * cout * cin * r * r
    for tune in tunes:
        for _ in range(reps):
            K.conv_igemm(x, filt, out, geom, stats=stats, bn_tile=tune)
        torch.cuda.synchronize()
        if os.environ.get("NCU"):
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.conv_igemm(x, filt, out, geom, stats=stats, bn_tile=tune)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print("case %-34s tune %9d (BN%d mt%d st%d %s kg%d)  %7.1f us  %6.0f TF/s" % (
            (n, cin, cout, h, w, r, stride), tune, tune & 0xfff, (tune >> 12) & 15, (tune >> 16) & 15,
            "PAIR" if tune & P else ("P" if tune & (1 << 20) else "-"), (tune >> 24) & 7, us, flops / us / 1e6), flush=True)
print("done")
