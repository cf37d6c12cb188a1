"""Times the device input pipeline (csrc/input.cu) at the dataset sizes and reports it against the
HBM roofline: algorithmic bytes = decoded uint8 source read once + fp32 NCHW image / uint8 label
written once."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, dataset as D

build.build()
peak = 6548.2
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        peak = float(json.load(f).get("hbm_gbps_sustained", peak))
except Exception:
    pass
for name, (h0, w0), (a, b) in (("GTA5 1914x1052 -> (512,1024)", (1052, 1914), (512, 1024)),
                               ("Cityscapes 2048x1024 -> (512,1024)", (1024, 2048), (512, 1024)),
                               ("Cityscapes 2048x1024 -> (720,1280)", (1024, 2048), (720, 1280))):
    n = 8
    img = torch.randint(0, 256, (n, h0, w0, 3), dtype=torch.uint8, device="cuda")
    lab = torch.randint(0, 34, (n, h0, w0), dtype=torch.uint8, device="cuda")
    pre = D.DevicePreprocess(a, b, D.GTA5_ID_TO_TRAINID)
    for fn, label, nbytes in ((lambda: pre.images(img), "image", n * (h0 * w0 * 3 + a * b * 12)),
                              (lambda: pre.labels(lab), "label", n * (a * b * 2))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print("%-40s %-6s %8.1f us/batch of %d  %7.1f GB/s  (%.3f of %.0f GB/s)  %.0f img/s"
              % (name, label, us, n, nbytes / us / 1e3, nbytes / us / 1e3 / peak, peak, n / us * 1e6))
