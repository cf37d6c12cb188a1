"""Runs the fused up-sampling kernels once each (for `ncu --set full -k regex:upsample`) and times them."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K

build.build()
dev = "cuda"
n, h, w, H, W = 8, 64, 128, 512, 1024
reps = int(os.environ.get("REPS", "1"))
torch.manual_seed(0)
lr = torch.randn(n, h, w, 32, device=dev)
labels = torch.randint(0, 19, (n, H, W), device=dev)
p = torch.empty((n, H, W, 32), dtype=torch.bfloat16, device=dev)
dp = torch.randn(n, H, W, 32, device=dev).to(torch.bfloat16)
d_lr = torch.zeros_like(lr)
acc = torch.zeros(2, dtype=torch.float64, device=dev)


def timeit(name, fn):
    fn()
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("%-28s %8.1f us" % (name, e0.elapsed_time(e1) / reps * 1e3))


timeit("fwd softmax", lambda: K.upsample_fwd(lr, H, W, 19, K.UP_SOFTMAX, out=p, p_ld=32))
timeit("fwd ce", lambda: K.upsample_fwd(lr, H, W, 19, K.UP_CE, labels=labels, ignore_index=255, acc=acc))
timeit("bwd ce (+loss)", lambda: K.upsample_bwd(lr, H, W, 19, K.UP_CE, d_lr, labels=labels, ignore_index=255, loss_acc=acc))
timeit("bwd softmax", lambda: K.upsample_bwd(lr, H, W, 19, K.UP_SOFTMAX, d_lr, grad_in=dp, grad_is_bf16=1, p_ld=32))
torch.cuda.synchronize()
print("done")
