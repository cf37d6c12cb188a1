"""Knock-out profile of the graph-replayed DA step: for each kernel family the step is captured
with that family's launches SKIPPED (results are garbage, timing is not) and the drop in ms/step
is the family's true cost inside the graph — warm caches, no launch gaps, unlike ncu's serialised
cold-cache per-kernel times.  Diagnostic only; nothing here ships.

Usage: python scripts/knockout.py [family ...]     (default: every family)
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import _lib, build, train as T
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator

build.build()
dev = torch.device("cuda", 0)
FAMILIES = {
    "conv_igemm": ["b200_conv_igemm"], "conv_wgrad": ["b200_conv_wgrad"],
    "bn_bwd": ["b200_bn_act_bwd_fused"], "bn_fwd": ["b200_bn_norm_act"],
    "upsample": ["b200_upsample_fwd", "b200_upsample_bwd"], "act_bwd_bias": ["b200_act_bwd_bias"],
    "stem": ["b200_stem_im2col"], "attention": ["b200_fc_small_fwd", "b200_fc_small_bwd", "b200_pool_sum",
                                                                 "b200_scale_add_bcast", "b200_upsum_dot_reduce"],
    "depthwise": ["b200_dwconv_s2_fwd", "b200_dwconv_s2_dgrad", "b200_dwconv_s2_wgrad"],
    "classifier": ["b200_classifier_fwd", "b200_classifier_dgrad", "b200_classifier_wgrad"],
    "pack": ["b200_pack_filters_batched"],
}
SKIP = set()
_orig_call = _lib.call


def call(name, *args, **kw):
    if name in SKIP:
        return
    return _orig_call(name, *args, **kw)


_lib.call = call
import dasemanticsegmentationaml_b200.kernels as K  # noqa: E402
K.call = call


def measure(skip):
    SKIP.clear()
    SKIP.update(skip)
    torch.manual_seed(0)
    model = BiSeNet("STDCNet813", 19).to(dev)
    disc = FCDiscriminator(19).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-4, fused=True)
    opt_d = torch.optim.Adam(disc.parameters(), lr=1e-3, betas=(0.9, 0.99), fused=True, capturable=True)
    g = torch.Generator().manual_seed(100)
    buf = {"images": torch.randn(8, 3, 512, 1024, generator=g).to(dev),
           "labels": torch.randint(0, 19, (8, 512, 1024), generator=g).to(dev),
           "images_t": torch.randn(8, 3, 512, 1024, generator=g).to(dev)}
    gs = T.GraphedStep(lambda **kw: T.train_da_step(model, disc, opt, opt_d, kw["images"], kw["labels"], kw["images_t"]),
                       buf, warmup=2)
    for _ in range(3):
        gs()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gs()
    e1.record()
    torch.cuda.synchronize()
    del gs
    return e0.elapsed_time(e1) / 10


base = measure([])
print("full step            %7.3f ms" % base, flush=True)
torch_only = None
for fam in (sys.argv[1:] or list(FAMILIES)):
    ms = measure(FAMILIES[fam])
    print("without %-12s %7.3f ms   -> %-12s costs %6.3f ms (%4.1f %%)" % (fam, ms, fam, base - ms, 100 * (base - ms) / base), flush=True)
ms = measure([n for v in FAMILIES.values() for n in v] + ["b200_bce_const_fwd", "b200_bce_const_bwd", "b200_scale_f32", "b200_cast_f32_bf16"])
print("without all of ours  %7.3f ms   (torch optimizers, zero fills, copies, graph overhead)" % ms)
