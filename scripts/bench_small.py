"""GPU time of the small (non-tensor-core) kernels at the layer shapes of the DA step.

Each kernel is captured `REPS` times into a CUDA graph and the graph is replayed: what is measured
is the kernel inside a graph (no host launch overhead, warm instruction cache), which is how the
train step runs it.  Tensors of the big shapes exceed L2 only for the first layers; the note column
gives the algorithmic HBM bytes and the fraction of the measured copy bandwidth.

Usage: python scripts/bench_small.py [name-filter ...]
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dasemanticsegmentationaml_b200 import build, kernels as K, ops
from dasemanticsegmentationaml_b200.model import BiSeNet, FCDiscriminator

build.build()
dev = torch.device("cuda", 0)
BF, F32 = torch.bfloat16, torch.float32
PEAK = 6548.2
REPS = 10
FILT = sys.argv[1:]


def act(n, h, w, c):
    return torch.randn(n, h, w, c, device=dev).to(BF)


def run(name, fn, nbytes=0):
    if FILT and not any(f in name for f in FILT):
        return
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (3 * REPS) * 1e3
    note = ""
    if nbytes:
        note = "%7.1f MB  %6.0f GB/s (%.2f of HBM)" % (nbytes / 1e6, nbytes / us / 1e3, nbytes / us / 1e3 / PEAK)
    print("%-46s %8.1f us  %s" % (name, us, note), flush=True)


N = 8
# ---- depthwise 3x3 s2 (+ avg-pool skip, BN statistics): CatBottleneck stride-2 blocks
for c, h, w in ((128, 128, 256), (256, 64, 128), (512, 32, 64)):
    x = act(N, h, w, c)
    wt = torch.randn(c, 1, 3, 3, device=dev)
    z, pool = act(N, h // 2, w // 2, c), act(N, h // 2, w // 2, c)
    stats = torch.zeros(2, c, device=dev)
    dz, dpool, dx = act(N, h // 2, w // 2, c), act(N, h // 2, w // 2, c), act(N, h, w, c)
    dw, e = torch.zeros(c, 1, 3, 3, device=dev), x.numel() * 2
    run("dwconv_s2_fwd  C%d %dx%d +pool +stats" % (c, h, w), lambda: K.dwconv_s2_fwd(x, 3, wt, None, z, pool, 0, 0.0, stats), e + e // 2)
    run("dwconv_s2_dgrad C%d %dx%d +dpool" % (c, h, w), lambda: K.dwconv_s2_dgrad(dz, dpool, 3, wt, dx), e + e // 2)
    run("dwconv_s2_wgrad C%d %dx%d" % (c, h, w), lambda: K.dwconv_s2_wgrad(dz, x, 3, dw, None), e + e // 4)
# ---- depthwise 4x4 s2 (+bias) of the depthwise-separable discriminators (config[3]; the pointwise k=1, p=1 convs
#      grow every map by 2 pixels)
for c, h, w in ((32, 512, 1024), (64, 258, 514), (128, 131, 259), (256, 67, 131), (512, 35, 67)):
    x = act(N, h, w, c)
    wt, bs = torch.randn(c, 1, 4, 4, device=dev), torch.randn(c, device=dev)
    ho, wo = (h + 2 - 4) // 2 + 1, (w + 2 - 4) // 2 + 1
    z, dz, dx = act(N, ho, wo, c), act(N, ho, wo, c), act(N, h, w, c)
    dw, db, e = torch.zeros(c, 1, 4, 4, device=dev), torch.zeros(c, device=dev), x.numel() * 2
    run("dwconv_s2_fwd  k4 C%d %dx%d +bias" % (c, h, w), lambda: K.dwconv_s2_fwd(x, 4, wt, bs, z, None, 0, 0.0, None), e + e // 4)
    run("dwconv_s2_dgrad k4 C%d %dx%d" % (c, h, w), lambda: K.dwconv_s2_dgrad(dz, None, 4, wt, dx), e + e // 4)
    run("dwconv_s2_wgrad k4 C%d %dx%d +dbias" % (c, h, w), lambda: K.dwconv_s2_wgrad(dz, x, 4, dw, db), e + e // 4)
# ---- discriminator head (Cout = 1, 4x4 s2) on [8, 32, 64, 512]
x = act(N, 32, 64, 512)
wt, b = torch.randn(1, 512, 4, 4, device=dev) * 0.01, torch.zeros(1, device=dev)
out = torch.empty(N, 16, 32, 1, device=dev)
dout = torch.randn(N, 16, 32, 1, device=dev)
dx, dw, db = act(N, 32, 64, 512), torch.zeros(1, 512, 4, 4, device=dev), torch.zeros(1, device=dev)
e = x.numel() * 2
run("classifier_fwd   [8,32,64,512]", lambda: K.classifier_fwd(x, wt, b, out), e)
run("classifier_dgrad [8,32,64,512]", lambda: K.classifier_dgrad(dout, wt, dx), e)
run("classifier_wgrad [8,32,64,512]", lambda: K.classifier_wgrad(dout, x, dw, db), e)
# ---- filter packing of whole modules
seg, disc = BiSeNet("STDCNet813", 19).to(dev), FCDiscriminator(19).to(dev)
xs = torch.randn(2, 3, 64, 128, device=dev)
seg.train(), disc.train()
o = seg.forward_lowres(xs)
(o[0].float().sum() + o[1].float().sum() + o[2].float().sum()).backward()
p = torch.softmax(torch.randn(2, 19, 64, 128, device=dev), 1).to(BF).contiguous(memory_format=torch.channels_last)
p.requires_grad_(True)
disc(p).float().sum().backward()
nseg = sum(q.numel() for q in seg.parameters() if q.dim() == 4)
ndis = sum(q.numel() for q in disc.parameters() if q.dim() == 4)
run("pack_filters BiSeNet (%.1f M weights)" % (nseg / 1e6), lambda: ops.refresh_packs(seg, force=True), nseg * 8)
run("pack_filters FCDiscriminator (%.1f M)" % (ndis / 1e6), lambda: ops.refresh_packs(disc, force=True), ndis * 8)
# ---- attention pieces
for name, c, h, w in (("feat32", 1024, 16, 32), ("arm32", 128, 16, 32), ("arm16", 128, 32, 64), ("ffm", 256, 64, 128)):
    x = act(N, h, w, c)
    out = torch.zeros(N, c, device=dev)
    run("pool_sum %s C%d %dx%d" % (name, c, h, w), lambda: K.pool_sum(x, out), x.numel() * 2)
for name, c, hs, ws, ho, wo, has_t in (("arm32 -> x2", 128, 16, 32, 32, 64, False), ("arm16 -> x2", 128, 32, 64, 64, 128, True),
                                       ("ffm", 256, 64, 128, 64, 128, False)):
    a, s, v = act(N, hs, ws, c), torch.rand(N, c, device=dev), torch.rand(N, c, device=dev)
    t = act(N, hs, ws, c) if has_t else None
    o = act(N, ho, wo, c)
    run("scale_add_bcast %s C%d" % (name, c), lambda: K.scale_add_bcast(a, s, 0.0, v, 1.0, t, o), (a.numel() + o.numel()) * 2)
    dsum, dot, vs = act(N, hs, ws, c), torch.zeros(N, c, device=dev), torch.zeros(N, c, device=dev)
    run("upsum_dot_reduce %s C%d" % (name, c), lambda: K.upsum_dot_reduce(o, a, dsum, hs, ws, dot, vs), (a.numel() * 2 + o.numel()) * 2)
# ---- tiny dense layers
for name, cin, co, bn, actn in (("conv_avg 1024->128 BN relu", 1024, 128, True, 1), ("arm 128->128 BN sigmoid", 128, 128, True, 3),
                                ("ffm 256->64 relu", 256, 64, False, 1), ("ffm 64->256 sigmoid", 64, 256, False, 3)):
    inp, W = torch.randn(N, cin, device=dev), torch.randn(co, cin, device=dev) * 0.05
    bnp = (torch.ones(co, device=dev), torch.zeros(co, device=dev), torch.zeros(co, device=dev), torch.ones(co, device=dev)) if bn else None
    pre, outp = torch.empty(N, co, device=dev), torch.empty(N, co, device=dev)
    mean, rstd = torch.empty(co, device=dev), torch.empty(co, device=dev)
    run("fc_small_fwd %s" % name, lambda: K.fc_small_fwd(inp, 1.0, W, bnp, True, actn, pre, outp, mean, rstd))
    dout, scratch = torch.randn(N, co, device=dev), torch.empty(N, co, device=dev)
    dW, dg, dbt, din = torch.zeros_like(W), torch.zeros(co, device=dev), torch.zeros(co, device=dev), torch.empty(N, cin, device=dev)
    run("fc_small_bwd %s" % name, lambda: K.fc_small_bwd(dout, outp, pre, inp, 1.0, W, bn, True, bnp[0] if bn else None, mean, rstd,
                                                          actn, scratch, dW, dg if bn else None, dbt if bn else None, din))
# ---- LeakyReLU backward + bias gradient of the dense discriminator
for c, h, w in ((64, 256, 512), (128, 128, 256), (256, 64, 128), (512, 32, 64)):
    dy, a, dz, dbias = act(N, h, w, c), act(N, h, w, c), act(N, h, w, c), torch.zeros(c, device=dev)
    run("act_bwd_bias C%d %dx%d" % (c, h, w), lambda: K.act_bwd_bias(dy, None, a, dz, 2, 0.2, dbias), dy.numel() * 6)
# ---- stem unfold
img = torch.randn(N, 3, 512, 1024, device=dev)
col = act(N, 256, 512, 32)
run("stem_im2col 512x1024", lambda: K.stem_im2col(img, col), img.numel() * 4 + col.numel() * 2)
# ---- BatchNorm passes (forward normalise+act, fused backward) at the step's layer shapes
for npix, c in ((1048576, 32), (262144, 64), (262144, 128), (65536, 256), (65536, 128), (65536, 64), (65536, 32),
                (16384, 512), (16384, 256), (16384, 128), (16384, 64), (4096, 512), (4096, 256), (4096, 128)):
    z, a, dy, dy2, dz = (act(1, 1, npix, c) for _ in range(5))
    stats = torch.stack([z.float().sum((0, 1, 2)), z.float().square().sum((0, 1, 2))]).contiguous()
    gamma, beta, rm, rv = torch.ones(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.ones(c, device=dev)
    buf, red = torch.empty(4, c, device=dev), torch.zeros(1 + K.BN_RED_REPLICAS, 2, c, device=dev)
    e = npix * c
    run("bn_norm_act px%d C%d" % (npix, c), lambda: K.bn_norm_act(z, a, stats, float(npix), gamma, beta, rm, rv, True, 1, 0.0,
                                                                    buf[0], buf[1], buf[2], buf[3]), 4 * e)
    run("bn_act_bwd_fused px%d C%d" % (npix, c), lambda: K.bn_act_bwd_fused(dy, None, z, dz, buf[0], buf[1], buf[2], buf[3], red, 1, 0.0), 6 * e)
    run("bn_act_bwd_fused px%d C%d +dy2" % (npix, c), lambda: K.bn_act_bwd_fused(dy, dy2, z, dz, buf[0], buf[1], buf[2], buf[3], red, 1, 0.0), 8 * e)
    def two_kernel():
        K.bn_act_bwd_reduce(dy, None, z, buf[0], buf[1], buf[2], buf[3], 1, 0.0, red[0])
        K.bn_act_bwd_apply(dy, None, z, dz, buf[0], buf[1], buf[2], buf[3], red[0], 1, 0.0)
    run("bn_act_bwd reduce+apply px%d C%d" % (npix, c), two_kernel, 6 * e)
