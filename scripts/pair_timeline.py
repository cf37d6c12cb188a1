"""Phase timeline of CTA pair 0 of the CTA-pair conv kernel (b200_debug_timeline) for a few launch shapes.

    python scripts/pair_timeline.py [name ...]

Prints, per shape: total kernel time by CUDA events inside a graph, set-up / first-operand / exit offsets and
the per-tile MMA and epilogue windows in microseconds relative to kernel entry."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dasemanticsegmentationaml_b200 import build, kernels as K
from dasemanticsegmentationaml_b200._lib import call, ptr

build.build()
dev = torch.device("cuda", 0)
BF = torch.bfloat16
NT = 30


def shapes():
    out = {}
    # name -> (x shape, filter [rows, slabs, cin_pad], geometry, stats)
    out["pvf"] = ((8, 257, 1026, 64), (64, 8, 64), K.pairview_fwd_geometry(512, 1024), False)
    out["pvd"] = ((8, 256, 512, 64), (64, 8, 64), K.pairview_dgrad_geometry(512, 1024), False)
    out["d2dgrad"] = ((8, 128, 256, 128), (64, 16, 128), K.dgrad_geometry(256, 512, 4, 4, 2, 1), False)
    out["s16"] = ((8, 32, 64, 128), (64, 9, 128), K.fwd_geometry(32, 64, 3, 3, 1, 1), True)      # M16384 N64 K128x9
    out["s32"] = ((8, 16, 32, 512), (256, 9, 512), K.fwd_geometry(16, 32, 3, 3, 1, 1), True)      # M4096 N256 K512x9
    out["s8"] = ((8, 64, 128, 128), (64, 9, 128), K.fwd_geometry(64, 128, 3, 3, 1, 1), True)      # M65536 N64 K128x9
    out["big"] = ((8, 64, 128, 256), (256, 9, 256), K.fwd_geometry(64, 128, 3, 3, 1, 1), True)    # M65536 N256 K256x9
    out["d2fwd"] = ((8, 256, 512, 64), (128, 16, 64), K.fwd_geometry(256, 512, 4, 4, 2, 1), False)
    out["stem"] = ((8, 512, 1024, 32), (32, 1, 32), K.fwd_geometry(512, 1024, 1, 1, 1, 0), True)   # M1048576 N32 K32x1
    out["c1x1"] = ((8, 128, 256, 64), (128, 1, 64), K.fwd_geometry(128, 256, 1, 1, 1, 0), True)    # M262144 N128 K64x1
    out["f1x1"] = ((8, 64, 128, 384), (256, 1, 384), K.fwd_geometry(64, 128, 1, 1, 1, 0), True)    # M65536 N256 K384x1
    return out


def main():
    knob = int(os.environ.get("KNOB", "0"))
    call("b200_debug_knob", K.c_int(knob))
    names = sys.argv[1:] or list(shapes())
    tunes = {}
    for i, a in enumerate(list(names)):
        if "=" in a:
            n, t = a.split("=")
            tunes[i] = int(t, 0)
            names[i] = n
    buf = torch.zeros(8 + 4 * NT, dtype=torch.int64, device=dev)
    for idx, name in enumerate(names):
        xs, fs, geom, st = shapes()[name]
        x = torch.randn(xs, device=dev).to(BF)
        filt = (torch.randn(fs, device=dev) * 0.05).to(BF)
        out = torch.empty((xs[0], geom.Hout, geom.Wout, fs[0]), device=dev, dtype=BF)
        stats = torch.zeros(2, fs[0], device=dev) if st else None
        tune = tunes.get(idx)
        if tune is None:
            key = K.conv_key(xs[0], xs[1], xs[2], fs[2], fs[0], geom, 0, st)
            tune = K.TUNED.get(key, 0)

        def run():
            K.conv_igemm(x, filt, out, geom, stats=stats, bn_tile=tune)
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        buf.zero_()
        call("b200_debug_timeline", ptr(buf))
        run()
        torch.cuda.synchronize()
        call("b200_debug_timeline", None)
        t = buf.cpu().tolist()
        t0 = t[0]
        if t0 == 0:
            print("%-8s tune=%#x  %.1f us/launch in a graph  (not the CTA-pair kernel: no timeline)" % (name, tune, us))
            continue
        rel = lambda v: (v - t0) / 1e3 if v else float("nan")
        print("%-8s tune=%#x  %.1f us/launch in a graph; pair 0: set-up %.2f  first operands %.2f  exit %.2f us" %
              (name, tune, us, rel(t[1]), rel(t[2]), rel(t[3])))
        if t[4]:
            e1 = t[8 + 4 + 2]
            print("  tile 1, warp 2 (us after the accumulator was full): first TMEM batch in registers %.2f, first batch done %.2f, "
                  "tile done %.2f" % tuple((v - e1) / 1e3 for v in t[4:7]))
        rows = []
        for lt in range(NT):
            a = t[8 + 4 * lt: 12 + 4 * lt]
            if a[0] == 0:
                break
            rows.append("  tile %2d: mma %.2f-%.2f  accumulator held by the epilogue %.2f-%.2f" % (lt, rel(a[0]), rel(a[1]), rel(a[2]), rel(a[3])))
        print("\n".join(rows[:6] + (["  ..."] if len(rows) > 8 else []) + rows[-2:] if len(rows) > 8 else rows))


main()
