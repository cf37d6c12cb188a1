"""Rebuilds the launches recorded in ``tuned_tiles.json`` from their shape keys.

A key is what ``kernels.conv_key`` / ``kernels.wgrad_key`` produce:

    conv  N Hin Win c<cin_pad> r<rows_pad> <f|d><R>x<S>s<stride>p<pad>[ f32][ st][ mk]
    wgrad N Ho  Wo  co<Cout> ci<Cin> <R>x<S>s<stride>[p<pad>]

(older wgrad keys carry no padding: 3x3 -> 1, 4x4 -> 1, 1x1 -> 0, except the point-wise convs of the
depthwise-separable discriminators, discriminator.py:36-45, whose padding=1 shows in the map size).
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TUNED_PATH = os.path.join(ROOT, "dasemanticsegmentationaml_b200", "tuned_tiles.json")

# output sizes of the padding=1 point-wise layers at 512x1024 (258x514 -> ... -> 35x67)
_PW_PAD1 = {(258, 514), (131, 259), (67, 131), (35, 67)}


def load_tuned():
    with open(TUNED_PATH) as f:
        return {k: int(v) for k, v in json.load(f).items()}


def conv_keys():
    return sorted(k for k in load_tuned() if k.startswith("conv "))


def wgrad_keys():
    return sorted(k for k in load_tuned() if k.startswith("wgrad "))


def parse_conv_key(key):
    m = re.match(r"conv (\d+) (\d+) (\d+) c64 r(\d+) pv([fd])4x4s2p1$", key)
    if m:   # pair view of the zero-bordered probability map (kernels.pairview_*_geometry)
        n, hin, win, rows = (int(m.group(i)) for i in range(1, 5))
        fwd = m.group(5) == "f"
        return dict(n=n, hin=hin, win=win, cin_pad=64, rows=rows, dgrad=not fwd, r=4, s=4, stride=2, pad=1, f32=False,
                    stats=False, pairview=(2 * hin - 2, win - 2) if fwd else (2 * hin, 2 * win))
    m = re.match(r"conv (\d+) (\d+) (\d+) c(\d+) r(\d+) ([fd])(\d+)x(\d+)s(\d+)p(\d+)( f32)?( st)?( mk)?$", key)
    assert m, key
    n, hin, win, cin_pad, rows = (int(m.group(i)) for i in range(1, 6))
    return dict(n=n, hin=hin, win=win, cin_pad=cin_pad, rows=rows, dgrad=m.group(6) == "d", r=int(m.group(7)),
                s=int(m.group(8)), stride=int(m.group(9)), pad=int(m.group(10)), f32=bool(m.group(11)),
                stats=bool(m.group(12)), mask=bool(m.group(13)))


def parse_wgrad_key(key):
    m = re.match(r"wgrad (\d+) (\d+) (\d+) co(\d+) ci64 8x1s1 v(\d+)$", key)
    if m:   # pair view: 8 taps over the [N, Ho + 1, 2 Wo + 2, 64] view
        n, ho, wo, cout = (int(m.group(i)) for i in range(1, 5))
        return dict(n=n, ho=ho, wo=wo, cout=cout, cin=64, r=8, s=1, stride=1, pad=0, hin=ho + 1, win=2 * wo + 2,
                    pairview=(2 * ho, 2 * wo))
    m = re.match(r"wgrad (\d+) (\d+) (\d+) co(\d+) ci(\d+) (\d+)x(\d+)s(\d+)(?:p(\d+))?$", key)
    assert m, key
    n, ho, wo, cout, cin, r, s, stride = (int(m.group(i)) for i in range(1, 9))
    if m.group(9) is not None:
        pad = int(m.group(9))
    elif r == 1:
        pad = 1 if (ho, wo) in _PW_PAD1 else 0
    else:
        pad = 1
    hin = (ho - 1) * stride + r - 2 * pad
    win = (wo - 1) * stride + s - 2 * pad
    if stride == 2 and r == 4:          # 4x4 s2 p1: Hin = 2 Ho exactly
        hin, win = 2 * ho, 2 * wo
    elif stride == 2 and r == 3:        # 3x3 s2 p1 on an even map: Hin = 2 Ho
        hin, win = 2 * ho, 2 * wo
    return dict(n=n, ho=ho, wo=wo, cout=cout, cin=cin, r=r, s=s, stride=stride, pad=pad, hin=hin, win=win)


def igemm_reference(x, filt, geom, n_out_rows=None):
    """fp32 restatement of ONE b200_conv_igemm launch straight from its tap tables (independent of
    which convolution the tables encode):

        out[n, h*os + oa, w*os + ob, co] = sum_{tap, ci} x[n, h*is + dh, w*is + dw, ci] * filt[co, slab, ci]

    x: [N, Hin, Win, C] (C <= cin_pad), filt: packed [rows, n_slabs, cin_pad]; out-of-range input
    pixels read as zero.  Used by scripts/tune_conv.py to reject tile configurations that produce
    wrong results, and by the tests as a second opinion next to F.conv2d."""
    import torch
    import torch.nn.functional as F
    n, hin, win, c = x.shape
    rows = filt.shape[0] if n_out_rows is None else n_out_rows
    xf = x.float()
    wf = filt.float()[:rows, :, :c]
    out = torch.zeros((n, geom.Hout, geom.Wout, rows), dtype=torch.float32, device=x.device)
    ist, ost = geom.in_stride, geom.out_stride
    for cl in geom.classes:
        ho, wo = cl["Ho"], cl["Wo"]
        acc = torch.zeros((n, ho, wo, rows), dtype=torch.float32, device=x.device)
        for dh, dw, slab in cl["taps"]:
            # input rows h*ist + dh for h in [0, ho): pad so that every index is in range
            top = max(0, -dh)
            left = max(0, -dw)
            bottom = max(0, (ho - 1) * ist + dh - (hin - 1))
            right = max(0, (wo - 1) * ist + dw - (win - 1))
            xp = F.pad(xf, (0, 0, left, right, top, bottom))
            sub = xp[:, dh + top: dh + top + (ho - 1) * ist + 1: ist, dw + left: dw + left + (wo - 1) * ist + 1: ist]
            acc += torch.einsum("nhwc,oc->nhwo", sub, wf[:, slab])
        out[:, cl["oa"]::ost, cl["ob"]::ost][:, :ho, :wo] = acc
    return out


def wgrad_reference(dz, x, taps, stride):
    """fp32 restatement of ONE b200_conv_wgrad launch from its tap list:
        dW[co, ci, rs] = sum_{n, h, w} dz[n, h, w, co] * x[n, h*stride + dh, w*stride + dw, ci]   for (dh, dw, rs) in taps
    (out-of-range input pixels read as zero)."""
    import torch
    import torch.nn.functional as F
    n, ho, wo, cout = dz.shape
    _, hin, win, cin = x.shape
    dzf, xf = dz.float(), x.float()
    out = torch.zeros((cout, cin, len(taps)), dtype=torch.float32, device=x.device)
    for dh, dw, rs in taps:
        top, left = max(0, -dh), max(0, -dw)
        bottom = max(0, (ho - 1) * stride + dh - (hin - 1))
        right = max(0, (wo - 1) * stride + dw - (win - 1))
        xp = F.pad(xf, (0, 0, left, right, top, bottom))
        sub = xp[:, dh + top: dh + top + (ho - 1) * stride + 1: stride, dw + left: dw + left + (wo - 1) * stride + 1: stride]
        out[:, :, rs] = torch.einsum("nhwo,nhwc->oc", dzf, sub)
    return out
