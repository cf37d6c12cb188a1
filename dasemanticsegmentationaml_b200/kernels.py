"""Thin tensor-level wrappers over the C-ABI entry points of libb200seg.so.

Activations are NHWC bf16 tensors ``[N, H, W, C]`` whose innermost stride is 1; a channel slice of a
wider buffer (``buf[..., a:b]``) is passed as-is — the kernels take the pixel stride from
``stride(2)``, which is how concatenations are built without a copy.
Every function launches on the current torch stream and returns immediately.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, c_int64, check, int_array, lib, ptr, stream


def round_up(v, m):
    return (v + m - 1) // m * m


def _nhwc_meta(t):
    assert t.dim() == 4 and t.stride(3) == 1, "expected an NHWC tensor with unit channel stride"
    n, h, w, c = t.shape
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "NHWC view must be pixel-dense"
    return n, h, w, c, ld


# ------------------------------------------------------------------ filter packing
def pack_filter(weight, transpose=False):
    """fp32 ``[Cout, Cin, R, S]`` -> bf16 ``[rows_pad, R*S, inner_pad]`` K-major GEMM operand.

    ``transpose=False``: rows = Cout (forward).  ``transpose=True``: rows = Cin (data gradient).
    rows are padded to a multiple of 16, the inner (reduction) channel count to 32 or 64.
    """
    cout, cin, r, s = weight.shape
    rows, inner = (cin, cout) if transpose else (cout, cin)
    rows_pad = round_up(rows, 16)
    inner_pad = round_up(inner, 32) if inner <= 32 else round_up(inner, 64)
    out = torch.empty((rows_pad, r * s, inner_pad), dtype=torch.bfloat16, device=weight.device)
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    check(lib().b200_pack_filter(ptr(w), ptr(out), c_int(cout), c_int(cin), c_int(r * s),
                                 c_int(rows_pad), c_int(inner_pad), c_int(1 if transpose else 0),
                                 stream()), "b200_pack_filter")
    return out


# ------------------------------------------------------------------ geometry
def fwd_geometry(hin, win, r, s, stride, pad):
    ho = (hin + 2 * pad - r) // stride + 1
    wo = (win + 2 * pad - s) // stride + 1
    taps = [(i - pad, j - pad, i * s + j) for i in range(r) for j in range(s)]
    return dict(classes=[dict(Ho=ho, Wo=wo, oa=0, ob=0, taps=taps)], in_stride=stride, out_stride=1,
                Hout=ho, Wout=wo)


def dgrad_geometry(hin, win, r, s, stride, pad):
    """Geometry of dx[N,hin,win,Cin] = conv_transpose(dz); dz is the kernel's *input*."""
    if stride == 1:
        taps = [(pad - i, pad - j, i * s + j) for i in range(r) for j in range(s)]
        return dict(classes=[dict(Ho=hin, Wo=win, oa=0, ob=0, taps=taps)], in_stride=1, out_stride=1,
                    Hout=hin, Wout=win)
    assert stride == 2
    classes = []
    for a in range(2):
        for b in range(2):
            taps = [((a + pad - i) // 2, (b + pad - j) // 2, i * s + j)
                    for i in range(r) if (a + pad - i) % 2 == 0
                    for j in range(s) if (b + pad - j) % 2 == 0]
            hc, wc = (hin - a + 1) // 2, (win - b + 1) // 2
            if not taps:
                raise ValueError("strided dgrad class without taps (1x1 stride-2 conv?)")
            if hc > 0 and wc > 0:
                classes.append(dict(Ho=hc, Wo=wc, oa=a, ob=b, taps=taps))
    return dict(classes=classes, in_stride=1, out_stride=2, Hout=hin, Wout=win)


def conv_igemm(x, filt, out, geom, bias=None, act=0, slope=0.0, stats=None, bn_tile=0):
    """out[N,Hout,Wout,rows_pad] = implicit-GEMM conv of x with a packed filter (see pack_filter)."""
    n, hin, win, cin, in_ld = _nhwc_meta(x)
    n2, hout, wout, cout_view, out_ld = _nhwc_meta(out)
    rows_pad, n_slabs, cin_pad = filt.shape
    assert n2 == n and hout == geom["Hout"] and wout == geom["Wout"], "output shape mismatch"
    assert cout_view == rows_pad, "output view must expose exactly the packed filter's rows"
    assert x.dtype == torch.bfloat16 and filt.dtype == torch.bfloat16
    out_f32 = 1 if out.dtype == torch.float32 else 0
    if not out_f32:
        assert out.dtype == torch.bfloat16
    classes = geom["classes"]
    tmax = max(len(c["taps"]) for c in classes)
    flat = []
    for c in classes:
        for t in range(tmax):
            flat += list(c["taps"][t]) if t < len(c["taps"]) else [0, 0, 0]
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.dim() == 2 and stats.shape[0] == 2
    check(lib().b200_conv_igemm(
        ptr(x), c_int(in_ld), c_int(0), c_int(cin), c_int(n), c_int(hin), c_int(win),
        ptr(filt), c_int(rows_pad), c_int(cin_pad), c_int(n_slabs),
        ptr(out), c_int(out_ld), c_int(0), c_int(hout), c_int(wout), c_int(out_f32),
        c_int(len(classes)), int_array([c["Ho"] for c in classes]), int_array([c["Wo"] for c in classes]),
        int_array([c["oa"] for c in classes]), int_array([c["ob"] for c in classes]),
        int_array([len(c["taps"]) for c in classes]), int_array(flat), c_int(tmax),
        c_int(geom["in_stride"]), c_int(geom["out_stride"]), ptr(bias), c_int(act), c_float(slope),
        ptr(stats), c_int(0 if stats is None else stats.shape[1]), c_int(bn_tile), stream()),
        "b200_conv_igemm")
    return out


def conv_wgrad(dz, x, dw, r, s, stride, pad):
    """dw[Cout,Cin,R,S] (fp32) += sum_pixels dz (x) x ; dz [N,Ho,Wo,>=Cout], x [N,Hin,Win,>=Cin]."""
    n, ho, wo, _, dz_ld = _nhwc_meta(dz)
    n2, hin, win, _, x_ld = _nhwc_meta(x)
    cout, cin = dw.shape[0], dw.shape[1]
    assert n2 == n and dw.dtype == torch.float32 and dw.is_contiguous()
    taps = []
    for i in range(r):
        for j in range(s):
            taps += [i - pad, j - pad, i * s + j]
    check(lib().b200_conv_wgrad(
        ptr(dz), c_int(dz_ld), c_int(0), c_int(cout), c_int(n), c_int(ho), c_int(wo),
        ptr(x), c_int(x_ld), c_int(0), c_int(cin), c_int(hin), c_int(win),
        c_int(r * s), int_array(taps), c_int(r * s), c_int(stride), ptr(dw), stream()),
        "b200_conv_wgrad")
    return dw


# ------------------------------------------------------------------ metrics
def fast_hist_accumulate(label, pred, n, hist, bad):
    """hist (int64 [n*n], device) += bincount(n*label+pred) over label in [0,n)."""
    assert label.numel() == pred.numel()
    label = label.contiguous()
    pred = pred.contiguous()
    check(lib().b200_fast_hist(ptr(label), c_int(label.element_size()), ptr(pred),
                               c_int(pred.element_size()), c_int64(label.numel()), c_int(n),
                               ptr(hist), ptr(bad), stream()), "b200_fast_hist")
    return hist
