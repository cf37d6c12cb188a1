"""Parity of EXACTLY the kernel configurations the benchmark runs: every entry of
``tuned_tiles.json`` (N = 8 shapes; persistent kernel, resident filter, mt = 2 sub-tiles, tuned
weight-gradient splits / K-pixel boxes) is rebuilt from its shape key, launched with the tuned word
AND with the launcher's heuristic (tune = 0), and both are compared with torch's fp32 convolution
on the same bf16-rounded operands:

    outputs  rel-L2 <= 4e-3 (bf16, one rounding) / 1e-4 (fp32);   BN statistics 1e-3;   dW 1e-3

(bf16 products are exact in fp32, so only the accumulation order and the output rounding differ.)
The tile shape does not change the order in which a single output element accumulates its taps and
channels, so tuned and heuristic outputs must also be bit-identical (the check formerly done by
hand with scripts/ab_persist.py).  Reference: model/stdcnet.py:13-15, model/discriminator.py:17-26.
"""
import zlib

import pytest
import torch
import torch.nn.functional as F

from tests import tuned_cases as TC

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("key", TC.conv_keys())
def test_tuned_conv_entry(cuda_lib, key):
    from dasemanticsegmentationaml_b200 import kernels as K
    c = TC.parse_conv_key(key)
    tune = TC.load_tuned()[key]
    n, hin, win, cin_pad, rows = c["n"], c["hin"], c["win"], c["cin_pad"], c["rows"]
    r, s, stride, pad = c["r"], c["s"], c["stride"], c["pad"]
    g = _gen(zlib.crc32(key.encode()) % 1000)
    # the input is a channel slice of a wider buffer (concat-free layout) on some shapes
    in_ld = cin_pad + (32 if (hin * win) % 3 == 0 else 0)
    xbuf = torch.randn(n, hin, win, in_ld, device="cuda", generator=g).to(BF)
    x = xbuf[..., :cin_pad]
    if c.get("pairview"):
        # the dense discriminator's first layer over column pairs of the zero-bordered map: checked against the
        # tap-table restatement (tests/test_modules_gpu.py::test_dense_discriminator_pair_view ties the pair view
        # itself to Conv2d(19, 64, 4, 2, 1))
        geom = (K.pairview_dgrad_geometry if c["dgrad"] else K.pairview_fwd_geometry)(*c["pairview"])
        filt = (torch.randn(rows, 8, 64, device="cuda", generator=g) / 512 ** 0.5).to(BF)
        ref = TC.igemm_reference(x, filt, geom)
    elif not c["dgrad"]:
        cout, cin = rows, cin_pad
        wgt = (torch.randn(cout, cin, r, s, device="cuda", generator=g) / (cin * r * s) ** 0.5).to(BF).float()
        filt = K.pack_filter(wgt, transpose=False)
        geom = K.fwd_geometry(hin, win, r, s, stride, pad)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), wgt, stride=stride, padding=pad).permute(0, 2, 3, 1)
    else:
        # data gradient: x is dz [N, Ho, Wo, Cout]; the filter is packed transposed (rows = Cin)
        cout, cin = cin_pad, rows
        dh = 2 * hin if stride == 2 else (hin - 1) + r - 2 * pad
        dw_ = 2 * win if stride == 2 else (win - 1) + s - 2 * pad
        wgt = (torch.randn(cout, cin, r, s, device="cuda", generator=g) / (cout * r * s) ** 0.5).to(BF).float()
        filt = K.pack_filter(wgt, transpose=True)
        geom = K.dgrad_geometry(dh, dw_, r, s, stride, pad)
        ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wgt, stride=stride, padding=pad,
                                 output_padding=(dh - ((hin - 1) * stride + r - 2 * pad),
                                                 dw_ - ((win - 1) * stride + s - 2 * pad))).permute(0, 2, 3, 1)
    assert filt.shape[0] == rows and filt.shape[2] == cin_pad
    assert tuple(ref.shape[1:3]) == (geom.Hout, geom.Wout)
    odt = torch.float32 if c["f32"] else BF
    # discriminator forward layers carry bias + LeakyReLU in the epilogue (discriminator.py:17-25)
    with_bias = (not c["dgrad"]) and (r == 4 or (r == 1 and pad == 1))
    bias = torch.randn(rows, device="cuda", generator=g) if with_bias else None
    if with_bias:
        ref = F.leaky_relu(ref + bias, 0.2)
    # their data gradients carry the producer's LeakyReLU backward + bias gradient
    # (a " mk" key IS that launch; older keys without the suffix are exercised both ways)
    with_mask = (c.get("mask") or (c["dgrad"] and (r == 4 or (r == 1 and pad == 1)))) and not c["stats"] and rows % 16 == 0 \
        and not c.get("pairview")
    mask = None
    if with_mask:
        mask = torch.randn(n, geom.Hout, geom.Wout, rows, device="cuda", generator=g).to(BF)
        ref = ref * torch.where(mask.float() > 0, 1.0, 0.2)
    outs = []
    for word in (tune, 0):
        out_ld = rows + (16 if rows % 32 == 0 and (hin + win) % 2 == 0 else 0)
        obuf = torch.full((n, geom.Hout, geom.Wout, out_ld), 7.0, device="cuda", dtype=odt)
        out = obuf[..., :rows]
        want_stats = c["stats"] or with_mask
        stats = torch.zeros(2, rows, device="cuda") if want_stats else None
        K.conv_igemm(x, filt, out, geom, bias=bias, act=2 if with_bias else 0, slope=0.2, stats=stats, bn_tile=word,
                     mask=mask, mask_slope=0.2, stats_sum_only=with_mask)
        torch.cuda.synchronize()
        tol = 1e-4 if c["f32"] else 4e-3
        err = rel_l2(out, ref)
        assert err < tol, (key, word, err)
        if out_ld > rows:
            assert (obuf[..., rows:] == 7.0).all(), "wrote outside the channel slice"
        if want_stats:
            flat = (out if not c["f32"] else ref).float().reshape(-1, rows)   # statistics of the stored values
            assert rel_l2(stats[0], flat.sum(0)) < 1e-3, (key, word)
            if not with_mask:
                assert rel_l2(stats[1], (flat * flat).sum(0)) < 1e-3, (key, word)
        outs.append(out.clone())
        del obuf
    # (the input-halo mode -- tune bit 23, or the heuristic on 3x3-like stride-1 taps -- accumulates channel
    #  blocks outer / taps inner, i.e. in another fp32 order: both are within tolerance, not bit-identical)
    if rel_l2(outs[0], outs[1]) > 0:
        assert rel_l2(outs[0], outs[1]) < 4e-3, "tuned and heuristic tiles disagree: %s" % key


@pytest.mark.parametrize("key", TC.wgrad_keys())
def test_tuned_wgrad_entry(cuda_lib, key):
    from dasemanticsegmentationaml_b200 import kernels as K
    c = TC.parse_wgrad_key(key)
    tune = TC.load_tuned()[key]
    n, ho, wo, cout, cin = c["n"], c["ho"], c["wo"], c["cout"], c["cin"]
    r, s, stride, pad, hin, win = c["r"], c["s"], c["stride"], c["pad"], c["hin"], c["win"]
    g = _gen(zlib.crc32(key.encode()) % 1000)
    if c.get("pairview"):
        taps = K.pairview_fwd_geometry(*c["pairview"]).classes[0]["taps"]
        x = torch.randn(n, hin, win, 64, device="cuda", generator=g).to(BF)
        dz = (torch.randn(n, ho, wo, cout, device="cuda", generator=g) * 0.1).to(BF)
        ref = TC.wgrad_reference(dz, x, taps, 1)
        for word in (tune, 0):
            sc = torch.zeros(8, cout, 64, device="cuda")
            K.conv_wgrad(dz, x, sc.view(cout, 64, 8, 1), 8, 1, 1, 0, tune=word, scratch=sc, taps=taps)
            torch.cuda.synchronize()
            assert rel_l2(sc.permute(1, 2, 0), ref) < 1e-3, (key, word)
        return
    x_c = K.round_up(cin, 8)
    xbuf = torch.zeros(n, hin, win, K.round_up(cin, 32), device="cuda", dtype=BF)
    xbuf[..., :cin] = torch.randn(n, hin, win, cin, device="cuda", generator=g).to(BF)
    x = xbuf[..., :x_c]
    dz_c = K.round_up(cout, 16)
    dzbuf = torch.zeros(n, ho, wo, dz_c, device="cuda", dtype=BF)
    dzbuf[..., :cout] = (torch.randn(n, ho, wo, cout, device="cuda", generator=g) * 0.1).to(BF)
    wf = torch.zeros(cout, cin, r, s, device="cuda", requires_grad=True)
    y = F.conv2d(xbuf[..., :cin].float().permute(0, 3, 1, 2), wf, stride=stride, padding=pad)
    assert tuple(y.shape[2:]) == (ho, wo), (key, y.shape)
    y.backward(dzbuf[..., :cout].float().permute(0, 3, 1, 2))
    for word in (tune, 0):
        dw = torch.zeros(cout, cin, r, s, device="cuda")
        K.conv_wgrad(dzbuf, x, dw, r, s, stride, pad, tune=word)
        torch.cuda.synchronize()
        err = rel_l2(dw, wf.grad)
        assert err < 1e-3, (key, word, err)
        if r * s > 1 and cin > 32:     # tap-major scratch path with the same split
            ci_pad = K.round_up(cin, 16)
            sc = torch.zeros(r * s, cout, ci_pad, device="cuda")
            dws = torch.zeros(cout, cin, r, s, device="cuda")
            K.conv_wgrad(dzbuf, x, dws, r, s, stride, pad, tune=word, scratch=sc)
            K.wgrad_unscratch(sc, dws, torch.tensor([[0, 0, cout, cin, r * s, ci_pad]], dtype=torch.int64).cuda(), 1)
            torch.cuda.synchronize()
            assert rel_l2(dws, wf.grad) < 1e-3, (key, word, "scratch")
