"""Domain discriminators on the B200 kernels — drop-in for the reference's ``model/discriminator.py``
(same three class names, ``(num_classes, ndf=64)`` constructors, ``[N, 1, h, w]`` outputs and
state_dict keys, including the 58-entry layout of the shipped ``GTA5_model/GTA5_10_D1.pth``).

* ``FCDiscriminator``: the four 4x4/stride-2 convolutions are tcgen05 implicit GEMMs with bias and
  LeakyReLU(0.2) in the epilogue; the 1-channel classifier is a warp-per-pixel reduction kernel.
* ``DepthWiseSep[BN]FCDiscriminator``: HBM-bound — depthwise 4x4/s2 kernels (+bias, +LeakyReLU or
  BatchNorm statistics), point-wise 1x1 convs with the reference's ``padding=1`` (the map grows by
  two pixels per point-wise layer, discriminator.py:36-45) as implicit GEMMs whose TMA
  out-of-bounds fill supplies the zero border.
Outputs are fp32; inputs are the bf16 channels-last probabilities produced by
``losses.upsample_softmax`` (any other NCHW tensor is converted at the boundary).
"""
import torch
import torch.nn as nn

from .. import kernels as K
from .. import ops
from ._glue import B200Module, bn_tuple, to_nhwc, zero_bordered_base

F32 = torch.float32
BF16 = torch.bfloat16
SLOPE = 0.2


def _pad_vec(v, n, fill=0.0):
    if v.numel() == n:
        return v.detach()
    out = torch.full((n,), fill, dtype=F32, device=v.device)
    out[: v.numel()].copy_(v.detach())
    return out


class _Classifier:
    """Shared forward/backward of the final Conv2d(512, 1, 4, 2, 1)."""

    @staticmethod
    def fwd(conv, a):
        n, h, w, _ = a.shape
        ho, wo = (h + 2 - 4) // 2 + 1, (w + 2 - 4) // 2 + 1
        out = torch.empty((n, 1, ho, wo), dtype=F32, device=a.device)
        K.classifier_fwd(a, conv.weight.detach(), conv.bias.detach(), out)
        return out

    @staticmethod
    def bwd(conv, a, dout, need_dw):
        dout = dout.contiguous().float()
        da = torch.empty(a.shape, dtype=BF16, device=a.device)
        K.classifier_dgrad(dout, conv.weight.detach(), da)
        grads = {}
        if need_dw:
            dw = ops.zeros_f32(conv.weight.shape, a.device)
            db = ops.zeros_f32((1,), a.device)
            K.classifier_wgrad(dout, a, dw, db)
            grads = {conv.weight: dw, conv.bias: db}
        return da, grads


class _DiscriminatorBase(B200Module):
    def _fwd_api(self, x):
        out, ctx = self._fwd(to_nhwc(x))
        ctx["in_c"] = x.shape[1]
        return out, ctx

    def _bwd_api(self, ctx, dout, need_dx=True, need_dw=True):
        dx, grads = self._bwd(ctx, dout, need_dx=need_dx, need_dw=need_dw)
        if dx is not None:
            dx = dx.permute(0, 3, 1, 2)[:, : ctx["in_c"]]
        return dx, grads


class FCDiscriminator(_DiscriminatorBase):
    """reference discriminator.py:4-28"""

    def __init__(self, num_classes, ndf=64):
        super(FCDiscriminator, self).__init__()
        self.conv1 = nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1)
        self.conv3 = nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1)
        self.conv4 = nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1)
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    # train.py asks losses.upsample_softmax for the zero-bordered layout when the discriminator says so
    wants_zero_bordered_input = True

    def _fwd_api(self, x):
        xpad = zero_bordered_base(x) if self.conv1.weight.shape[0] % 64 == 0 else None
        if xpad is None:
            return super(FCDiscriminator, self)._fwd_api(x)
        out, ctx = self._fwd(None, xpad=xpad)
        ctx["in_c"] = x.shape[1]
        return out, ctx

    def _fwd(self, x, xpad=None):
        ctxs = []
        cur = x
        for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
            if conv is self.conv1 and xpad is not None:
                cur, c = ops.conv_pairview_fwd(xpad, conv.weight, conv.bias.detach(), K.ACT_LEAKY, SLOPE)
            else:
                cur, c = ops.conv_bias_act_fwd(cur, conv.weight, conv.bias.detach(), 2, 1, K.ACT_LEAKY, SLOPE)
            ctxs.append(c)
        out = _Classifier.fwd(self.classifier, cur)
        return out, {"convs": ctxs, "a4": cur}

    def _bwd(self, ctx, dout, need_dx=True, need_dw=True):
        d, grads = _Classifier.bwd(self.classifier, ctx["a4"], dout, need_dw)
        convs = (self.conv1, self.conv2, self.conv3, self.conv4)
        ready = None   # (dz, dbias) of layer i when the layer above already applied its LeakyReLU backward
        for i in (3, 2, 1, 0):
            # layers 3..1: LeakyReLU backward + bias gradient of the layer below ride in this layer's
            # data-gradient epilogue (one pass over the big activation gradients instead of three)
            d, dw, db = ops.conv_bias_act_bwd(ctx["convs"][i], d, None, need_dx=(i > 0 or need_dx), need_dw=need_dw,
                                              dz_dbias=ready, producer=ctx["convs"][i - 1] if i > 0 else None)
            ready = d if isinstance(d, tuple) else None
            if need_dw:
                grads[convs[i].weight] = dw
                grads[convs[i].bias] = db
            if i == 3:
                ops.REDUCER.hook()   # conv4 + classifier = 3/4 of the parameters: exchange them early
        return d, grads


class _DepthWiseSepBase(_DiscriminatorBase):
    _use_bn = False

    def _build(self, num_classes, ndf):
        chans = [num_classes, ndf, ndf * 2, ndf * 4, ndf * 8]
        for i in range(4):
            setattr(self, "conv%d_d" % (i + 1), nn.Conv2d(chans[i], chans[i], kernel_size=4, stride=2,
                                                           padding=1, groups=chans[i]))
            if self._use_bn:
                setattr(self, "bn%d_d" % (i + 1), nn.BatchNorm2d(chans[i]))
            setattr(self, "conv%d_p" % (i + 1), nn.Conv2d(chans[i], chans[i + 1], kernel_size=1, padding=1))
            if self._use_bn:
                setattr(self, "bn%d_p" % (i + 1), nn.BatchNorm2d(chans[i + 1]))
        self.classifier = nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1)
        self.leaky_relu = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    # depthwise weights / BN vectors of the 19-channel first stage are zero-padded to the 32
    # channels of the probability buffer (padding channels stay exactly zero through the stage)
    def _dw_params(self, i, cpad):
        conv = getattr(self, "conv%d_d" % i)
        c = conv.weight.shape[0]
        if c == cpad:
            w, b = conv.weight.detach(), conv.bias.detach()
        else:
            w = ops.zeros_f32((cpad, 1, 4, 4), conv.weight.device)
            w[:c].copy_(conv.weight.detach())
            b = _pad_vec(conv.bias, cpad)
        bn = None
        if self._use_bn:
            m = getattr(self, "bn%d_d" % i)
            if c == cpad:
                bn = bn_tuple(m, self.training)
            else:
                g, be, rm, rv = bn_tuple(m, self.training)
                bn = (_pad_vec(g, cpad), _pad_vec(be, cpad), _pad_vec(rm, cpad), _pad_vec(rv, cpad, 1.0))
        return conv, w, b, bn

    def _fwd(self, x):
        ctxs = []
        cur = x
        n = x.shape[0]
        for i in (1, 2, 3, 4):
            cpad = K.round_up(cur.shape[3], 8)
            if cur.shape[3] != cpad:  # 19 -> 32: widen the view over the zero padding channels
                assert cur.stride(2) >= cpad
                cur = cur.as_strided((n, cur.shape[1], cur.shape[2], cpad), cur.stride())
            conv_d, w, b, bn = self._dw_params(i, cpad)
            a, cd = ops.dw_bn_fwd(cur, w.view(cpad, 1, 4, 4), bn, self.training, None, K.ACT_LEAKY, SLOPE, bias=b)
            if bn is not None and self.training and bn[2] is not getattr(self, "bn%d_d" % i).running_mean:
                m = getattr(self, "bn%d_d" % i)
                c = m.running_mean.numel()
                m.running_mean.copy_(bn[2][:c])
                m.running_var.copy_(bn[3][:c])
            conv_p = getattr(self, "conv%d_p" % i)
            if self._use_bn:
                m = getattr(self, "bn%d_p" % i)
                cout = conv_p.weight.shape[0]
                stats = ops.zeros_f32((2, cout), x.device) if self.training else None
                z = ops.conv_raw_fwd(a, conv_p.weight, 1, 1, stats=stats, bias=conv_p.bias.detach())
                p, cb = ops.bn_act_fwd(z, stats, bn_tuple(m, self.training), self.training, K.ACT_LEAKY, SLOPE)
                ctxs.append((cd, ("bn", cb, a)))
            else:
                p, cp = ops.conv_bias_act_fwd(a, conv_p.weight, conv_p.bias.detach(), 1, 1, K.ACT_LEAKY, SLOPE)
                ctxs.append((cd, ("act", cp, a)))
            cur = p
        out = _Classifier.fwd(self.classifier, cur)
        return out, {"stages": ctxs, "a4": cur}

    def _bwd(self, ctx, dout, need_dx=True, need_dw=True):
        d, grads = _Classifier.bwd(self.classifier, ctx["a4"], dout, need_dw)
        for i in (4, 3, 2, 1):
            cd, (kind, cp, a) = ctx["stages"][i - 1]
            conv_p = getattr(self, "conv%d_p" % i)
            conv_d = getattr(self, "conv%d_d" % i)
            if kind == "bn":
                dz, dg, db = ops.bn_act_bwd(cp, d)
                m = getattr(self, "bn%d_p" % i)
                if need_dw:
                    grads[m.weight], grads[m.bias] = dg, db
                    grads[conv_p.weight] = ops.conv_wgrad(dz, a, conv_p.weight, 1, 1)
                    st = ops.zeros_f32((2, dz.shape[3]), dz.device)
                    K.channel_stats(dz, st)
                    grads[conv_p.bias] = st[0]
                d = ops.conv_dgrad(dz, conv_p.weight, 1, 1, a.shape[1], a.shape[2])
            else:
                d, dw, db = ops.conv_bias_act_bwd(cp, d, None, need_dx=True, need_dw=need_dw)
                if need_dw:
                    grads[conv_p.weight], grads[conv_p.bias] = dw, db
            c = conv_d.weight.shape[0]
            if d.shape[3] != cd.x.shape[3]:
                d = d[..., : cd.x.shape[3]] if d.shape[3] > cd.x.shape[3] else d
            dx, dw, dbias, dg, db = ops.dw_bn_bwd(cd, d, None, need_dx=(i > 1 or need_dx))
            if need_dw:
                grads[conv_d.weight] = dw[:c].contiguous() if dw.shape[0] != c else dw
                grads[conv_d.bias] = dbias[:c].contiguous() if dbias.shape[0] != c else dbias
                if self._use_bn:
                    m = getattr(self, "bn%d_d" % i)
                    grads[m.weight] = dg[:c].contiguous() if dg.shape[0] != c else dg
                    grads[m.bias] = db[:c].contiguous() if db.shape[0] != c else db
            d = dx
        return d, grads


class DepthWiseSepFCDiscriminator(_DepthWiseSepBase):
    """reference discriminator.py:30-73"""
    _use_bn = False

    def __init__(self, num_classes, ndf=64):
        super(DepthWiseSepFCDiscriminator, self).__init__()
        self._build(num_classes, ndf)


class DepthWiseSepBNFCDiscriminator(_DepthWiseSepBase):
    """reference discriminator.py:75-134"""
    _use_bn = True

    def __init__(self, num_classes, ndf=64):
        super(DepthWiseSepBNFCDiscriminator, self).__init__()
        self._build(num_classes, ndf)
