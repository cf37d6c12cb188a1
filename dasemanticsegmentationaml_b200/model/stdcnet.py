"""STDCNet813 backbone on the B200 kernels — drop-in for the reference's ``model/stdcnet.py``
(same class names, constructor signatures, forward returns and state_dict keys).

Every ``ConvX`` is one tcgen05 implicit-GEMM launch (BatchNorm statistics reduced in its epilogue)
plus one vectorised normalise+ReLU pass that writes straight into the channel slice of the
bottleneck's output, so ``torch.cat`` (stdcnet.py:112) never runs; the stride-2 blocks' depthwise
conv and AvgPool2d skip (stdcnet.py:73-78) share one fused kernel.
"""
import math

import torch
import torch.nn as nn
from torch.nn import init

from .. import kernels as K
from .. import ops
from ._glue import B200Module, bn_tuple


class ConvX(B200Module):
    """conv(k, stride, pad=k//2, bias=False) -> BatchNorm2d -> ReLU   (reference stdcnet.py:6-15)."""

    def __init__(self, in_planes, out_planes, kernel=3, stride=1):
        super(ConvX, self).__init__()
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel, stride=stride,
                              padding=kernel // 2, bias=False)
        self.bn = nn.BatchNorm2d(out_planes)
        self.relu = nn.ReLU(inplace=True)

    def _fwd(self, x, out=None):
        return ops.conv_bn_act_fwd(x, self.conv.weight, bn_tuple(self.bn, self.training),
                                   self.conv.stride[0], self.conv.padding[0], self.training,
                                   K.ACT_RELU, out=out)

    def _bwd(self, ctx, dy1, dy2=None, need_dx=True):
        dx, dw, dg, db = ops.conv_bn_act_bwd(ctx, dy1, dy2, need_dx)
        return dx, {self.conv.weight: dw, self.bn.weight: dg, self.bn.bias: db}

    def _fwd_api(self, x):
        if x.shape[1] == 3:  # the stem reads the fp32 NCHW image directly
            a, ctx = self._fwd(x.float().contiguous())
            return a.permute(0, 3, 1, 2), ctx
        return super(ConvX, self)._fwd_api(x)


def _conv_plan(in_planes, out_planes, block_num, stride):
    """(cin, cout, kernel, stride) of conv_list, as laid out at reference stdcnet.py:36-47 / 82-92."""
    plan = []
    for idx in range(block_num):
        if idx == 0:
            plan.append((in_planes, out_planes // 2, 1, 1))
        elif idx == 1 and block_num == 2:
            plan.append((out_planes // 2, out_planes // 2, 3, stride))
        elif idx == 1 and block_num > 2:
            plan.append((out_planes // 2, out_planes // 4, 3, stride))
        elif idx < block_num - 1:
            plan.append((out_planes // int(math.pow(2, idx)), out_planes // int(math.pow(2, idx + 1)), 3, 1))
        else:
            plan.append((out_planes // int(math.pow(2, idx)), out_planes // int(math.pow(2, idx)), 3, 1))
    return plan


class CatBottleneck(B200Module):
    """STDC module, concatenation variant (reference stdcnet.py:66-113)."""

    def __init__(self, in_planes, out_planes, block_num=3, stride=1):
        super(CatBottleneck, self).__init__()
        assert block_num > 1, "block number should be larger than 1."
        self.conv_list = nn.ModuleList()
        self.stride = stride
        if stride == 2:
            self.avd_layer = nn.Sequential(
                nn.Conv2d(out_planes // 2, out_planes // 2, kernel_size=3, stride=2, padding=1,
                          groups=out_planes // 2, bias=False),
                nn.BatchNorm2d(out_planes // 2),
            )
            self.skip = nn.AvgPool2d(kernel_size=3, stride=2, padding=1)
        for cin, cout, k, s in _conv_plan(in_planes, out_planes, block_num, 1):
            self.conv_list.append(ConvX(cin, cout, kernel=k, stride=s))
        self.out_planes = out_planes

    def _fwd(self, x, out=None):
        n, h, w, _ = x.shape
        half = self.conv_list[0].conv.out_channels
        ctx = {}
        if self.stride == 2:
            a1, ctx["c0"] = self.conv_list[0]._fwd(x)
            ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            y = out if out is not None else ops.empty_act(n, ho, wo, self.out_planes, x.device)
            cur, ctx["dw"] = ops.dw_bn_fwd(a1, self.avd_layer[0].weight,
                                           bn_tuple(self.avd_layer[1], self.training), self.training,
                                           pool_out=y[..., :half])
        else:
            y = out if out is not None else ops.empty_act(n, h, w, self.out_planes, x.device)
            cur, ctx["c0"] = self.conv_list[0]._fwd(x, out=y[..., :half])
        off = half
        ctx["tail"] = []
        for conv in list(self.conv_list)[1:]:
            c = conv.conv.out_channels
            cur, cc = conv._fwd(cur, out=y[..., off:off + c])
            ctx["tail"].append((cc, off, c))
            off += c
        return y, ctx

    def _bwd(self, ctx, dy, need_dx=True):
        grads = {}
        half = self.conv_list[0].conv.out_channels
        dnext = None
        tail = list(self.conv_list)[1:]
        for conv, (cc, off, c) in zip(reversed(tail), reversed(ctx["tail"])):
            dnext, g = conv._bwd(cc, dy[..., off:off + c], dnext, True)
            grads.update(g)
        if self.stride == 2:
            da1, dw, _, dg, db = ops.dw_bn_bwd(ctx["dw"], dnext, dpool=dy[..., :half])
            grads.update({self.avd_layer[0].weight: dw, self.avd_layer[1].weight: dg,
                          self.avd_layer[1].bias: db})
            dx, g = self.conv_list[0]._bwd(ctx["c0"], da1, None, need_dx)
        else:
            dx, g = self.conv_list[0]._bwd(ctx["c0"], dy[..., :half], dnext, need_dx)
        grads.update(g)
        return dx, grads


class AddBottleneck(B200Module):
    """STDC module, residual-add variant (reference stdcnet.py:17-64).  The reference never
    instantiates it (STDCNet813 defaults to type="cat", ContextPath passes no type); reachable through
    ``STDCNet813(type="add")``.  Same kernels as CatBottleneck plus one fused add for the residual."""

    def __init__(self, in_planes, out_planes, block_num=3, stride=1):
        super(AddBottleneck, self).__init__()
        assert block_num > 1, "block number should be larger than 1."
        self.conv_list = nn.ModuleList()
        self.stride = stride
        if stride == 2:
            self.avd_layer = nn.Sequential(
                nn.Conv2d(out_planes // 2, out_planes // 2, kernel_size=3, stride=2, padding=1,
                          groups=out_planes // 2, bias=False),
                nn.BatchNorm2d(out_planes // 2),
            )
            self.skip = nn.Sequential(
                nn.Conv2d(in_planes, in_planes, kernel_size=3, stride=2, padding=1, groups=in_planes, bias=False),
                nn.BatchNorm2d(in_planes),
                nn.Conv2d(in_planes, out_planes, kernel_size=1, bias=False),
                nn.BatchNorm2d(out_planes),
            )
        for cin, cout, k, s in _conv_plan(in_planes, out_planes, block_num, 1):
            self.conv_list.append(ConvX(cin, cout, kernel=k, stride=s))
        self.out_planes = out_planes

    def _fwd(self, x, out=None):
        """cat(conv_list outputs) + skip(x)   (reference stdcnet.py:49-64)."""
        n, h, w, _ = x.shape
        half = self.conv_list[0].conv.out_channels
        ctx = {}
        if self.stride == 2:
            a1, ctx["c0"] = self.conv_list[0]._fwd(x)
            ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            y = ops.empty_act(n, ho, wo, self.out_planes, x.device)
            cur, ctx["dw"] = ops.dw_bn_fwd(a1, self.avd_layer[0].weight, bn_tuple(self.avd_layer[1], self.training),
                                           self.training, out=y[..., :half])
            # skip = dw3x3 s2 -> BN -> 1x1 conv -> BN (no activation anywhere, stdcnet.py:30-35)
            s1, ctx["skip_dw"] = ops.dw_bn_fwd(x, self.skip[0].weight, bn_tuple(self.skip[1], self.training), self.training)
            res, ctx["skip_pw"] = ops.conv_bn_act_fwd(s1, self.skip[2].weight, bn_tuple(self.skip[3], self.training), 1, 0,
                                                      self.training, act=K.ACT_NONE)
        else:
            if x.shape[3] != self.out_planes:
                raise ValueError("AddBottleneck(stride=1) adds its input to its output: in_planes must equal out_planes")
            y = ops.empty_act(n, h, w, self.out_planes, x.device)
            cur, ctx["c0"] = self.conv_list[0]._fwd(x, out=y[..., :half])
            res = x
        off = half
        ctx["tail"] = []
        for conv in list(self.conv_list)[1:]:
            c = conv.conv.out_channels
            cur, cc = conv._fwd(cur, out=y[..., off:off + c])
            ctx["tail"].append((cc, off, c))
            off += c
        return ops.add_acts(y, res, out), ctx

    def _bwd(self, ctx, dy, need_dx=True):
        grads = {}
        half = self.conv_list[0].conv.out_channels
        dnext = None
        tail = list(self.conv_list)[1:]
        for conv, (cc, off, c) in zip(reversed(tail), reversed(ctx["tail"])):
            dnext, g = conv._bwd(cc, dy[..., off:off + c], dnext, True)
            grads.update(g)
        if self.stride == 2:
            da1, dw, _, dg, db = ops.dw_bn_bwd(ctx["dw"], dy[..., :half], dy2=dnext)
            grads.update({self.avd_layer[0].weight: dw, self.avd_layer[1].weight: dg, self.avd_layer[1].bias: db})
            dx, g = self.conv_list[0]._bwd(ctx["c0"], da1, None, need_dx)
            grads.update(g)
            ds1, dwp, dgp, dbp = ops.conv_bn_act_bwd(ctx["skip_pw"], dy, None, True)
            grads.update({self.skip[2].weight: dwp, self.skip[3].weight: dgp, self.skip[3].bias: dbp})
            dxs, dws, _, dgs, dbs = ops.dw_bn_bwd(ctx["skip_dw"], ds1, need_dx=need_dx)
            grads.update({self.skip[0].weight: dws, self.skip[1].weight: dgs, self.skip[1].bias: dbs})
            if need_dx:
                dx = ops.add_acts(dx, dxs)
        else:
            dx, g = self.conv_list[0]._bwd(ctx["c0"], dy[..., :half], dnext, need_dx)
            grads.update(g)
            if need_dx:
                dx = ops.add_acts(dx, dy)
        return dx, grads


class STDCNet813(B200Module):
    """STDC1 backbone (reference stdcnet.py:116-204): returns the 1/2 .. 1/32 feature maps."""

    def __init__(self, base=64, layers=[2, 2, 2], block_num=4, type="cat", num_classes=1000, dropout=0.20,
                 pretrain_model='', use_conv_last=False):
        super(STDCNet813, self).__init__()
        block = CatBottleneck if type == "cat" else AddBottleneck
        self.use_conv_last = use_conv_last
        self.features = self._make_layers(base, layers, block_num, block)
        # classifier head of the ImageNet model: never used by the segmentation forward, kept so
        # that state_dict() has the reference's keys (conv_last.*, fc.weight, bn.*, linear.weight)
        self.conv_last = ConvX(base * 16, max(1024, base * 16), 1, 1)
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(max(1024, base * 16), max(1024, base * 16), bias=False)
        self.bn = nn.BatchNorm1d(max(1024, base * 16))
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(p=dropout)
        self.linear = nn.Linear(max(1024, base * 16), num_classes, bias=False)

        # aliases: the same modules registered a second time -> duplicated state_dict keys
        self.x2 = nn.Sequential(self.features[:1])
        self.x4 = nn.Sequential(self.features[1:2])
        self.x8 = nn.Sequential(self.features[2:4])
        self.x16 = nn.Sequential(self.features[4:6])
        self.x32 = nn.Sequential(self.features[6:])
        self._stage_ends = (1, 2, 4, 6, len(self.features))
        print('use pretrain model {}'.format(pretrain_model))
        if "STDCNet" in pretrain_model:
            self.init_weight(pretrain_model)
        else:
            self.init_params()

    def init_weight(self, pretrain_model):
        state_dict = torch.load(pretrain_model)["state_dict"]
        own = self.state_dict()
        for k, v in state_dict.items():
            own.update({k: v})
        self.load_state_dict(own)

    def init_params(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    init.constant_(m.bias, 0)

    def _make_layers(self, base, layers, block_num, block):
        features = [ConvX(3, base // 2, 3, 2), ConvX(base // 2, base, 3, 2)]
        for i, layer in enumerate(layers):
            for j in range(layer):
                if i == 0 and j == 0:
                    features.append(block(base, base * 4, block_num, 2))
                elif j == 0:
                    features.append(block(base * int(math.pow(2, i + 1)), base * int(math.pow(2, i + 2)), block_num, 2))
                else:
                    features.append(block(base * int(math.pow(2, i + 2)), base * int(math.pow(2, i + 2)), block_num, 1))
        return nn.Sequential(*features)

    # -- composite forward/backward over the five stages --------------------------------------
    def _fwd(self, img, feat8_out=None):
        feats, ctxs = [], []
        cur = img
        last = len(self.features) - 1
        for i, layer in enumerate(self.features):
            out = feat8_out if (feat8_out is not None and i + 1 == self._stage_ends[2]) else None
            cur, c = layer._fwd(cur, out=out)
            ctxs.append(c)
            if i + 1 in self._stage_ends:
                feats.append(cur)
        ctx = {"layers": ctxs}
        if self.use_conv_last:
            feats[4], ctx["last"] = self.conv_last._fwd(feats[4])
        return tuple(feats), ctx

    def _bwd(self, ctx, d2=None, d4=None, d8=None, d16=None, d32=None, need_dx=False):
        grads = {}
        if self.use_conv_last and d32 is not None:
            d32, g = self.conv_last._bwd(ctx["last"], d32)
            grads.update(g)
        ext = {self._stage_ends[0] - 1: d2, self._stage_ends[1] - 1: d4, self._stage_ends[2] - 1: d8,
               self._stage_ends[3] - 1: d16, self._stage_ends[4] - 1: d32}
        d = None
        for i in range(len(self.features) - 1, -1, -1):
            e = ext.get(i)
            if e is not None:
                d = e if d is None else ops.add_acts(d, e)
            if d is None:
                continue
            layer = self.features[i]
            if isinstance(layer, ConvX):
                d, g = layer._bwd(ctx["layers"][i], d, None, i > 0)
            else:
                d, g = layer._bwd(ctx["layers"][i], d, True)
            grads.update(g)
            if i == self._stage_ends[2]:
                # stages 4 and 5 hold ~85 % of the parameters and are done first: start averaging
                # their gradients over the ranks while the high-resolution stages still run
                ops.REDUCER.hook()
        return None, grads

    def _fwd_api(self, x):
        feats, ctx = self._fwd(x.float().contiguous())
        return tuple(f.permute(0, 3, 1, 2) for f in feats), ctx

    def forward_impl(self, x):
        raise NotImplementedError("the ImageNet classification head (reference stdcnet.py:196-204) "
                                  "is dead code for segmentation and not part of the hot path")
