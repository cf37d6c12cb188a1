"""ORACLE — test infrastructure, not product code.

A plain fp32 / CPU restatement of the reference's hot path (STDC-BiSeNet, the three domain
discriminators, the losses and the mIoU metric), written as pure functions over a ``state_dict``
so that it can be evaluated on exactly the weights the CUDA modules hold.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
this package; the product (``dasemanticsegmentationaml_b200``) never does.

Parity pin: ``tests/golden/make_golden.py`` executes the UNMODIFIED reference (imported from
/root/reference, which only exists in the build container) on seeded inputs and stores its outputs
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement against those
vectors.  The arithmetic itself lives in PyTorch / numpy (un-vendored, versions unpinned by the
reference's requirements.txt; torch 2.11.0, numpy 2.3 here).

Every function cites the reference lines it restates (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# Quantisation-matched mode (tests only): the CUDA path stores every conv output and every
# activation as bf16; with QUANT = True the oracle rounds the same tensors (straight-through in the
# backward), so that a comparison isolates arithmetic differences from the storage format
# (SURVEY 8c "quantisation-matched oracle").  Off for the golden-vector tests.
QUANT = False


def _q(x):
    if not QUANT:
        return x
    # straight-through: rounding the oracle's gradients as well was tried and LOWERS the agreement
    # (the CUDA path sums some gradients in fp32 before its single rounding, so the rounding points
    # do not coincide and the two noises add instead of cancelling)
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


# --------------------------------------------------------------------------- building blocks
def _bn(sd, prefix, x, training):
    """nn.BatchNorm2d forward (train: batch statistics + running-stat update, eval: running stats)."""
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and (prefix + ".num_batches_tracked") in sd:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training,
                        BN_MOMENTUM, BN_EPS)


def conv_bn_relu(sd, prefix, x, stride, padding, training):
    """ConvX (model/stdcnet.py:6-15) and ConvBNReLU (model/model_stages.py:11-29):
    bias-free conv -> BatchNorm2d -> ReLU."""
    y = _q(F.conv2d(x, sd[prefix + ".conv.weight"], None, stride, padding))
    return _q(F.relu(_bn(sd, prefix + ".bn", y, training)))


def cat_bottleneck(sd, prefix, x, out_planes, stride, training, block_num=4):
    """CatBottleneck.forward (model/stdcnet.py:94-113) for block_num=4."""
    out1 = conv_bn_relu(sd, prefix + ".conv_list.0", x, 1, 0, training)
    feats = []
    cur = out1
    if stride == 2:
        c = out_planes // 2
        cur = _q(F.conv2d(out1, sd[prefix + ".avd_layer.0.weight"], None, 2, 1, 1, c))  # stdcnet.py:73-77
        cur = _q(_bn(sd, prefix + ".avd_layer.1", cur, training))
    for idx in range(1, block_num):
        cur = conv_bn_relu(sd, prefix + ".conv_list.%d" % idx, cur, 1, 1, training)
        feats.append(cur)
    skip = _q(F.avg_pool2d(out1, 3, 2, 1)) if stride == 2 else out1  # stdcnet.py:78,108-110
    return torch.cat([skip] + feats, dim=1)


def add_bottleneck(sd, prefix, x, out_planes, stride, training, block_num=4):
    """AddBottleneck.forward (model/stdcnet.py:49-64) for block_num=4: the first list entry is the
    (depthwise-strided) 1x1 output itself, and the skip is x (stride 1) or
    dw3x3/s2 -> BN -> 1x1 -> BN of x (stride 2, stdcnet.py:30-35)."""
    cur = conv_bn_relu(sd, prefix + ".conv_list.0", x, 1, 0, training)
    if stride == 2:
        c = out_planes // 2
        cur = _q(F.conv2d(cur, sd[prefix + ".avd_layer.0.weight"], None, 2, 1, 1, c))    # stdcnet.py:24-28
        cur = _q(_bn(sd, prefix + ".avd_layer.1", cur, training))
    feats = [cur]
    for idx in range(1, block_num):
        cur = conv_bn_relu(sd, prefix + ".conv_list.%d" % idx, cur, 1, 1, training)
        feats.append(cur)
    if stride == 2:
        cin = x.shape[1]
        skip = _q(F.conv2d(x, sd[prefix + ".skip.0.weight"], None, 2, 1, 1, cin))
        skip = _q(_bn(sd, prefix + ".skip.1", skip, training))
        skip = _q(F.conv2d(skip, sd[prefix + ".skip.2.weight"]))
        skip = _q(_bn(sd, prefix + ".skip.3", skip, training))
    else:
        skip = x
    return _q(torch.cat(feats, dim=1) + skip)


def stdcnet813(sd, prefix, x, training, block="cat", use_conv_last=False):
    """STDCNet813.forward (model/stdcnet.py:185-194) with layers=[2,2,2], block_num=4, base=64;
    `block` = the constructor's `type` ("cat" is what BiSeNet builds), `use_conv_last` adds the 1x1
    ConvX on feat32 (stdcnet.py:126,191-192)."""
    f = prefix + ".features"
    blk = cat_bottleneck if block == "cat" else add_bottleneck
    feat2 = conv_bn_relu(sd, f + ".0", x, 2, 1, training)
    feat4 = conv_bn_relu(sd, f + ".1", feat2, 2, 1, training)
    feat8 = blk(sd, f + ".2", feat4, 256, 2, training)
    feat8 = blk(sd, f + ".3", feat8, 256, 1, training)
    feat16 = blk(sd, f + ".4", feat8, 512, 2, training)
    feat16 = blk(sd, f + ".5", feat16, 512, 1, training)
    feat32 = blk(sd, f + ".6", feat16, 1024, 2, training)
    feat32 = blk(sd, f + ".7", feat32, 1024, 1, training)
    if use_conv_last:
        feat32 = conv_bn_relu(sd, prefix + ".conv_last", feat32, 1, 0, training)
    return feat2, feat4, feat8, feat16, feat32


def attention_refinement(sd, prefix, x, training):
    """AttentionRefinementModule.forward (model/model_stages.py:77-85)."""
    feat = conv_bn_relu(sd, prefix + ".conv", x, 1, 1, training)
    atten = F.avg_pool2d(feat, feat.shape[2:])
    atten = F.conv2d(atten, sd[prefix + ".conv_atten.weight"])
    atten = torch.sigmoid(_bn(sd, prefix + ".bn_atten", atten, training))
    return feat * atten


def context_path(sd, prefix, x, training):
    """ContextPath.forward (model/model_stages.py:112-135).  The optional 1x1 ConvX on feat32
    (use_conv_last, model_stages.py:98 -> stdcnet.py:191-192) is applied when the state holds its weights."""
    last = (prefix + ".backbone.conv_last.conv.weight") in sd
    feat2, feat4, feat8, feat16, feat32 = stdcnet813(sd, prefix + ".backbone", x, training, use_conv_last=last)
    avg = F.avg_pool2d(feat32, feat32.shape[2:])
    avg = conv_bn_relu(sd, prefix + ".conv_avg", avg, 1, 0, training)
    avg_up = F.interpolate(avg, feat32.shape[2:], mode="nearest")
    feat32_sum = _q(attention_refinement(sd, prefix + ".arm32", feat32, training) + avg_up)
    feat32_up = F.interpolate(feat32_sum, feat16.shape[2:], mode="nearest")
    feat32_up = conv_bn_relu(sd, prefix + ".conv_head32", feat32_up, 1, 1, training)
    feat16_sum = _q(attention_refinement(sd, prefix + ".arm16", feat16, training) + feat32_up)
    feat16_up = F.interpolate(feat16_sum, feat8.shape[2:], mode="nearest")
    feat16_up = conv_bn_relu(sd, prefix + ".conv_head16", feat16_up, 1, 1, training)
    return feat2, feat4, feat8, feat16, feat16_up, feat32_up


def feature_fusion(sd, prefix, fsp, fcp, training):
    """FeatureFusionModule.forward (model/model_stages.py:175-185)."""
    feat = conv_bn_relu(sd, prefix + ".convblk", torch.cat([fsp, fcp], dim=1), 1, 0, training)
    atten = F.avg_pool2d(feat, feat.shape[2:])
    atten = F.relu(F.conv2d(atten, sd[prefix + ".conv1.weight"]))
    atten = torch.sigmoid(F.conv2d(atten, sd[prefix + ".conv2.weight"]))
    return _q(feat * atten + feat)


def seg_head(sd, prefix, x, training):
    """BiSeNetOutput.forward (model/model_stages.py:45-48)."""
    return F.conv2d(conv_bn_relu(sd, prefix + ".conv", x, 1, 1, training), sd[prefix + ".conv_out.weight"])


def bisenet_lowres(sd, x, training):
    """BiSeNet.forward up to the three class-logit maps before up-sampling (model_stages.py:229-238)."""
    _, _, feat8, _, feat_cp8, feat_cp16 = context_path(sd, "cp", x, training)
    fuse = feature_fusion(sd, "ffm", feat8, feat_cp8, training)
    return (seg_head(sd, "conv_out", fuse, training), seg_head(sd, "conv_out16", feat_cp8, training),
            seg_head(sd, "conv_out32", feat_cp16, training))


def bisenet_forward(sd, x, training=False):
    """BiSeNet.forward (model/model_stages.py:229-244): three [N, classes, H, W] logit maps."""
    size = x.shape[2:]
    return tuple(F.interpolate(t, size, mode="bilinear", align_corners=True)
                 for t in bisenet_lowres(sd, x, training))


# --------------------------------------------------------------------------- discriminators
def fc_discriminator(sd, x, prefix=""):
    """FCDiscriminator.forward (model/discriminator.py:17-28)."""
    for name in ("conv1", "conv2", "conv3", "conv4"):
        x = _q(F.leaky_relu(F.conv2d(x, sd[prefix + name + ".weight"], sd[prefix + name + ".bias"], 2, 1), 0.2))
    return F.conv2d(x, sd[prefix + "classifier.weight"], sd[prefix + "classifier.bias"], 2, 1)


def dwsep_discriminator(sd, x, batch_norm=False, training=True, prefix=""):
    """DepthWiseSepFCDiscriminator.forward (model/discriminator.py:51-73) and the BatchNorm variant
    DepthWiseSepBNFCDiscriminator.forward (discriminator.py:103-134).  The point-wise convs keep the
    reference's kernel_size=1, padding=1 (discriminator.py:36,39,42,45), which grows the map."""
    for i in (1, 2, 3, 4):
        wd = sd[prefix + "conv%d_d.weight" % i]
        x = F.conv2d(x, wd, sd[prefix + "conv%d_d.bias" % i], 2, 1, 1, wd.shape[0])
        if batch_norm:
            x = _bn(sd, prefix + "bn%d_d" % i, _q(x), training)
        x = _q(F.leaky_relu(x, 0.2))
        x = F.conv2d(x, sd[prefix + "conv%d_p.weight" % i], sd[prefix + "conv%d_p.bias" % i], 1, 1)
        if batch_norm:
            x = _bn(sd, prefix + "bn%d_p" % i, _q(x), training)
        x = _q(F.leaky_relu(x, 0.2))
    return F.conv2d(x, sd[prefix + "classifier.weight"], sd[prefix + "classifier.bias"], 2, 1)


# --------------------------------------------------------------------------- losses
def cross_entropy_ignore(logits, target, ignore_index=255):
    """torch.nn.CrossEntropyLoss(ignore_index=255) as used at train.py:66,86-89,135,214-217."""
    return F.cross_entropy(logits, target, ignore_index=ignore_index)


def ohem_cross_entropy(logits, target, threshold, keep_num):
    """OHEM_CrossEntroy_Loss.forward (utils.py:263-271)."""
    loss = F.cross_entropy(logits, target, reduction="none").view(-1)
    loss, _ = torch.sort(loss, descending=True)
    if loss[keep_num] > threshold:
        loss = loss[loss > threshold]
    else:
        loss = loss[:keep_num]
    return loss.mean()


def bce_with_logits_const(logits, target_value):
    """BCEWithLogitsLoss against an all-zeros / all-ones map (train.py:173,231-232,249-250,258)."""
    return F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, float(target_value)))


# --------------------------------------------------------------------------- metrics (numpy)
def fast_hist(a, b, n):
    """utils.py:161-167 — called as fast_hist(label, pred, n) at train.py:47."""
    k = (a >= 0) & (a < n)
    return np.bincount(n * a[k].astype(int) + b[k], minlength=n ** 2).reshape(n, n)


def per_class_iu(hist):
    """utils.py:170-172."""
    epsilon = 1e-5
    return np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist) + epsilon)


def reverse_one_hot(image):
    """utils.py:98-122: [C, H, W] scores -> [H, W] int64 class map."""
    return torch.argmax(image.permute(1, 2, 0), dim=-1)


def compute_global_accuracy(pred, label):
    """utils.py:151-159 (vectorised: the reference counts equal positions in a Python loop)."""
    pred = np.asarray(pred).reshape(-1)
    label = np.asarray(label).reshape(-1)
    return float((pred == label).sum()) / float(len(label))


def poly_lr(init_lr, it, max_iter=300, power=0.9):
    """utils.py:11-26."""
    return init_lr * (1 - it / max_iter) ** power


# --------------------------------------------------------------------------- evaluation / steps
def eval_batch(sd, images, labels, n_classes=19):
    """val() body (train.py:30-47) for a batch: forward, per-sample argmax, summed confusion matrix."""
    with torch.no_grad():
        out = bisenet_forward(sd, images, training=False)[0]
    hist = np.zeros((n_classes, n_classes), dtype=np.int64)
    for i in range(images.shape[0]):
        pred = reverse_one_hot(out[i]).numpy()
        hist += fast_hist(labels[i].numpy().reshape(-1), pred.reshape(-1), n_classes)
    return out, hist


def supervised_loss(sd, images, labels, training=True):
    """Loss composition of train() (train.py:83-89) / the source pass of train_DA (train.py:211-217)."""
    out, out16, out32 = bisenet_forward(sd, images, training)
    return (cross_entropy_ignore(out, labels) + cross_entropy_ignore(out16, labels)
            + cross_entropy_ignore(out32, labels)), (out, out16, out32)


def discriminator_forward(kind, sd, x, training=True):
    if kind == "dense":
        return fc_discriminator(sd, x)
    return dwsep_discriminator(sd, x, batch_norm=(kind == "dwsep_bn"), training=training)


def da_step(seg_sd, d_sd, kind, images, labels, images_t, opt_seg, opt_d, lambda_adv=0.001):
    """One iteration of train_DA's inner loop (train.py:192-262) without AMP (fp32 throughout).

    ``seg_sd`` / ``d_sd`` hold leaf tensors with requires_grad set on the parameters; ``opt_seg`` /
    ``opt_d`` are torch optimizers over those leaves (SGD / Adam as at train.py:170-172).
    Returns (loss_seg, loss_adv_for_G, loss_D_source, loss_D_target) as floats.
    """
    d_params = [v for v in d_sd.values() if v.is_floating_point() and v.requires_grad]
    opt_seg.zero_grad()
    opt_d.zero_grad()
    for p in d_params:  # train.py:207-208
        p.requires_grad_(False)
    loss, (out, _, _) = supervised_loss(seg_sd, images, labels, True)
    loss.backward()
    opt_seg.step()  # train.py:219-221
    out_t, _, _ = bisenet_forward(seg_sd, images_t, True)  # train.py:223-224
    opt_seg.zero_grad()
    d_out = discriminator_forward(kind, d_sd, F.softmax(out_t, dim=1))
    loss_adv_g = bce_with_logits_const(d_out, 0.0)  # train.py:230-232
    (loss_adv_g * lambda_adv).backward()
    opt_seg.step()  # train.py:234-237
    for p in d_params:  # train.py:240-241
        p.requires_grad_(True)
    out, out_t = out.detach(), out_t.detach()
    d_out = discriminator_forward(kind, d_sd, F.softmax(out, dim=1))
    loss_d_src = bce_with_logits_const(d_out, 0.0)  # train.py:247-250
    loss_d_src.backward()
    opt_d.step()
    d_out = discriminator_forward(kind, d_sd, F.softmax(out_t, dim=1))
    loss_d_tgt = bce_with_logits_const(d_out, 1.0)  # train.py:256-258
    opt_d.zero_grad()
    loss_d_tgt.backward()
    opt_d.step()  # train.py:259-262
    return float(loss), float(loss_adv_g), float(loss_d_src), float(loss_d_tgt)


def da_step_nni(seg_sd, d_sd, kind, images, labels, images_t, opt_seg, opt_d, lambda_adv=0.001):
    """One iteration of train_nni.py's inner loop (train_nni.py:105-163), fp32, no AMP: the
    adversarial term and the discriminator see softmax(out32) instead of softmax(out), gradients of
    both passes accumulate and each optimizer steps ONCE at the end.
    Returns (loss_seg, loss_adv_for_G * lambda, loss_D_source, loss_D_target) as floats."""
    d_params = [v for v in d_sd.values() if v.is_floating_point() and v.requires_grad]
    for p in d_params:  # train_nni.py:108-109
        p.requires_grad_(False)
    opt_seg.zero_grad()
    opt_d.zero_grad()
    loss, (_, _, out32) = supervised_loss(seg_sd, images, labels, True)  # train_nni.py:118-124
    loss.backward()
    _, _, out32_t = bisenet_forward(seg_sd, images_t, True)  # train_nni.py:131-136
    d_out = discriminator_forward(kind, d_sd, F.softmax(out32_t, dim=1))
    loss_d1 = bce_with_logits_const(d_out, 0.0) * lambda_adv
    loss_d1.backward()
    for p in d_params:  # train_nni.py:141-142
        p.requires_grad_(True)
    out32, out32_t = out32.detach(), out32_t.detach()
    d_out = discriminator_forward(kind, d_sd, F.softmax(out32, dim=1))
    loss_d_src = bce_with_logits_const(d_out, 0.0)  # train_nni.py:147-151
    loss_d_src.backward()
    d_out = discriminator_forward(kind, d_sd, F.softmax(out32_t, dim=1))
    loss_d_tgt = bce_with_logits_const(d_out, 1.0)  # train_nni.py:153-157
    loss_d_tgt.backward()
    opt_seg.step()  # train_nni.py:159-161
    opt_d.step()
    return float(loss), float(loss_d1), float(loss_d_src), float(loss_d_tgt)


# --------------------------------------------------------------------------- synthetic weights
def _kaiming(shape, gen, a=0.0, mode="fan_in"):
    fan_in = shape[1] * shape[2] * shape[3]
    fan_out = shape[0] * shape[2] * shape[3]
    fan = fan_in if mode == "fan_in" else fan_out
    std = math.sqrt(2.0 / (1 + a * a)) / math.sqrt(fan)
    return torch.randn(shape, generator=gen) * std


def _add_bn(sd, prefix, c, gen, randomize):
    sd[prefix + ".weight"] = torch.ones(c) if not randomize else 1.0 + 0.1 * torch.randn(c, generator=gen)
    sd[prefix + ".bias"] = torch.zeros(c) if not randomize else 0.1 * torch.randn(c, generator=gen)
    sd[prefix + ".running_mean"] = torch.zeros(c) if not randomize else 0.1 * torch.randn(c, generator=gen)
    sd[prefix + ".running_var"] = torch.ones(c) if not randomize else 1.0 + 0.2 * torch.rand(c, generator=gen)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def make_bisenet_state(seed=0, n_classes=19, randomize_bn=False):
    """Random-init weights with the reference's state_dict layout *without* the alias keys and the
    dead classifier head (those are added by the product module; the oracle never reads them).
    Init follows STDCNet813.init_params (stdcnet.py:155-167, kaiming fan_out) for the backbone and
    kaiming_normal_(a=1) for everything else (model_stages.py:31-35 and siblings)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}

    def convx(prefix, cin, cout, k, backbone):
        sd[prefix + ".conv.weight"] = _kaiming((cout, cin, k, k), gen, 0.0 if backbone else 1.0,
                                               "fan_out" if backbone else "fan_in")
        _add_bn(sd, prefix + ".bn", cout, gen, randomize_bn)

    def block(prefix, cin, cout, stride):
        convx(prefix + ".conv_list.0", cin, cout // 2, 1, True)
        convx(prefix + ".conv_list.1", cout // 2, cout // 4, 3, True)
        convx(prefix + ".conv_list.2", cout // 4, cout // 8, 3, True)
        convx(prefix + ".conv_list.3", cout // 8, cout // 8, 3, True)
        if stride == 2:
            sd[prefix + ".avd_layer.0.weight"] = _kaiming((cout // 2, 1, 3, 3), gen, 0.0, "fan_out")
            _add_bn(sd, prefix + ".avd_layer.1", cout // 2, gen, randomize_bn)

    f = "cp.backbone.features"
    convx(f + ".0", 3, 32, 3, True)
    convx(f + ".1", 32, 64, 3, True)
    block(f + ".2", 64, 256, 2)
    block(f + ".3", 256, 256, 1)
    block(f + ".4", 256, 512, 2)
    block(f + ".5", 512, 512, 1)
    block(f + ".6", 512, 1024, 2)
    block(f + ".7", 1024, 1024, 1)
    for name, cin in (("cp.arm16", 512), ("cp.arm32", 1024)):
        convx(name + ".conv", cin, 128, 3, False)
        sd[name + ".conv_atten.weight"] = _kaiming((128, 128, 1, 1), gen, 1.0)
        _add_bn(sd, name + ".bn_atten", 128, gen, randomize_bn)
    convx("cp.conv_head32", 128, 128, 3, False)
    convx("cp.conv_head16", 128, 128, 3, False)
    convx("cp.conv_avg", 1024, 128, 1, False)
    convx("ffm.convblk", 384, 256, 1, False)
    sd["ffm.conv1.weight"] = _kaiming((64, 256, 1, 1), gen, 1.0)
    sd["ffm.conv2.weight"] = _kaiming((256, 64, 1, 1), gen, 1.0)
    for name, cin, mid in (("conv_out", 256, 256), ("conv_out16", 128, 64), ("conv_out32", 128, 64)):
        convx(name + ".conv", cin, mid, 3, False)
        sd[name + ".conv_out.weight"] = _kaiming((n_classes, mid, 1, 1), gen, 1.0)
    return sd


def make_backbone_state(seed=0, block="cat", use_conv_last=False, randomize_bn=True, prefix="bb."):
    """Random weights for a stand-alone STDCNet813 (no alias keys, no ImageNet head except
    conv_last when asked): `block` = "cat" | "add" (stdcnet.py:117,119-122)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}

    def convx(name, cin, cout, k):
        sd[name + ".conv.weight"] = _kaiming((cout, cin, k, k), gen, 0.0, "fan_out")
        _add_bn(sd, name + ".bn", cout, gen, randomize_bn)

    def blk(name, cin, cout, stride):
        convx(name + ".conv_list.0", cin, cout // 2, 1)
        convx(name + ".conv_list.1", cout // 2, cout // 4, 3)
        convx(name + ".conv_list.2", cout // 4, cout // 8, 3)
        convx(name + ".conv_list.3", cout // 8, cout // 8, 3)
        if stride == 2:
            sd[name + ".avd_layer.0.weight"] = _kaiming((cout // 2, 1, 3, 3), gen, 0.0, "fan_out")
            _add_bn(sd, name + ".avd_layer.1", cout // 2, gen, randomize_bn)
            if block == "add":
                sd[name + ".skip.0.weight"] = _kaiming((cin, 1, 3, 3), gen, 0.0, "fan_out")
                _add_bn(sd, name + ".skip.1", cin, gen, randomize_bn)
                sd[name + ".skip.2.weight"] = _kaiming((cout, cin, 1, 1), gen, 0.0, "fan_out")
                _add_bn(sd, name + ".skip.3", cout, gen, randomize_bn)

    f = prefix + "features"
    convx(f + ".0", 3, 32, 3)
    convx(f + ".1", 32, 64, 3)
    for i, (cin, cout, stride) in enumerate(((64, 256, 2), (256, 256, 1), (256, 512, 2), (512, 512, 1),
                                             (512, 1024, 2), (1024, 1024, 1))):
        blk(f + ".%d" % (i + 2), cin, cout, stride)
    if use_conv_last:
        convx(prefix + "conv_last", 1024, 1024, 1)
    return sd


def make_discriminator_state(kind, seed=0, n_classes=19, ndf=64):
    """Weights for FCDiscriminator / DepthWiseSep[BN]FCDiscriminator with nn.Conv2d's default init
    scale (uniform +-1/sqrt(fan_in)); key layout of discriminator.py:9-13, 34-47, 79-97."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin_per_group, k):
        bound = 1.0 / math.sqrt(cin_per_group * k * k)
        sd[name + ".weight"] = (torch.rand((cout, cin_per_group, k, k), generator=gen) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand((cout,), generator=gen) * 2 - 1) * bound

    chans = [n_classes, ndf, ndf * 2, ndf * 4, ndf * 8]
    if kind == "dense":
        for i in range(4):
            conv("conv%d" % (i + 1), chans[i + 1], chans[i], 4)
    else:
        for i in range(4):
            conv("conv%d_d" % (i + 1), chans[i], 1, 4)
            if kind == "dwsep_bn":
                _add_bn(sd, "bn%d_d" % (i + 1), chans[i], gen, False)
            conv("conv%d_p" % (i + 1), chans[i + 1], chans[i], 1)
            if kind == "dwsep_bn":
                _add_bn(sd, "bn%d_p" % (i + 1), chans[i + 1], gen, False)
    conv("classifier", 1, chans[4], 4)
    return sd


def clone_state(sd, requires_grad=False):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and "running_" not in k:
            t.requires_grad_(True)
        out[k] = t
    return out
