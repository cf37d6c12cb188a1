"""CPU restatement of the reference's input pipeline — TEST INFRASTRUCTURE ONLY (imported by tests/,
``__graft_entry__.smoke()`` and bench.py's CPU legs; the product never touches it).

Reference path (dataset/cityscapes.py:61-69, dataset/GTAV.py:81-100):
    image = pil_loader(path).resize(self.resize, Image.BILINEAR)        # uint8 RGB
    label = Image.open(path).resize(self.resize, Image.NEAREST)         # uint8 L
    image = Normalize(mean, std)(ToTensor()(image))                     # fp32 CHW
    label = PILToTensor()(label)                                        # uint8 [1, H, W]
    label = convert_labels(label)          (GTAV only: 34 ids -> 19 train ids, in place, sequential)

The arithmetic lives in un-vendored Pillow (12.2.0 in this image; the reference pins no version) and
torchvision 0.26.  What is restated here is Pillow's published algorithm (src/libImaging/Resample.c
``precompute_coeffs`` / ``normalize_coeffs_8bpc`` / ``ImagingResampleHorizontal_8bpc`` /
``ImagingResampleVertical_8bpc``, and Geometry.c ``ImagingScaleAffine`` for NEAREST).  Parity is
pinned: tests/test_input_oracle.py checks every function bit-for-bit against Pillow / torchvision
themselves on random images (both libraries are part of the image, also on the GPU box) and
against committed fixtures generated through the reference's own dataset classes
(tests/golden/make_golden_input.py).

NOTE the reference passes ``(height, width)`` to ``Image.resize``, which takes ``(width, height)``:
``CityScapes(mode, root, 512, 1024)`` yields tensors of shape [3, 1024, 512].  ``resize_args`` keeps
that quirk: callers pass the same two numbers in the same order.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1).
    Returns (xmin int32[out], count int32[out], coeff int32[out, ksize])."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    count = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx], count[xx] = lo, n
    return xmin, count, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis(img, out_size, axis):
    """One 8-bit pass (ImagingResampleHorizontal_8bpc / Vertical_8bpc) along ``axis`` of [H, W, C]."""
    in_size = img.shape[axis]
    xmin, count, kk = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(count[xx]):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, out_w, out_h):
    """``Image.resize((out_w, out_h), Image.BILINEAR)`` on a uint8 [H, W, C] array: horizontal pass
    first, then vertical, each rounding to uint8 (Resample.c ImagingResampleInner)."""
    h, w = img.shape[:2]
    if out_w != w:
        img = _resample_axis(img, out_w, 1)
    if out_h != h:
        img = _resample_axis(img, out_h, 0)
    return img


def nearest_index(in_size, out_size):
    """Geometry.c ImagingScaleAffine: source index per output coordinate; the position is advanced
    by repeated double additions exactly as the C loop does."""
    a = float(in_size) / out_size
    idx = np.zeros(out_size, np.int64)
    xo = 0.0 + a * 0.5
    for x in range(out_size):
        xin = -1 if xo < 0.0 else int(xo)
        idx[x] = min(max(xin, 0), in_size - 1) if 0 <= xin < in_size else -1
        xo += a
    return idx


def resize_nearest_u8(lab, out_w, out_h):
    """``Image.resize((out_w, out_h), Image.NEAREST)`` on a uint8 [H, W] array."""
    h, w = lab.shape
    ix, iy = nearest_index(w, out_w), nearest_index(h, out_h)
    assert (ix >= 0).all() and (iy >= 0).all()
    return lab[iy][:, ix]


def to_tensor_normalize(img):
    """ToTensor + Normalize (torchvision): float32 ((u8 / 255) - mean) / std, CHW."""
    x = img.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    mean = np.asarray(MEAN, np.float32)[:, None, None]
    std = np.asarray(STD, np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


def convert_labels(label, lb_map):
    """GTAV.convert_labels (GTAV.py:97-100): sequential in-place remap in dict order."""
    label = label.copy()
    for k, v in lb_map.items():
        label[label == k] = v
    return label


def gta5_lb_map(info):
    """GTAV.py:26-28: {id: trainId} from the parsed gta5_info.json list."""
    return {el["id"]: el["trainId"] for el in info}


def cityscapes_item(image_u8, label_u8, a, b):
    """CityScapes.__getitem__ (cityscapes.py:61-69) after decoding; (a, b) = the dataset's
    (height, width) ctor arguments, handed to PIL as (width, height)."""
    return to_tensor_normalize(resize_bilinear_u8(image_u8, a, b)), resize_nearest_u8(label_u8, a, b)[None]


def gtav_item(image_u8, label_u8, a, b, lb_map):
    """GtaV.__getitem__ with aug_type=None (GTAV.py:81-91)."""
    img, lab = cityscapes_item(image_u8, label_u8, a, b)
    return img, convert_labels(lab, {k: v for k, v in lb_map.items() if 0 <= k <= 255})
