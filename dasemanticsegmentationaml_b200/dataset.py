"""Device-side input pipeline behind the reference's dataset classes.

Reference (dataset/cityscapes.py:12-75, dataset/GTAV.py:13-100): PIL decode -> ``resize(BILINEAR)`` /
``resize(NEAREST)`` -> ``ToTensor`` + ``Normalize`` / ``PILToTensor`` (+ the GTA5 34 -> 19 id remap),
all per item on CPU workers.  Here ``__getitem__`` only decodes; resize + normalise + remap run on
the GPU for the whole batch (``DevicePreprocess``), bit-identical to the CPU loader:

    ds = GtaV(root, None, 512, 1024)                      # same ctor as the reference
    loader = DataLoader(ds, batch_size=8, collate_fn=collate_raw, pin_memory=True, ...)
    for raw_images, raw_labels in loader:
        images, labels = ds.preprocess(raw_images, raw_labels)   # fp32 [N,3,H,W], uint8 [N,1,H,W] on the GPU

Kept quirk: the reference hands ``(height, width)`` to ``Image.resize``, which expects
``(width, height)`` — ``CityScapes('train', root, 512, 1024)`` yields [3, 1024, 512] tensors, and so
does this module.  Augmentations (``aug_type``) are out of scope and rejected.
"""
import os

import numpy as np
import torch

from . import kernels as K

MEAN = (0.485, 0.456, 0.406)   # cityscapes.py:20, GTAV.py:30
STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2    # Pillow Resample.c

# dataset/gta5_info.json reduced to {id: trainId} (GTAV.py:26-28); the standard Cityscapes label
# table.  id -1 (license plate) cannot occur in a uint8 label image.
GTA5_ID_TO_TRAINID = {0: 255, 1: 255, 2: 255, 3: 255, 4: 255, 5: 255, 6: 255, 7: 0, 8: 1, 9: 255, 10: 255,
                      11: 2, 12: 3, 13: 4, 14: 255, 15: 255, 16: 255, 17: 5, 18: 255, 19: 6, 20: 7, 21: 8,
                      22: 9, 23: 10, 24: 11, 25: 12, 26: 13, 27: 14, 28: 15, 29: 255, 30: 255, 31: 16,
                      32: 17, 33: 18, 34: 255, -1: 255}


def bilinear_tables(in_size, out_size):
    """Pillow's ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the triangle filter, vectorised
    in float64: bounds int32 [out, 2] = (first source index, taps), coefficients int32 [out, ksize]."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)     # C (int) cast truncates toward zero
    hi = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    cnt = hi - lo
    t = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((t + lo[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w[t >= cnt[:, None]] = 0.0
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):                                            # same left-to-right summation as C
        ww = ww + w[:, j]
    w = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = (0.5 + w * float(1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)
    kk[t.repeat(out_size, 0) >= cnt[:, None]] = 0
    bounds = np.stack([lo, cnt], 1).astype(np.int32)
    return bounds, kk


def nearest_table(in_size, out_size):
    """Pillow's ``ImagingScaleAffine`` source index per output coordinate (position advanced by
    repeated float64 additions, as the C loop does)."""
    a = float(in_size) / out_size
    pos = np.cumsum(np.concatenate([[a * 0.5], np.full(out_size - 1, a)]))   # sequential adds
    idx = pos.astype(np.int64)
    if (idx < 0).any() or (idx >= in_size).any():
        raise ValueError("nearest resize %d -> %d leaves the source image" % (in_size, out_size))
    return idx.astype(np.int32)


def normalize_table():
    """float32 [3, 256]: ((v / 255) - mean) / std with torchvision's fp32 operation order."""
    v = torch.arange(256, dtype=torch.float32).div(255)
    mean = torch.tensor(MEAN, dtype=torch.float32)[:, None]
    std = torch.tensor(STD, dtype=torch.float32)[:, None]
    return ((v[None, :] - mean) / std).contiguous()


def label_table(lb_map=None):
    """uint8 [256] remap; equals GtaV.convert_labels' sequential in-place loop because no trainId is
    visited as an id after it was produced (checked in tests/test_input_oracle.py)."""
    lut = np.arange(256, dtype=np.uint8)
    if lb_map:
        seq = lut.copy()
        for k, v in lb_map.items():                     # GTAV.py:97-100, literally
            if 0 <= k <= 255:
                seq[seq == k] = v
        lut = seq
    return torch.from_numpy(lut)


class ResizeTables(object):
    """Device-resident tables for one (source size -> output size) pair."""

    def __init__(self, h0, w0, out_h, out_w, device):
        xb, xk = bilinear_tables(w0, out_w)
        yb, yk = bilinear_tables(h0, out_h)
        first = yb[0::8, 0]
        last_rows = np.minimum(np.arange(0, out_h, 8) + 7, out_h - 1)
        self.max_rows = int((yb[last_rows, 0] + yb[last_rows, 1] - first).max())
        dev = torch.device(device)
        self.xb, self.xk = torch.from_numpy(xb).to(dev), torch.from_numpy(xk).to(dev)
        self.yb, self.yk = torch.from_numpy(yb).to(dev), torch.from_numpy(yk).to(dev)
        self.ix = torch.from_numpy(nearest_table(w0, out_w)).to(dev)
        self.iy = torch.from_numpy(nearest_table(h0, out_h)).to(dev)


class DevicePreprocess(object):
    """resize + ToTensor + Normalize (+ label remap) of a decoded uint8 batch on the GPU.

    ``height, width`` are the dataset constructors' arguments, in their order; as in the reference
    they reach ``Image.resize`` as ``(width, height)`` = ``(height, width)``, so the output is
    ``[N, 3, width, height]``."""

    def __init__(self, height, width, lb_map=None, device="cuda", label_dtype=torch.uint8):
        self.out_w, self.out_h = int(height), int(width)
        self.device = torch.device(device)
        self.label_dtype = label_dtype
        self._tables = {}
        self._norm_host, self._lut_host = normalize_table(), label_table(lb_map)
        self._norm = self._lut = None       # moved to the device on first use (datasets are built on CPU)

    def tables(self, h0, w0):
        if self._norm is None:
            self._norm, self._lut = self._norm_host.to(self.device), self._lut_host.to(self.device)
        key = (h0, w0)
        if key not in self._tables:
            self._tables[key] = ResizeTables(h0, w0, self.out_h, self.out_w, self.device)
        return self._tables[key]

    def images(self, raw):
        """uint8 [N, H0, W0, 3] (CPU or CUDA) -> fp32 [N, 3, out_h, out_w] on the device."""
        raw = raw.to(self.device, non_blocking=True).contiguous()
        t = self.tables(raw.shape[1], raw.shape[2])
        out = torch.empty((raw.shape[0], 3, self.out_h, self.out_w), dtype=torch.float32, device=self.device)
        return K.image_resize_normalize(raw, t.xb, t.xk, t.yb, t.yk, self._norm, out, t.max_rows)

    def labels(self, raw):
        """uint8 [N, H0, W0] -> label_dtype [N, 1, out_h, out_w] (PILToTensor keeps a channel axis)."""
        raw = raw.to(self.device, non_blocking=True).contiguous()
        t = self.tables(raw.shape[1], raw.shape[2])
        out = torch.empty((raw.shape[0], 1, self.out_h, self.out_w), dtype=self.label_dtype, device=self.device)
        return K.label_resize_remap(raw, t.ix, t.iy, self._lut, out)

    def __call__(self, raw_images, raw_labels):
        if isinstance(raw_images, (list, tuple)):       # mixed source sizes: one launch per item
            imgs = torch.cat([self.images(r[None]) for r in raw_images])
            labs = torch.cat([self.labels(r[None]) for r in raw_labels])
            return imgs, labs
        return self.images(raw_images), self.labels(raw_labels)


def collate_raw(batch):
    """DataLoader collate for undecorated items: stacks when every source image has the same size,
    otherwise returns lists (``DevicePreprocess`` accepts both)."""
    imgs, labs = zip(*batch)
    if all(i.shape == imgs[0].shape for i in imgs) and all(l.shape == labs[0].shape for l in labs):
        return torch.stack(imgs), torch.stack(labs)
    return list(imgs), list(labs)


def _decode(image_path, label_path):
    from PIL import Image
    with open(image_path, "rb") as f:
        image = np.asarray(Image.open(f).convert("RGB"))      # dataset/utils.py:11-14 pil_loader
    label = np.asarray(Image.open(label_path))
    return torch.from_numpy(image.copy()), torch.from_numpy(label.copy())


_IMG_EXT = (".png", ".jpg", ".jpeg")


class CityScapes(torch.utils.data.Dataset):
    """``CityScapes(mode, root, height, width)`` (dataset/cityscapes.py:12-75): same file discovery
    and pairing; items are the DECODED uint8 image [H0, W0, 3] and label [H0, W0] —
    ``self.preprocess`` finishes the job on the GPU."""

    def __init__(self, mode, root, height, width, device="cuda"):
        super().__init__()
        self.split = mode
        self.resize = (height, width)
        root = os.path.normpath(root)
        image_dir, label_dir = os.path.join(root, "images", mode), os.path.join(root, "gtFine", mode)
        self.root = os.path.normpath(image_dir)
        images, labels = [], []
        for city in os.listdir(image_dir):
            folder = os.path.join(image_dir, city)
            if os.path.isdir(folder):
                images += [os.path.join(folder, f) for f in os.listdir(folder) if f.lower().endswith(_IMG_EXT)]
        for city in os.listdir(label_dir):
            folder = os.path.join(label_dir, city)
            if os.path.isdir(folder):
                labels += [os.path.join(folder, f) for f in os.listdir(folder)
                           if f.lower().endswith(_IMG_EXT) and "color" not in f.lower()]
        self.data = list(zip(sorted(images), sorted(labels)))
        self.preprocess = DevicePreprocess(height, width, None, device)

    def __getitem__(self, idx):
        return _decode(*self.data[idx])

    def __len__(self):
        return len(self.data)


class GtaV(torch.utils.data.Dataset):
    """``GtaV(root, aug_type, height, width)`` (dataset/GTAV.py:13-100) with ``aug_type=None``."""

    def __init__(self, root, aug_type, height, width, device="cuda"):
        super().__init__()
        if aug_type is not None:
            raise NotImplementedError("augmentation pipelines (GTAV.py:32-52) are outside the accelerated path")
        self.root = os.path.normpath(root)
        self.resize = (height, width)
        self.lb_map = dict(GTA5_ID_TO_TRAINID)
        images = sorted(self.root + "/images/" + f for f in os.listdir(os.path.join(self.root, "images")))
        labels = sorted(self.root + "/labels/" + f for f in os.listdir(os.path.join(self.root, "labels")))
        self.data = list(zip(images, labels))
        self.preprocess = DevicePreprocess(height, width, self.lb_map, device)

    def __getitem__(self, idx):
        return _decode(*self.data[idx])

    def __len__(self):
        return len(self.data)

    def convert_labels(self, label):
        """Host-side equivalent of GTAV.py:97-100 for callers that still want it (uint8 tensor)."""
        return label_table(self.lb_map)[label.long()]
