"""Pins oracle/input_oracle.py: (1) against fixtures produced by the reference's own dataset classes
(tests/golden/make_golden_input.py), (2) against Pillow / torchvision themselves on random images."""
import os
import re

import numpy as np
import pytest

from oracle import input_oracle as I

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_input.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def lb_map(gold):
    return {int(k): int(v) for k, v in zip(gold["lb_map_ids"], gold["lb_map_train"])}


def test_dataset_items_bit_exact(gold):
    n = 0
    for key in gold.files:
        m = re.match(r"(cs|gta)_out_img(\d)_(\d+)x(\d+)", key)
        if not m:
            continue
        ds, i, a, b = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
        img, lab = gold["%s_img%d" % (ds, i)], gold["%s_lab%d" % (ds, i)]
        if ds == "cs":
            x, y = I.cityscapes_item(img, lab, a, b)
        else:
            x, y = I.gtav_item(img, lab, a, b, lb_map(gold))
        assert x.dtype == np.float32 and x.shape == gold[key].shape
        assert np.array_equal(x, gold[key]), key
        assert np.array_equal(y, gold[key.replace("_img", "_lab")]), key
        n += 1
    assert n == 8


def test_label_map_is_a_plain_lookup(gold):
    """The sequential in-place remap of GTAV.convert_labels never chains (every trainId was already
    visited as an id), so a 256-entry table reproduces it; ids >= 35 pass through."""
    m = lb_map(gold)
    ids = np.arange(256, dtype=np.uint8)
    seq = I.convert_labels(ids, {k: v for k, v in m.items() if 0 <= k <= 255})
    lut = ids.copy()
    for k, v in m.items():
        if 0 <= k <= 255:
            lut[k] = v
    assert np.array_equal(seq, lut)


@pytest.mark.parametrize("shape", [(64, 96, 48, 32), (61, 97, 50, 33), (40, 50, 100, 77), (33, 70, 70, 33),
                                   (263, 479, 128, 256)])
def test_resize_matches_pillow(shape):
    from PIL import Image
    h, w, ow, oh = shape
    rng = np.random.default_rng(h * w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    lab = rng.integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(I.resize_bilinear_u8(img, ow, oh), np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
    assert np.array_equal(I.resize_nearest_u8(lab, ow, oh), np.array(Image.fromarray(lab).resize((ow, oh), Image.NEAREST)))


def test_nearest_index_tables_at_dataset_sizes():
    """GTA5 1914x1052 and Cityscapes 2048x1024 sources at the README's (512, 1024) arguments."""
    from PIL import Image
    for (h, w) in ((1052, 1914), (1024, 2048)):
        ramp_x = (np.arange(w) % 251).astype(np.uint8)[None, :].repeat(4, 0)
        ramp_y = (np.arange(h) % 251).astype(np.uint8)[:, None].repeat(4, 1)
        for out in (512, 1024, 720, 1280):
            ix = I.nearest_index(w, out)
            assert np.array_equal(ramp_x[:, ix], np.array(Image.fromarray(ramp_x).resize((out, 4), Image.NEAREST)))
            iy = I.nearest_index(h, out)
            assert np.array_equal(ramp_y[iy], np.array(Image.fromarray(ramp_y).resize((4, out), Image.NEAREST)))


def test_normalize_matches_torchvision():
    import torch
    from PIL import Image
    from torchvision import transforms
    vals = np.arange(256, dtype=np.uint8)
    img = np.stack([vals, vals[::-1], np.roll(vals, 7)], -1)[None].repeat(2, 0)  # [2, 256, 3]
    t = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=I.MEAN, std=I.STD)])
    ref = t(Image.fromarray(img)).numpy()
    assert np.array_equal(I.to_tensor_normalize(img), ref)
