"""Device input pipeline (csrc/input.cu through the C ABI) against the reference fixtures and the
CPU oracle: everything here is byte / table work, so the bar is bit-exact."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import input_oracle as I

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_input.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_reference_dataset_items_bit_exact(cuda_lib, gold):
    from dasemanticsegmentationaml_b200 import dataset as D
    n = 0
    for key in gold.files:
        m = re.match(r"(cs|gta)_out_img(\d)_(\d+)x(\d+)", key)
        if not m:
            continue
        ds, i, a, b = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
        pre = D.DevicePreprocess(a, b, D.GTA5_ID_TO_TRAINID if ds == "gta" else None)
        img = torch.from_numpy(gold["%s_img%d" % (ds, i)])[None]
        lab = torch.from_numpy(gold["%s_lab%d" % (ds, i)])[None]
        x, y = pre(img, lab)
        assert x.dtype == torch.float32 and y.dtype == torch.uint8
        assert np.array_equal(x[0].cpu().numpy(), gold[key]), key
        assert np.array_equal(y[0].cpu().numpy(), gold[key.replace("_img", "_lab")]), key
        n += 1
    assert n == 8
    lb = {int(k): int(v) for k, v in zip(gold["lb_map_ids"], gold["lb_map_train"])}
    assert lb == D.GTA5_ID_TO_TRAINID


@pytest.mark.parametrize("case", [((1052, 1914), (512, 1024)),     # GTA5 source, README arguments
                                  ((1024, 2048), (512, 1024)),     # Cityscapes source
                                  ((1024, 2048), (720, 1280)),
                                  ((300, 420), (700, 900)),        # up-sampling
                                  ((1024, 2048), (64, 96))])       # 16-21x down-sampling (wide bands)
def test_batches_match_oracle_at_dataset_sizes(cuda_lib, case):
    from dasemanticsegmentationaml_b200 import dataset as D
    (h0, w0), (a, b) = case
    rng = np.random.default_rng(h0 + a)
    img = rng.integers(0, 256, (2, h0, w0, 3), dtype=np.uint8)
    lab = rng.integers(0, 40, (2, h0, w0), dtype=np.uint8)
    pre = D.DevicePreprocess(a, b, D.GTA5_ID_TO_TRAINID, label_dtype=torch.int64)
    x, y = pre(torch.from_numpy(img), torch.from_numpy(lab))
    assert x.shape == (2, 3, b, a) and y.shape == (2, 1, b, a) and y.dtype == torch.int64
    for i in range(2):
        xo, yo = I.gtav_item(img[i], lab[i], a, b, D.GTA5_ID_TO_TRAINID)
        assert np.array_equal(x[i].cpu().numpy(), xo)
        assert np.array_equal(y[i].cpu().numpy(), yo.astype(np.int64))


def test_dataset_classes_end_to_end(cuda_lib, tmp_path):
    """PNG trees in the reference's directory layouts -> our Dataset classes -> DataLoader ->
    device preprocess, against the oracle's per-item pipeline (mixed source sizes included)."""
    from PIL import Image
    from dasemanticsegmentationaml_b200 import dataset as D
    rng = np.random.default_rng(3)
    gta = tmp_path / "gta"
    (gta / "images").mkdir(parents=True)
    (gta / "labels").mkdir()
    raw = []
    for i, (h, w) in enumerate([(105, 191), (105, 191), (64, 128)]):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        lab = rng.integers(0, 36, (h, w), dtype=np.uint8)
        Image.fromarray(img).save(str(gta / "images" / ("%05d.png" % i)))
        Image.fromarray(lab).save(str(gta / "labels" / ("%05d.png" % i)))
        raw.append((img, lab))
    ds = D.GtaV(str(gta), None, 32, 64)
    assert len(ds) == 3
    for batch_ids in ([0, 1], [1, 2]):
        ri, rl = D.collate_raw([ds[i] for i in batch_ids])
        x, y = ds.preprocess(ri, rl)
        for j, i in enumerate(batch_ids):
            xo, yo = I.gtav_item(raw[i][0], raw[i][1], 32, 64, D.GTA5_ID_TO_TRAINID)
            assert np.array_equal(x[j].cpu().numpy(), xo) and np.array_equal(y[j].cpu().numpy(), yo)
    with pytest.raises(NotImplementedError):
        D.GtaV(str(gta), "CS-HF", 32, 64)
    cs = tmp_path / "cs"
    for sub in ("images/val/c0", "gtFine/val/c0"):
        (cs / sub).mkdir(parents=True)
    img = rng.integers(0, 256, (61, 97, 3), dtype=np.uint8)
    lab = rng.integers(0, 19, (61, 97), dtype=np.uint8)
    Image.fromarray(img).save(str(cs / "images/val/c0/x_leftImg8bit.png"))
    Image.fromarray(lab).save(str(cs / "gtFine/val/c0/x_gtFine_labelTrainIds.png"))
    Image.fromarray(img).save(str(cs / "gtFine/val/c0/x_gtFine_color.png"))
    dsc = D.CityScapes("val", str(cs), 50, 40)
    assert len(dsc) == 1
    x, y = dsc.preprocess(*D.collate_raw([dsc[0]]))
    xo, yo = I.cityscapes_item(img, lab, 50, 40)
    assert np.array_equal(x[0].cpu().numpy(), xo) and np.array_equal(y[0].cpu().numpy(), yo)
