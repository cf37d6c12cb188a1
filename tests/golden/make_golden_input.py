"""Golden vectors for the input pipeline, produced by the UNMODIFIED reference dataset classes
(build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_input.py

Small PNG trees are written under a temp dir in the layouts `dataset/cityscapes.py:22-57` and
`dataset/GTAV.py:62-79` expect, `CityScapes('train', root, a, b)` / `GtaV(root, None, a, b)` are
instantiated from cwd=/root/reference (GTAV.py:26 opens ./dataset/gta5_info.json) and their items
are stored next to the decoded inputs in tests/golden/reference_input.npz.
"""
import json
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
os.chdir(REF)

from dataset.cityscapes import CityScapes  # noqa: E402  (reference)
from dataset.GTAV import GtaV  # noqa: E402


def main():
    rng = np.random.default_rng(42)
    out = {}
    with open(os.path.join(REF, "dataset", "gta5_info.json")) as f:
        info = json.load(f)
    out["lb_map_ids"] = np.array([el["id"] for el in info], np.int64)
    out["lb_map_train"] = np.array([el["trainId"] for el in info], np.int64)
    with tempfile.TemporaryDirectory() as tmp:
        # ---- Cityscapes layout: images/<split>/<city>/*.png, gtFine/<split>/<city>/*labelTrainIds.png
        cs = os.path.join(tmp, "cs")
        shapes = [(96, 200), (61, 97)]
        for i, (h, w) in enumerate(shapes):
            os.makedirs(os.path.join(cs, "images", "train", "city%d" % i))
            os.makedirs(os.path.join(cs, "gtFine", "train", "city%d" % i))
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            lab = rng.integers(0, 20, (h, w), dtype=np.uint8)
            lab[lab == 19] = 255
            Image.fromarray(img).save(os.path.join(cs, "images", "train", "city%d" % i, "a_%d_leftImg8bit.png" % i))
            Image.fromarray(lab).save(os.path.join(cs, "gtFine", "train", "city%d" % i, "a_%d_gtFine_labelTrainIds.png" % i))
            Image.fromarray(img).save(os.path.join(cs, "gtFine", "train", "city%d" % i, "a_%d_gtFine_color.png" % i))
            out["cs_img%d" % i], out["cs_lab%d" % i] = img, lab
        for (a, b) in ((32, 48), (50, 40)):
            ds = CityScapes("train", cs, a, b)
            assert len(ds) == 2
            for i in range(2):
                x, y = ds[i]
                out["cs_out_img%d_%dx%d" % (i, a, b)] = x.numpy()
                out["cs_out_lab%d_%dx%d" % (i, a, b)] = y.numpy()
        # ---- GTA5 layout: images/*.png, labels/*.png (34-id labels)
        gta = os.path.join(tmp, "gta")
        os.makedirs(os.path.join(gta, "images"))
        os.makedirs(os.path.join(gta, "labels"))
        for i, (h, w) in enumerate([(105, 191), (64, 128)]):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            lab = rng.integers(0, 40, (h, w), dtype=np.uint8)
            Image.fromarray(img).save(os.path.join(gta, "images", "%05d.png" % i))
            Image.fromarray(lab).save(os.path.join(gta, "labels", "%05d.png" % i))
            out["gta_img%d" % i], out["gta_lab%d" % i] = img, lab
        for (a, b) in ((32, 64), (51, 102)):
            ds = GtaV(gta, None, a, b)
            for i in range(2):
                x, y = ds[i]
                out["gta_out_img%d_%dx%d" % (i, a, b)] = x.numpy()
                out["gta_out_lab%d_%dx%d" % (i, a, b)] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "reference_input.npz"), **out)
    print("wrote", len(out), "arrays;", {k: v.shape for k, v in out.items() if "out_img0" in k})


if __name__ == "__main__":
    main()
