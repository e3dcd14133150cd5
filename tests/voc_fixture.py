"""Synthetic VOC12-shaped tree for the loader tests: JPEGImages/*.jpg of different sizes, SegmentationClass/*.png,
Annotations/*.xml, a list file in the reference's format and a cls_labels.npy dictionary."""
import os

import numpy as np
import PIL.Image

NAMES = ["2007_000032", "2007_000033", "2008_000123", "2009_004567", "2010_000001", "2011_003271"]
SIZES = [(500, 375), (375, 500), (224, 224), (320, 240), (499, 333), (281, 500)]       # (W, H)


def make_voc_tree(root: str, n: int = 6, seed: int = 0):
    from vision_transformer_cam_b200.voc12 import CAT_LIST
    rng = np.random.RandomState(seed)
    for d in ("JPEGImages", "SegmentationClass", "Annotations", "lists"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    labels = {}
    lines = []
    for i in range(n):
        name, (w, h) = NAMES[i], SIZES[i]
        # smooth random image (JPEG-friendly) + noise
        base = rng.rand(h // 16 + 2, w // 16 + 2, 3)
        img = np.kron(base, np.ones((16, 16, 1)))[:h, :w] * 200 + rng.rand(h, w, 3) * 55
        PIL.Image.fromarray(img.astype(np.uint8)).save(os.path.join(root, "JPEGImages", name + ".jpg"), quality=92)
        seg = rng.randint(0, 21, size=(h // 8 + 1, w // 8 + 1)).astype(np.uint8)
        seg = np.kron(seg, np.ones((8, 8), np.uint8))[:h, :w].copy()
        seg[:3] = 255
        im = PIL.Image.fromarray(seg, mode="P")
        im.putpalette([(k * 37 + c * 91) % 256 for k in range(256) for c in range(3)])      # any 256-entry palette keeps the indices
        im.save(os.path.join(root, "SegmentationClass", name + ".png"))
        lab = (rng.rand(20) < 0.12).astype(np.float32)
        lab[rng.randint(20)] = 1.0
        labels[name] = lab
        objs = "".join(f"<object><name>{CAT_LIST[c]}</name></object>" for c in np.nonzero(lab)[0]) + "<object><name>head</name></object>"
        with open(os.path.join(root, "Annotations", name + ".xml"), "w") as f:
            f.write(f"<annotation><filename>{name}.jpg</filename>{objs}</annotation>")
        lines.append(f"/JPEGImages/{name}.jpg /SegmentationClassAug/{name}.png")
    list_path = os.path.join(root, "lists", "val.txt")
    with open(list_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    npy_path = os.path.join(root, "lists", "cls_labels.npy")
    np.save(npy_path, labels, allow_pickle=True)
    return list_path, npy_path, labels
