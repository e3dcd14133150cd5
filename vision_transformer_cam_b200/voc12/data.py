"""VOC12 dataset classes with the reference's names and return conventions (voc12/data.py:27-118), plus the uint8 ingest
the fused forward wants: decode workers produce resized uint8 HWC images, batches are collated straight into pinned
uint8 [B,S,S,3] tensors, and ToTensor + Normalize (predict.py:72-75, validate.py:80-84) happen inside the patchify kernel
(`VisionTransformer.forward_cam_u8`) -- 4x fewer host->device bytes than the fp32 NCHW tensor of the reference, and
bit-identical patch values.

    loader = make_u8_loader("voc12/val.txt", voc_root, batch_size=256, seg_label_flag=False)
    for names, u8, labels in loader:                         # u8: pinned uint8 [B,224,224,3]
        out = model.forward_cam_u8(u8.to("cuda", non_blocking=True))

Host-side only (PIL / numpy / torch.utils.data); nothing here touches the GPU."""
from __future__ import annotations

import os.path
from typing import List, Optional, Sequence

import numpy as np
import PIL.Image
import torch
from torch.utils.data import DataLoader, Dataset

IMG_FOLDER_NAME = "JPEGImages"
SEG_LABEL_FOLDER_NAME = "SegmentationClass"
ANNOT_FOLDER_NAME = "Annotations"

CAT_LIST = ['aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable', 'dog',
            'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor']
CAT_NAME_TO_NUM = {name: i for i, name in enumerate(CAT_LIST)}


def load_image_label_from_xml(img_name: str, voc12_root: str) -> np.ndarray:
    """Multi-hot float32[20] from the <name> tags of Annotations/<img_name>.xml (voc12/data.py:27-41)."""
    from xml.dom import minidom
    lab = np.zeros(len(CAT_LIST), np.float32)
    for el in minidom.parse(os.path.join(voc12_root, ANNOT_FOLDER_NAME, img_name + ".xml")).getElementsByTagName("name"):
        num = CAT_NAME_TO_NUM.get(el.firstChild.data)
        if num is not None:
            lab[num] = 1.0
    return lab


def load_image_label_list_from_xml(img_name_list: Sequence[str], voc12_root: str) -> List[np.ndarray]:
    return [load_image_label_from_xml(n, voc12_root) for n in img_name_list]


def load_image_label_list_from_npy(img_name_list: Sequence[str], npy_path: str = "voc12/cls_labels.npy") -> List[np.ndarray]:
    """Labels from the {name: multi-hot} dictionary file (voc12/data.py:49-53; the reference hard-codes the relative path)."""
    table = np.load(npy_path, allow_pickle=True).item()
    return [table[n] for n in img_name_list]


def get_img_path(img_name: str, voc12_root: str) -> str:
    return os.path.join(voc12_root, IMG_FOLDER_NAME, img_name + ".jpg")


def get_seg_label_path(img_name: str, voc12_root: str) -> str:
    return os.path.join(voc12_root, SEG_LABEL_FOLDER_NAME, img_name + ".png")


def load_img_name_list(dataset_path: str) -> List[str]:
    """List-file lines are '/JPEGImages/2007_000032.jpg /SegmentationClassAug/2007_000032.png' or bare names; the image
    name is the 11 characters in front of the first entry's extension (voc12/data.py:64-70)."""
    with open(dataset_path) as f:
        return [line.split(" ")[0][-15:-4] for line in f.read().splitlines() if line]


class VOC12ImageDataset(Dataset):
    """(name, img[, seg_label]) per item (voc12/data.py:73-100): RGB PIL image through `transform`; seg_label int64 [H,W]."""

    def __init__(self, img_name_list_path: str, voc12_root: str, transform=None, seg_label_flag: bool = False):
        self.img_name_list = load_img_name_list(img_name_list_path)
        self.voc12_root = voc12_root
        self.transform = transform
        self.seg_label_flag = seg_label_flag

    def __len__(self) -> int:
        return len(self.img_name_list)

    def __getitem__(self, idx: int):
        name = self.img_name_list[idx]
        img = PIL.Image.open(get_img_path(name, self.voc12_root)).convert("RGB")
        if self.transform:
            img = self.transform(img)
        if self.seg_label_flag:
            seg = torch.as_tensor(np.array(PIL.Image.open(get_seg_label_path(name, self.voc12_root))), dtype=torch.int64)
            return name, img, seg
        return name, img


class VOC12ClsDataset(VOC12ImageDataset):
    """(name, img, label[, seg_label]) with label float32[20] (voc12/data.py:103-118)."""

    def __init__(self, img_name_list_path: str, voc12_root: str, transform=None, seg_label_flag: bool = False,
                 cls_labels_path: str = "voc12/cls_labels.npy", labels_from_xml: bool = False):
        super().__init__(img_name_list_path, voc12_root, transform, seg_label_flag)
        self.label_list = (load_image_label_list_from_xml(self.img_name_list, voc12_root) if labels_from_xml
                           else load_image_label_list_from_npy(self.img_name_list, cls_labels_path))

    def __getitem__(self, idx: int):
        item = super().__getitem__(idx)
        label = torch.from_numpy(np.asarray(self.label_list[idx], dtype=np.float32))
        if self.seg_label_flag:
            name, img, seg = item
            return name, img, label, seg
        name, img = item
        return name, img, label


class U8Resize:
    """transforms.Resize([S, S]) of the reference pipelines, stopping BEFORE ToTensor / Normalize: PIL bilinear resize
    (what torchvision applies to a PIL image) -> uint8 HWC tensor.  The two remaining fp32 operations run on the GPU inside
    the patchify kernel, with the same rounding as torchvision's."""

    def __init__(self, size: int = 224):
        self.size = int(size)

    def __call__(self, img: PIL.Image.Image) -> torch.Tensor:
        if img.size != (self.size, self.size):
            img = img.resize((self.size, self.size), PIL.Image.BILINEAR)
        return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())


def u8_collate(items):
    """(name, u8 img[, label][, seg]) items -> (names, uint8 [B,S,S,3], [labels [B,C]], [list of seg label maps]).
    Segmentation maps keep their own sizes (VOC images differ), so they stay a list; the image batch is one contiguous
    tensor the DataLoader can pin."""
    names = [it[0] for it in items]
    imgs = torch.stack([it[1] for it in items])
    out = [names, imgs]
    rest = list(zip(*[it[2:] for it in items]))
    for col in rest:
        if col[0].dim() == 1:
            out.append(torch.stack(col))
        else:
            out.append(list(col))
    return tuple(out)


def make_u8_loader(img_name_list_path: str, voc12_root: str, batch_size: int = 256, img_size: int = 224, seg_label_flag: bool = False,
                   with_labels: bool = True, cls_labels_path: str = "voc12/cls_labels.npy", labels_from_xml: bool = False,
                   num_workers: Optional[int] = None, shuffle: bool = False, rank: int = 0, world: int = 1,
                   pin_memory: Optional[bool] = None) -> DataLoader:
    """DataLoader over VOC12 that yields pinned uint8 HWC batches for `forward_cam_u8`.  `rank` / `world` select this
    process's contiguous shard of the list (the partition of dist.shard_range), so N ranks read disjoint images."""
    from .. import dist as D
    if with_labels:
        ds: Dataset = VOC12ClsDataset(img_name_list_path, voc12_root, U8Resize(img_size), seg_label_flag, cls_labels_path, labels_from_xml)
    else:
        ds = VOC12ImageDataset(img_name_list_path, voc12_root, U8Resize(img_size), seg_label_flag)
    if world > 1:
        lo, hi = D.shard_range(len(ds), rank, world)
        ds = torch.utils.data.Subset(ds, range(lo, hi))
    if num_workers is None:
        num_workers = min(os.cpu_count() or 1, 16)
    if pin_memory is None:
        pin_memory = torch.cuda.is_available()
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=pin_memory, collate_fn=u8_collate,
                      persistent_workers=False, drop_last=False)
