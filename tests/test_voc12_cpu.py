"""CPU: the VOC12 ingest (vision_transformer_cam_b200/voc12) against the reference's conventions (voc12/data.py:27-118)
and against torchvision's Resize on the same decoded images."""
import numpy as np
import PIL.Image
import pytest
import torch

from voc_fixture import NAMES, SIZES, make_voc_tree


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("voc"))
    list_path, npy_path, labels = make_voc_tree(root)
    return dict(root=root, list=list_path, npy=npy_path, labels=labels)


def test_name_list_and_paths(tree, tmp_path):
    from vision_transformer_cam_b200 import voc12
    assert voc12.load_img_name_list(tree["list"]) == NAMES
    bare = tmp_path / "bare.txt"                      # train.txt style: the [-15:-4] slice expects an extension
    bare.write_text("\n".join(n + ".jpg" for n in NAMES) + "\n")
    assert voc12.load_img_name_list(str(bare)) == NAMES
    assert voc12.get_img_path("2007_000032", "/d").endswith("JPEGImages/2007_000032.jpg")
    assert voc12.get_seg_label_path("2007_000032", "/d").endswith("SegmentationClass/2007_000032.png")
    assert len(voc12.CAT_LIST) == 20 and voc12.CAT_NAME_TO_NUM["tvmonitor"] == 19


def test_labels_npy_equals_xml(tree):
    from vision_transformer_cam_b200 import voc12
    a = voc12.load_image_label_list_from_npy(NAMES, tree["npy"])
    b = voc12.load_image_label_list_from_xml(NAMES, tree["root"])          # 'head' (not a VOC class) is ignored
    for x, y, n in zip(a, b, NAMES):
        assert np.array_equal(x, y) and np.array_equal(x, tree["labels"][n]) and y.dtype == np.float32


def test_dataset_return_conventions(tree):
    from vision_transformer_cam_b200 import voc12
    ds = voc12.VOC12ImageDataset(tree["list"], tree["root"])
    name, img = ds[1]
    assert name == NAMES[1] and isinstance(img, PIL.Image.Image) and img.size == SIZES[1] and img.mode == "RGB" and len(ds) == 6
    ds = voc12.VOC12ClsDataset(tree["list"], tree["root"], seg_label_flag=True, cls_labels_path=tree["npy"])
    name, img, label, seg = ds[0]
    assert label.dtype == torch.float32 and label.shape == (20,) and seg.dtype == torch.int64 and seg.shape == (375, 500)
    assert int(seg[0, 0]) == 255 and int(seg[3:].max()) <= 20
    name, img, label = voc12.VOC12ClsDataset(tree["list"], tree["root"], cls_labels_path=tree["npy"])[2]
    assert torch.equal(label, torch.from_numpy(tree["labels"][NAMES[2]]))


def test_u8_resize_equals_torchvision(tree):
    """U8Resize == the uint8 image torchvision's Resize([224,224]) hands to ToTensor (validate.py:80-84), so that
    Normalize(ToTensor(.)) of the reference == the in-kernel normalisation of forward_cam_u8 on the same bytes."""
    from torchvision import transforms
    from vision_transformer_cam_b200 import voc12
    ref_t = transforms.Compose([transforms.Resize([224, 224]), transforms.PILToTensor()])
    for n in NAMES:
        img = PIL.Image.open(voc12.get_img_path(n, tree["root"])).convert("RGB")
        u8 = voc12.U8Resize(224)(img)
        assert u8.dtype == torch.uint8 and u8.shape == (224, 224, 3)
        assert torch.equal(u8.permute(2, 0, 1), ref_t(img))


def test_u8_loader_batches_and_shards(tree):
    from vision_transformer_cam_b200 import voc12
    loader = voc12.make_u8_loader(tree["list"], tree["root"], batch_size=4, seg_label_flag=True, cls_labels_path=tree["npy"], num_workers=2)
    batches = list(loader)
    assert [b[1].shape for b in batches] == [(4, 224, 224, 3), (2, 224, 224, 3)]
    names, u8, labels, segs = batches[0]
    assert names == NAMES[:4] and u8.dtype == torch.uint8 and u8.is_contiguous() and labels.shape == (4, 20)
    assert [tuple(s.shape) for s in segs] == [(h, w) for (w, h) in SIZES[:4]]
    # two ranks read disjoint contiguous shards that cover the list
    seen = []
    for r in range(2):
        for b in voc12.make_u8_loader(tree["list"], tree["root"], batch_size=8, with_labels=False, num_workers=0, rank=r, world=2):
            assert len(b) == 2
            seen += b[0]
    assert seen == NAMES
