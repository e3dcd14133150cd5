"""VOC12 ingest for the fused forward (mirrors the reference's voc12/data.py surface)."""
from .data import (CAT_LIST, CAT_NAME_TO_NUM, IMG_FOLDER_NAME, SEG_LABEL_FOLDER_NAME, ANNOT_FOLDER_NAME,  # noqa: F401
                   load_image_label_from_xml, load_image_label_list_from_xml, load_image_label_list_from_npy,
                   get_img_path, get_seg_label_path, load_img_name_list, VOC12ImageDataset, VOC12ClsDataset,
                   U8Resize, u8_collate, make_u8_loader)
