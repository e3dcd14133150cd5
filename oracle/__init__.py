"""CPU oracle for the ViT forward + CAM hot path of Jingfeng-Tang/vision_transformer_cam.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package
(`vision_transformer_cam_b200/`) imports this directory.  The only legal
importers are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and there only as the checker or as the
timed CPU baseline -- never as the product path.

The oracle is a plain PyTorch fp32 restatement (the path is floating point; the
reference itself is eager PyTorch) of

  * `vit_model.py:303-424`  VisionTransformer.forward_features / forward
  * `vit_model.py:103-140`  Attention.forward (incl. the layer>4 additive mask)
  * `vit_model.py:189-200`  Block.forward
  * `predict.py:189-247,261-269`   attention rollout + per-layer CLS maps
  * `validate.py:132-258`          high-weight-patch pseudo segmentation + bg map
  * `utils.py:30-88,248-262`       ConfusionMatrix, cam_norm, compute_mAP
  * `t.py:55-75`                   classic CAM definition

generalised only in `197 -> N` and `12 -> num_heads` (the reference hard-codes
those, so configs 4/5 of BASELINE.json have no runnable reference).

Parity pin: the reference ships no golden vectors or tests, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the
build container through `oracle/ref_shim.py`; the generating script is
`tests/golden/make_golden.py` and the fixtures live in `tests/golden/*.npz`.
"""
from .vit_forward import VitConfig, init_state_dict, forward, make_images, peaked  # noqa: F401
from . import postproc  # noqa: F401
