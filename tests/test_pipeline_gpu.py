"""GPU: the predict / validate callers (pipeline.py) against the CPU oracle, and the sharded extraction (1 rank here;
the 2-rank NCCL run is in test_multigpu.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(lib_built):
    import vision_transformer_cam_b200 as V
    from oracle import vit_forward as VF
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = V.vit_base_patch16_224_in21k(num_classes=20, has_logits=False)
    sd = VF.peaked({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(sd)
    return dict(model=model.to(dev).eval(), sd=sd, VF=VF, dev=dev)


def cosine(a, b):
    a, b = a.double().cpu().flatten(), torch.as_tensor(b).double().flatten()
    return float(a @ b / (a.norm() * b.norm()))


def test_predict_matches_oracle(env):
    from vision_transformer_cam_b200 import pipeline
    from oracle import postproc as PP
    x = env["VF"].make_images(3, 2)
    ref = env["VF"].forward(env["sd"], x, env["VF"].VIT_B16_224, mask_norm="batch")
    p = pipeline.predict(env["model"], x.to(env["dev"]), (375, 500))
    assert p.rollout.shape == (2, 375, 500) and p.layer_maps.shape == (12, 2, 375, 500) and p.layer_maps.dtype == torch.uint8
    # free-running decisions in the peaked regime: a flipped background bit changes later layers (graded with teacher
    # forcing in test_forward_gpu.py); here: layers before the first mask exactly, the rest loosely
    c_roll = cosine(p.rollout, PP.rollout_map(ref["P"], (375, 500)))
    c_lm = cosine(p.layer_maps[:5].float(), PP.layer_maps(ref["P"], (375, 500), as_u8=True)[:5].float())
    c_cam = cosine(p.cam, PP.classic_cam(ref["X"][-1], env["sd"]["head1.weight"]))
    print("predict (peaked, free running): rollout", c_roll, "layer maps 0-4", c_lm, "cam", c_cam)
    assert c_lm >= 0.999 and c_roll >= 0.95 and c_cam >= 0.95
    assert float((p.hwp_scores.cpu() - torch.sigmoid(ref["hwp"])).abs().max()) < 0.3


def test_validator_counters(env):
    from vision_transformer_cam_b200 import pipeline
    from oracle import postproc as PP
    x = env["VF"].make_images(20, 4).to(env["dev"])
    g = torch.Generator().manual_seed(9)
    seg_gt = torch.randint(0, 21, (4, 120, 160), generator=g).to(torch.uint8)
    seg_gt[:, :5] = 255
    target = (torch.rand((4, 20), generator=g) < 0.2).float()
    target[:, 3] = 1
    v = pipeline.Validator(env["model"], 20, device=env["dev"])
    seg = torch.cat([v.step(x[:2], target[:2], seg_gt[:2]), v.step(x[2:], target[2:], seg_gt[2:])])
    res = v.finalize()
    ref = PP.confmat_update(None, seg_gt.numpy(), seg.cpu().numpy())
    assert np.array_equal(res["confmat"].cpu().numpy(), ref)                           # integer counters: bit exact
    assert res["confmat"].sum().item() == 4 * 115 * 160
    assert 0.0 <= res["mAP"] <= 1.0 and seg.shape == (4, 120, 160)
    # per-image independence of the validate path (mask_norm='image'): one call of 4 == two calls of 2
    v2 = pipeline.Validator(env["model"], 20, device=env["dev"])
    assert torch.equal(v2.step(x, target, seg_gt), seg)


def test_extract_cams_single_rank(env):
    from vision_transformer_cam_b200 import pipeline
    VF, dev = env["VF"], env["dev"]
    get = lambda lo, hi: VF.make_images(100 + lo, hi - lo).to(dev)
    out = pipeline.extract_cams_sharded(env["model"], get, n_items=7, batch=3)
    assert out["cam"].shape == (7, 20, 14, 14) and out["rollout"].shape == (7, 196) and out["hwp_logits"].shape == (7, 20)
    assert bool(torch.isfinite(out["cam"]).all()) and float(out["cam"].max()) <= 1.0 + 1e-6


def test_device_feeder_overlapped_ingest_equals_direct(env):
    """Pinned-host batches streamed through DeviceFeeder (copy of batch i+1 on a side stream during batch i) give exactly
    the results of device-resident batches, in order, for more batches than buffers."""
    from vision_transformer_cam_b200 import pipeline as PIPE
    VF, dev, model = env["VF"], env["dev"], env["model"]
    host = [VF.make_images(4 * i, 4).pin_memory() for i in range(5)]
    direct = [model.forward_cam(h.to(dev), mask_norm="image").logits.clone() for h in host]
    fed = []
    for x in PIPE.DeviceFeeder(dev).stream(iter(host)):
        assert x.is_cuda
        fed.append(model.forward_cam(x, mask_norm="image").logits.clone())
    assert len(fed) == 5 and all(torch.equal(a, b) for a, b in zip(fed, direct))
    mixed = list(PIPE.DeviceFeeder(dev).stream([host[0], host[1].to(dev), host[2]]))
    assert len(mixed) == 3 and torch.equal(mixed[1].cpu(), host[1])
    out = PIPE.extract_cams_sharded(model, lambda lo, hi: torch.cat(host)[lo:hi].pin_memory(), 20, batch=6, with_rollout=False)
    ref = PIPE.extract_cams_sharded(model, lambda lo, hi: torch.cat(host)[lo:hi].to(dev), 20, batch=6, with_rollout=False)
    assert torch.equal(out["cam"], ref["cam"]) and out["cam"].shape[0] == 20
