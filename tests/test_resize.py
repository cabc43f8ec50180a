"""Input-side bilinear resize: the oracle against Pillow (what detectron2's ResizeTransform calls), the
device kernel against both, and the TTA mapper / predictor paths that use it."""
import numpy as np
import pytest
import torch

from oracle import resize as oresize

SIZES = [  # (h, w, new_h, new_w): up, down, mixed, one axis unchanged, strong down-scaling, odd sizes
    (100, 140, 96, 134), (37, 53, 111, 80), (256, 256, 100, 300), (64, 64, 64, 100), (64, 100, 32, 100),
    (333, 500, 800, 1201), (600, 600, 37, 41), (50, 70, 50, 70), (1, 9, 5, 3), (128, 160, 800, 1000),
]


def _pil(img_chw, nh, nw):
    from PIL import Image

    hwc = np.ascontiguousarray(img_chw.transpose(1, 2, 0))
    return np.asarray(Image.fromarray(hwc).resize((nw, nh), Image.BILINEAR)).transpose(2, 0, 1)


@pytest.mark.parametrize("h,w,nh,nw", SIZES)
def test_oracle_matches_pillow(h, w, nh, nw):
    pytest.importorskip("PIL")
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    assert np.array_equal(oresize.resize_bilinear_u8(img, nh, nw), _pil(img, nh, nw))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,nh,nw", SIZES)
def test_device_matches_oracle_and_pillow(h, w, nh, nw):
    from dafne_b200.modeling import resize_bilinear_u8

    rng = np.random.default_rng(h * 1000 + w + 1)
    img = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    img[0, : h // 2] = 255  # saturated and constant regions: the fixed-point rounding must reproduce them exactly
    img[1, :, : w // 3] = 0
    got = resize_bilinear_u8(torch.from_numpy(img).cuda(), nh, nw).cpu().numpy()
    assert np.array_equal(got, oresize.resize_bilinear_u8(img, nh, nw))
    try:
        assert np.array_equal(got, _pil(img, nh, nw))
    except ImportError:
        pass


@pytest.mark.gpu
def test_device_resize_batched_planes():
    from dafne_b200.modeling import resize_bilinear_u8

    rng = np.random.default_rng(5)
    batch = rng.integers(0, 256, (4, 3, 90, 120), dtype=np.uint8)
    got = resize_bilinear_u8(torch.from_numpy(batch).cuda(), 144, 100).cpu().numpy()
    for i in range(4):
        assert np.array_equal(got[i], oresize.resize_bilinear_u8(batch[i], 144, 100))


@pytest.mark.gpu
def test_tta_mapper_device_copies_equal_host_copies():
    from dafne_b200 import tta
    from dafne_b200.config import get_cfg

    cfg = get_cfg()
    cfg.TEST.AUG.MIN_SIZES = [96, 128, 200]
    cfg.TEST.AUG.MAX_SIZE = 240
    g = torch.Generator().manual_seed(0)
    image = torch.randint(0, 256, (3, 100, 140), dtype=torch.uint8, generator=g)
    inp = {"image": image, "height": 200, "width": 280}
    host = tta.DotaDatasetMapperTTA(cfg)(inp)
    dev = tta.DotaDatasetMapperTTA(cfg, device="cuda:0")(inp)
    assert len(host) == len(dev) == 9
    for a, b in zip(host, dev):
        assert b["image"].is_cuda and torch.equal(a["image"], b["image"].cpu())
        pts = np.array([[10.0, 20.0], [279.0, 199.0]], np.float32)
        assert np.array_equal(a["transforms"].apply_coords(pts.copy()), b["transforms"].apply_coords(pts.copy()))


@pytest.mark.gpu
def test_default_predictor_resizes_like_detectron2():
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import DefaultPredictor, resize_bilinear_u8

    cfg = get_cfg()
    cfg.MODEL.DEVICE = "cuda:0"
    cfg.INPUT.MIN_SIZE_TEST, cfg.INPUT.MAX_SIZE_TEST = 192, 256
    pred = DefaultPredictor(cfg)
    img = np.random.default_rng(0).integers(0, 256, (96, 160, 3), dtype=np.uint8)
    out = pred(img)["instances"]
    assert out.image_size == (96, 160)  # results are in the coordinates of the original image
    # the same as resizing by hand (ResizeShortestEdge: 96 x 160 -> 154 x 256, capped by MAX_SIZE_TEST) and asking
    # the model to rescale to the original size
    chw = torch.as_tensor(np.ascontiguousarray(img.transpose(2, 0, 1))).cuda()
    resized = resize_bilinear_u8(chw, 154, 256)
    want = pred.model([{"image": resized, "height": 96, "width": 160}])[0]["instances"]
    assert len(out) == len(want) and torch.equal(out.pred_corners, want.pred_corners)
