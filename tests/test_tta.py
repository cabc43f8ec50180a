"""Test-time augmentation: the mapper's transforms on the CPU, the device wrapper against the oracle's
restatement of the merge (tta.py:232-268) on the GPU."""
import numpy as np
import pytest
import torch

from dafne_b200 import tta
from dafne_b200.config import get_cfg


def _cfg(min_sizes=(96, 128), max_size=160, hflip=True, vflip=True):
    cfg = get_cfg()
    cfg.TEST.AUG.MIN_SIZES = list(min_sizes)
    cfg.TEST.AUG.MAX_SIZE = max_size
    cfg.TEST.AUG.HFLIP = hflip
    cfg.TEST.AUG.VFLIP = vflip
    return cfg


def test_resize_shortest_edge_sizes_like_detectron2():
    t = tta.resize_shortest_edge_transform(480, 640, 800, 1333)
    assert (t.new_h, t.new_w) == (800, 1067)
    t = tta.resize_shortest_edge_transform(500, 2000, 800, 1333)  # capped by max_size
    assert (t.new_h, t.new_w) == (333, 1333)
    t = tta.resize_shortest_edge_transform(1024, 1024, 1200, 1200)
    assert (t.new_h, t.new_w) == (1200, 1200)


def test_transforms_agree_with_the_oracle_restatement():
    """Product transforms (dafne_b200/tta.py) against the oracle's independent restatement (oracle/tta.py): output
    sizes of ResizeShortestEdge over a sweep of image sizes, and apply_coords / inverse of every copy's chain, bit for
    bit in float32."""
    from oracle import tta as otta

    rng = np.random.default_rng(1)
    for _ in range(300):
        h, w = int(rng.integers(32, 2400)), int(rng.integers(32, 2400))
        size, max_size = int(rng.choice([400, 512, 600, 800, 1024, 1200, 1333])), int(rng.choice([800, 1333, 2000]))
        t = tta.resize_shortest_edge_transform(h, w, size, max_size)
        assert (t.new_h, t.new_w) == otta.shortest_edge_size(h, w, size, max_size), (h, w, size, max_size)
    cfg = _cfg(min_sizes=(96, 128, 200), max_size=180)
    g = torch.Generator().manual_seed(0)
    image = torch.randint(0, 256, (3, 100, 140), dtype=torch.uint8, generator=g)
    copies = tta.DotaDatasetMapperTTA(cfg)({"image": image, "height": 300, "width": 420})
    chains = [otta.chain_of_copy(100, 140, ms, 180, flip, orig_hw=(300, 420)) for ms in (96, 128, 200)
              for flip in ("", "h", "v")]
    pts = rng.uniform(-5, 260, (64, 2)).astype(np.float32)
    for c, chain in zip(copies, chains):
        tfm = c["transforms"]
        assert np.array_equal(tfm.apply_coords(pts.copy()), otta.apply_coords(chain, pts))
        assert np.array_equal(tfm.inverse().apply_coords(pts.copy()), otta.apply_coords(otta.inverse_chain(chain), pts))


def test_transform_inverse_round_trip_and_device_free_math():
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 300, (50, 2)).astype(np.float32)
    tl = tta.ResizeTransform(300, 400, 150, 260) + tta.TransformList([tta.HFlipTransform(260), tta.VFlipTransform(150)])
    fwd = tl.apply_coords(pts.copy())
    back = tl.inverse().apply_coords(fwd.copy())
    assert np.abs(back - pts).max() < 1e-3
    img = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    out = tl.apply_image(img)
    assert out.shape == (150, 260, 3) and out.dtype == np.uint8
    # flips are exact pixel permutations of the resized image
    res = tta.ResizeTransform(300, 400, 150, 260).apply_image(img)
    assert np.array_equal(out, res[::-1, ::-1])


def test_mapper_builds_the_reference_s_copies():
    cfg = _cfg()
    g = torch.Generator().manual_seed(0)
    image = torch.randint(0, 256, (3, 100, 140), dtype=torch.uint8, generator=g)
    copies = tta.DotaDatasetMapperTTA(cfg)({"image": image, "height": 200, "width": 280})
    assert len(copies) == 2 * 3  # len(MIN_SIZES) x (plain, hflip, vflip)   (tta.py:69-135)
    shapes = [tuple(c["image"].shape) for c in copies]
    assert shapes[0] == shapes[1] == shapes[2] == (3, 96, 134)
    assert shapes[3] == shapes[4] == shapes[5] == (3, 114, 160)  # capped by MAX_SIZE
    for c in copies:
        # the transform chain maps original-image coordinates to this copy's coordinates
        h, w = c["image"].shape[1:]
        corner = c["transforms"].apply_coords(np.array([[280.0, 200.0]], np.float32))
        assert abs(abs(corner[0, 0] - w / 2) - w / 2) < 1e-2 and abs(abs(corner[0, 1] - h / 2) - h / 2) < 1e-2
    assert torch.equal(copies[1]["image"], copies[0]["image"].flip(2))
    assert torch.equal(copies[2]["image"], copies[0]["image"].flip(1))
    with pytest.raises(NotImplementedError):
        cfg.TEST.AUG.ROTATION_ANGLES = [90]
        tta.DotaDatasetMapperTTA(cfg)


@pytest.mark.gpu
def test_tta_wrapper_equals_oracle_merge_of_the_copies():
    from dafne_b200.modeling import build_model
    from oracle import tta as otta

    cfg = _cfg(min_sizes=(160, 192, 256), max_size=320)
    cfg.MODEL.DEVICE = "cuda:0"
    model = build_model(cfg)
    wrapper = tta.OneStageRCNNWithTTA(cfg, model)
    g = torch.Generator().manual_seed(4)
    image = torch.randint(0, 256, (3, 200, 256), dtype=torch.uint8, generator=g)
    inp = {"image": image, "height": 400, "width": 512}
    out = wrapper([inp])[0]["instances"]
    # the oracle merges what the SAME model returns for each augmented copy
    copies = wrapper.tta_mapper(inp)
    tfms = [c.pop("transforms") for c in copies]
    per_copy = wrapper._batch_inference(copies)
    corners = [o["instances"].pred_corners.cpu().numpy() for o in per_copy]
    scores = [o["instances"].scores.cpu().numpy() for o in per_copy]
    classes = [o["instances"].pred_classes.cpu().numpy() for o in per_copy]
    assert sum(len(s) for s in scores) > 50, "the synthetic model should detect something in every copy"
    spec = model.spec
    # the oracle applies ITS OWN restatement of the transforms: the chains are rebuilt from the mapper's configuration
    # (sizes and flips), not taken from the product's transform objects
    H, W = image.shape[1:]
    chains = [otta.chain_of_copy(H, W, ms, 320, flip, orig_hw=(400, 512)) for ms in (160, 192, 256)
              for flip in ("", "h", "v")]
    assert len(chains) == len(tfms)
    want_c, want_s, want_k, _ = otta.merge_detections(corners, scores, classes, chains, spec.nms_thresh,
                                                      spec.post_nms_topk, spec.vehicle_merge)
    assert len(out) == len(want_s)
    assert np.array_equal(out.scores.cpu().numpy(), want_s)
    assert np.array_equal(out.pred_classes.cpu().numpy(), want_k)
    assert np.array_equal(out.pred_corners.cpu().numpy(), want_c)
    # corners are in the coordinates of the ORIGINAL image (height / width), not of the input tensor
    assert out.pred_corners[:, 0::2].max() > 256


@pytest.mark.gpu
def test_tta_cross_image_batches_equal_the_per_image_loop():
    """A call with several images batches same-shaped copies ACROSS images (cross_image_batch, not in the reference): the
    merged detections of every image are bit-identical to the reference-shaped loop over the images (tta.py:178-196)."""
    from dafne_b200.modeling import build_model

    cfg = _cfg(min_sizes=(160, 192, 256), max_size=320)
    cfg.MODEL.DEVICE = "cuda:0"
    model = build_model(cfg)
    g = torch.Generator().manual_seed(9)
    # three images of one size and one of another: its copies form groups of their own
    inputs = [{"image": torch.randint(0, 256, (3, 200, 256), dtype=torch.uint8, generator=g), "height": 400, "width": 512}
              for _ in range(3)]
    inputs.append({"image": torch.randint(0, 256, (3, 224, 224), dtype=torch.uint8, generator=g), "height": 224,
                   "width": 224})
    loop = tta.OneStageRCNNWithTTA(cfg, model, cross_image_batch=0)
    batched = tta.OneStageRCNNWithTTA(cfg, model, cross_image_batch=5)  # 9 copies per shape of the first three: 5 + 4
    want = loop(inputs)
    for rep in range(2):  # the second call replays the captured graphs
        got = batched(inputs)
        assert len(got) == len(want) == 4
        for a, b in zip(want, got):
            ia, ib = a["instances"], b["instances"]
            assert len(ia) == len(ib) and len(ia) > 20
            assert torch.equal(ia.pred_corners, ib.pred_corners)
            assert torch.equal(ia.scores, ib.scores)
            assert torch.equal(ia.pred_classes, ib.pred_classes)
            assert ia.image_size == ib.image_size
    assert not torch.equal(want[0]["instances"].scores, want[1]["instances"].scores)


@pytest.mark.gpu
def test_poly_nms_beyond_the_shared_memory_sort():
    """The TTA union can hold 27 copies x 1000 boxes: more than the 16 384 keys one CTA sorts in shared memory."""
    from dafne_b200.modeling import batched_nms_poly
    from oracle import postprocess as opost

    rng = np.random.default_rng(2)
    n = 20000
    cx, cy = rng.uniform(0, 1500, n), rng.uniform(0, 1500, n)
    w, h, a = rng.uniform(20, 60, n), rng.uniform(8, 20, n), rng.uniform(0, np.pi, n)
    dx = np.stack([-w, w, w, -w], 1) / 2
    dy = np.stack([-h, -h, h, h], 1) / 2
    boxes = np.stack([cx[:, None] + dx * np.cos(a)[:, None] - dy * np.sin(a)[:, None],
                      cy[:, None] + dx * np.sin(a)[:, None] + dy * np.cos(a)[:, None]], 2).reshape(n, 8).astype(np.float32)
    scores = (rng.permutation(n) / n).astype(np.float32)
    classes = rng.integers(0, 3, n)
    keep = batched_nms_poly(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(),
                            torch.from_numpy(classes).cuda(), 0.1).cpu().numpy()
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    span = (boxes.max() - boxes.min()) + np.float32(1.0)
    shifted = boxes + (classes.astype(np.float32) * span)[:, None]
    want = order[opost.greedy_nms(shifted[order], 0.1)]
    assert np.array_equal(keep, want)
