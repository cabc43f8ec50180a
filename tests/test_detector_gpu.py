"""GPU tests of the reference-facing surface: OneStageDetector(cfg)(batched_inputs), DefaultPredictor, detect_host."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import build_model

    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(ROOT, "configs", "dota10_r50_1024.yaml"))
    cfg.MODEL.DEVICE = "cuda:0"
    return build_model(cfg)


def _inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [{"image": torch.randint(0, 256, (3, 256, 320), dtype=torch.uint8, generator=g), "height": 512, "width": 640},
            {"image": torch.randint(0, 256, (3, 200, 300), dtype=torch.uint8, generator=g)}]


def test_forward_returns_instances_with_reference_fields(model):
    out = model(_inputs())
    assert len(out) == 2
    for o, size in zip(out, [(512, 640), (200, 300)]):
        inst = o["instances"]
        assert inst.image_size == size
        n = len(inst)
        assert inst.pred_boxes.tensor.shape == (n, 4) and inst.pred_corners.shape == (n, 8)
        assert inst.scores.shape == (n,) and inst.centerness.shape == (n,) and inst.locations.shape == (n, 2)
        assert inst.pred_classes.dtype == torch.int64 and inst.fpn_levels.dtype == torch.int64
        s = inst.scores
        assert bool((s[:-1] >= s[1:]).all()) and n > 0 and n <= 1000


def test_batch_result_equals_padded_single_semantics(model):
    """Reference semantics for mixed sizes: zero-pad to the batch maximum; GN statistics see the padding. Running the
    small image alone therefore differs slightly from running it in the batch, but a batch of one reproduces itself."""
    a = model(_inputs(3)[:1])[0]["instances"]
    b = model(_inputs(3)[:1])[0]["instances"]
    assert torch.equal(a.pred_classes, b.pred_classes) and torch.equal(a.pred_corners, b.pred_corners)


def test_float_images_and_uint8_images_agree(model):
    inp = _inputs(5)[:1]
    a = model(inp)[0]["instances"]
    inp_f = [{**inp[0], "image": inp[0]["image"].float()}]
    b = model(inp_f)[0]["instances"]
    assert torch.equal(a.pred_classes, b.pred_classes) and torch.equal(a.scores, b.scores)


def test_do_postprocess_false_then_manual_rescale(model):
    """detectron2's ProposalNetwork.forward runs detector_postprocess on every result: with do_postprocess=False the
    boxes are still scaled / clipped / filtered and image_size is the output size; only the corner / location rescale
    (one_stage_detector.py:78-98) is left to the caller -- the TTA call shape (tta.py:190-194)."""
    inp = _inputs(7)[:1]
    raw = model.inference(inp, do_postprocess=False)
    inst = raw[0]["instances"]
    assert inst.image_size == (512, 640)
    full = model(inp)[0]["instances"]
    assert len(inst) == len(full)
    assert torch.equal(inst.pred_boxes.tensor, full.pred_boxes.tensor)  # already in output coordinates
    assert torch.equal(inst.scores, full.scores) and torch.equal(inst.pred_classes, full.pred_classes)
    assert not torch.equal(inst.pred_corners, full.pred_corners)  # still in input coordinates
    model._postprocess(raw, inp)  # in place, like the reference's
    assert torch.equal(inst.pred_corners, full.pred_corners)
    assert torch.equal(inst.locations, full.locations)


def test_select_over_all_levels_like_tta(model):
    """tta.py:264-268: concatenate instances of several runs, then NMS + top-k over the union."""
    from dafne_b200.structures import Instances

    inp = _inputs(9)[:1]
    a = model.inference(inp, do_postprocess=False)[0]["instances"]
    merged = Instances.cat([a, a])  # duplicates must collapse back to the original set
    out = model.proposal_generator.dafne_outputs.select_over_all_levels([merged])[0]
    assert len(out) == len(a)
    assert torch.equal(out.scores, a.scores)


def test_default_predictor_call_shape(model):
    from dafne_b200.modeling import DefaultPredictor

    pred = DefaultPredictor(model.cfg)
    img = np.random.default_rng(0).integers(0, 256, (128, 160, 3), dtype=np.uint8)
    out = pred(img)
    assert "instances" in out and out["instances"].image_size == (128, 160)


def test_detect_host_equals_device_path(model):
    eng = model._get_engine()
    g = torch.Generator().manual_seed(13)
    batch = torch.randint(0, 256, (2, 3, 256, 256), dtype=torch.uint8, generator=g)
    sizes = [(256, 256), (256, 256)]
    dets_d, counts_d = eng.detect(batch.cuda(), sizes)
    hd, hc = eng.detect_host(batch.pin_memory(), sizes)
    torch.cuda.synchronize()
    assert torch.equal(hc, counts_d.cpu())
    n = int(hc[0])
    assert torch.equal(hd[0, :n], dets_d[0, :n].cpu())


def test_pipelined_host_calls_equal_blocking_call(model):
    """dafne_detect_host_begin / _end (two batches in flight, H2D on the copy stream) returns what dafne_detect_host
    returns, batch by batch, and refuses a third batch in flight."""
    from dafne_b200._capi import DafneError
    from dafne_b200.engine import DET

    eng = model._engine
    g = torch.Generator().manual_seed(11)
    batches = [torch.randint(0, 256, (2, 3, 128, 160), dtype=torch.uint8, generator=g).pin_memory() for _ in range(4)]
    sizes = [(128, 160), (100, 150)]
    cap = 1064
    want = []
    for b in batches:
        d, c = eng.detect_host(b, sizes, None, None, None, cap)
        want.append((d.clone(), c.clone()))
    bufs = [(torch.empty(2, cap, DET).pin_memory(), torch.empty(2, dtype=torch.int32).pin_memory()) for _ in range(2)]
    prev = None
    for i, b in enumerate(batches):
        t = eng.detect_host_begin(b, sizes, None, bufs[i % 2][0], bufs[i % 2][1], cap)
        if prev is not None:
            eng.detect_host_end(prev[0])
            k = prev[1]
            n = want[k][1]
            assert torch.equal(bufs[k % 2][1], n)
            for j in range(2):
                assert torch.equal(bufs[k % 2][0][j, : n[j]], want[k][0][j, : n[j]])
        prev = (t, i)
    t_extra = eng.detect_host_begin(batches[0], sizes, None, bufs[0][0], bufs[0][1], cap)
    with pytest.raises(DafneError):  # slots of batch 3 and of the extra batch are both in flight
        eng.detect_host_begin(batches[1], sizes, None, bufs[1][0], bufs[1][1], cap)
    eng.detect_host_end(prev[0])
    eng.detect_host_end(t_extra)
    torch.cuda.synchronize()


def test_unsupported_config_fails_loudly():
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import build_model

    cfg = get_cfg()
    cfg.MODEL.DAFNE.CORNER_PREDICTION = "direct"
    with pytest.raises(NotImplementedError):
        build_model(cfg)


def test_graph_replay_equals_eager_step(model):
    """dafne_graph_capture / _launch: the step replayed as one CUDA graph returns the eager step's bits, and follows new
    image contents written into the captured input buffer."""
    from dafne_b200.engine import DafneEngine, DetectionWire

    eng = DafneEngine(model.spec, torch.device("cuda:0"))
    eng.load_state_dict(model.state_dict())
    g = torch.Generator().manual_seed(21)
    a = torch.randint(0, 256, (2, 3, 192, 256), dtype=torch.uint8, generator=g).cuda()
    b = torch.randint(0, 256, (2, 3, 192, 256), dtype=torch.uint8, generator=g).cuda()
    sizes = [(192, 256), (180, 250)]
    want_a = [t.clone() for t in eng.detect(a, sizes)]
    want_b = [t.clone() for t in eng.detect(b, sizes)]
    buf = a.clone()
    wire = DetectionWire(2, model.spec.post_nms_topk + 64, torch.device("cuda:0"))
    dets, counts = eng.capture(buf, sizes, out=wire)
    eng.replay()
    torch.cuda.synchronize()
    assert torch.equal(counts, want_a[1]) and torch.equal(dets, want_a[0])
    buf.copy_(b)
    eng.replay()
    torch.cuda.synchronize()
    assert torch.equal(counts, want_b[1]) and torch.equal(dets, want_b[0])
    assert int(counts.min()) > 0
    launches, _ = eng.stats()
    assert launches > 4 * 100  # two eager steps, the capture's eager step and two replays are all accounted for
    eng.close()


def test_tta_deferred_batches_equal_the_per_batch_path(model):
    """OneStageRCNNWithTTA enqueues all batches before its one host sync: same outputs as the reference-shaped
    `_batch_inference` (one sync per batch)."""
    from dafne_b200 import tta

    cfg = model.cfg.clone() if hasattr(model.cfg, "clone") else model.cfg
    wrapper = tta.OneStageRCNNWithTTA(cfg, model)
    g = torch.Generator().manual_seed(5)
    inputs = [{"image": torch.randint(0, 256, (3, 160 + 32 * (k // 3), 192), dtype=torch.uint8, generator=g),
               "height": 320, "width": 384} for k in range(7)]
    a = wrapper._batch_inference(inputs)
    assert wrapper.use_cuda_graphs
    b = wrapper._batch_inference_deferred(inputs)  # captures one graph per batch shape
    c = wrapper._batch_inference_deferred(inputs)  # replays them
    # other image contents through the captured graphs, then the first ones again
    other = [{**d, "image": torch.randint(0, 256, tuple(d["image"].shape), dtype=torch.uint8, generator=g)} for d in inputs]
    o_eager = wrapper._batch_inference(other)
    o_graph = wrapper._batch_inference_deferred(other)
    assert len(a) == len(b) == len(c) == 7
    for x, y, z in zip(a + o_eager, b + o_graph, c + o_graph):
        for w in (y, z):
            assert torch.equal(x["instances"].pred_corners, w["instances"].pred_corners)
            assert torch.equal(x["instances"].scores, w["instances"].scores)
            assert torch.equal(x["instances"].pred_boxes.tensor, w["instances"].pred_boxes.tensor)
            assert x["instances"].image_size == w["instances"].image_size
    assert not torch.equal(a[0]["instances"].scores, o_eager[0]["instances"].scores)
    assert model.use_cuda_graphs is False  # the wrapper restores the model's own setting
