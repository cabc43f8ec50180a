"""CPU tests of the host logic: config loading, structures, spec, weight naming, C-ABI exports (no compute calls)."""
import glob
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_own_configs_resolve_to_specs():
    from dafne_b200.config import get_cfg
    from dafne_b200.spec import ModelSpec

    want = {"dota10_r50_1024.yaml": (50, 15, True, False), "dota10_r101_ms.yaml": (101, 15, True, True),
            "hrsc_r50_ms.yaml": (50, 1, True, False)}
    for f, (depth, c, sort_c, twc) in want.items():
        cfg = get_cfg()
        cfg.merge_from_file(os.path.join(ROOT, "configs", f))
        s = ModelSpec.from_cfg(cfg)
        assert (s.resnet_depth, s.num_classes, s.sort_corners, s.thresh_with_ctr) == (depth, c, sort_c, twc)
        assert s.score_thresh == 0.05 and s.pre_nms_topk == 2000 and s.nms_thresh == 0.1 and s.post_nms_topk == 1000
        assert s.pixel_mean == (123.675, 116.28, 103.53)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_reference_yaml_files_load_unchanged():
    from dafne_b200.config import get_cfg
    from dafne_b200.spec import ModelSpec

    table = {"dota-1.0_r101_ms": (101, 15, True, True), "dota-1.5_r101_ms": (101, 16, False, False),
             "hrsc_r101_ms": (101, 1, False, False), "hrsc_r50_ms": (50, 1, True, False),
             "ucas_aod_r101_ms": (101, 2, False, False)}
    for name, (depth, c, sort_c, twc) in table.items():
        cfg = get_cfg()
        cfg.merge_from_file(f"{REF}/configs/pre-trained/{name}.yaml")
        s = ModelSpec.from_cfg(cfg)
        assert (s.resnet_depth, s.num_classes, s.sort_corners, s.thresh_with_ctr) == (depth, c, sort_c, twc), name
    cfg = get_cfg()
    cfg.merge_from_file(f"{REF}/configs/dota-1.0/1024.yaml")  # _BASE_ chain + "(1024,)" literals
    assert cfg.INPUT.MIN_SIZE_TRAIN == (1024,) and cfg.SOLVER.IMS_PER_BATCH == 8
    assert ModelSpec.from_cfg(cfg).resnet_depth == 50
    for f in glob.glob(f"{REF}/configs/**/*.yaml", recursive=True):
        get_cfg().merge_from_file(f)  # every shipped YAML parses


def test_cfg_overrides_and_scope_guard():
    from dafne_b200.config import get_cfg
    from dafne_b200.spec import ModelSpec

    cfg = get_cfg()
    cfg.merge_from_list(["MODEL.DAFNE.NMS_TH", "0.2", "MODEL.RESNETS.DEPTH", 101, "MODEL.DAFNE.SORT_CORNERS", "False"])
    s = ModelSpec.from_cfg(cfg)
    assert s.nms_thresh == 0.2 and s.resnet_depth == 101 and s.sort_corners is False
    cfg.merge_from_list(["MODEL.DAFNE.USE_DEFORMABLE", True])
    with pytest.raises(NotImplementedError):
        ModelSpec.from_cfg(cfg)


def test_instances_and_boxes():
    from dafne_b200.structures import Boxes, Instances

    a = Instances((10, 20), scores=torch.tensor([0.9, 0.5, 0.1]), pred_boxes=Boxes(torch.zeros(3, 4)))
    assert len(a) == 3 and a.has("scores") and a.image_size == (10, 20)
    b = a[torch.tensor([True, False, True])]
    assert len(b) == 2 and b.scores.tolist() == pytest.approx([0.9, 0.1])
    c = Instances.cat([a, b])
    assert len(c) == 5 and isinstance(c.pred_boxes, Boxes)
    with pytest.raises(AssertionError):
        a.set("bad", torch.zeros(2))
    bx = Boxes(torch.tensor([[-5.0, 2.0, 30.0, 4.0], [1.0, 1.0, 1.0, 5.0]]))
    bx.clip((10, 20))
    assert bx.tensor[0].tolist() == [0.0, 2.0, 20.0, 4.0] and bx.nonempty().tolist() == [True, False]


def test_state_dict_names_follow_detectron2_tree():
    from dafne_b200.spec import ModelSpec
    from dafne_b200.weights import state_dict_shapes, synthetic_state_dict

    s50 = state_dict_shapes(ModelSpec(resnet_depth=50))
    s101 = state_dict_shapes(ModelSpec(resnet_depth=101))
    assert s50["backbone.bottom_up.stem.conv1.weight"] == (64, 3, 7, 7)
    assert s50["backbone.bottom_up.res3.0.shortcut.weight"] == (512, 256, 1, 1)
    assert s50["backbone.bottom_up.res5.2.conv2.norm.running_var"] == (512,)
    assert "backbone.bottom_up.res4.22.conv3.weight" in s101 and "backbone.bottom_up.res4.6.conv1.weight" not in s50
    assert s50["backbone.fpn_lateral5.weight"] == (256, 2048, 1, 1)
    assert s50["backbone.top_block.p7.bias"] == (256,)
    assert s50["proposal_generator.dafne_head.corners_tower.9.weight"] == (256, 256, 3, 3)
    assert s50["proposal_generator.dafne_head.cls_tower.10.bias"] == (256,)
    assert s50["proposal_generator.dafne_head.cls_logits.weight"] == (15, 256, 3, 3)
    assert s50["proposal_generator.dafne_head.scales.4.scale"] == (1,)
    n_params = sum(int(torch.tensor(v).prod()) for v in s101.values())
    assert 50e6 < n_params < 60e6  # ~54 M parameters
    sd = synthetic_state_dict(ModelSpec(resnet_depth=50), seed=0)
    sd2 = synthetic_state_dict(ModelSpec(resnet_depth=50), seed=0)
    assert list(sd) == list(s50) and all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_model_shell_has_reference_surface_without_gpu():
    from dafne_b200 import _capi
    from dafne_b200.config import get_cfg
    from dafne_b200.modeling import META_ARCH_REGISTRY, OneStageDetector

    cfg = get_cfg()
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg, init="zeros")
    assert isinstance(model, OneStageDetector) and not model.training
    keys = list(model.state_dict().keys())
    assert "backbone.bottom_up.res2.0.conv1.norm.running_mean" in keys
    assert hasattr(model.proposal_generator.dafne_outputs, "select_over_all_levels")
    sd = {k: torch.ones_like(v) for k, v in model.state_dict().items()}
    model.load_state_dict({"model": sd})  # detectron2 checkpoint wrapper
    assert float(model.state_dict()["backbone.fpn_output3.bias"].sum()) == 256.0
    with pytest.raises(NotImplementedError):
        model.train()
    if not torch.cuda.is_available():
        with pytest.raises(_capi.DafneError):  # no CPU fallback: the product path fails loudly without CUDA
            model([{"image": torch.zeros(3, 64, 64, dtype=torch.uint8)}])


def test_level_sizes():
    from dafne_b200.spec import ModelSpec

    assert ModelSpec().level_sizes(1024, 1024) == [(128, 128), (64, 64), (32, 32), (16, 16), (8, 8)]
    assert ModelSpec().level_sizes(800, 800) == [(100, 100), (50, 50), (25, 25), (13, 13), (7, 7)]


def test_library_loads_and_exports_every_declared_symbol():
    """The C-ABI library loads on a machine without a GPU and exports exactly what include/dafne_b200.h declares."""
    from dafne_b200 import _capi

    header = open(os.path.join(ROOT, "include", "dafne_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dafne_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    lib = _capi.lib()
    assert lib.dafne_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dafne_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    sass = subprocess.run(["cuobjdump", "-sass", _capi.LIB_PATH], capture_output=True, text=True).stdout
    if sass:  # tcgen05 / TMA really are in the binary (B200_PROFILING.md: UTCHMMA, UTMALDG, UTMASTG, LDTM)
        for mnemonic in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
            assert mnemonic in sass, mnemonic
