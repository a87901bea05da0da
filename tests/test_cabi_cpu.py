"""CPU-side checks of the boundary: the library builds, loads, exports every symbol that
include/podb200.h declares, the ctypes mirrors match the C struct layouts, argument validation
returns error codes (no compute without a GPU), and the product refuses to run without a device."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "podb200.h")


@pytest.fixture(scope="module")
def lib():
    from pod_compare_b200 import build, _cabi
    build.build()
    return _cabi.load_library()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pod_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from pod_compare_b200 import _cabi
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.EXPORTS) == names


def test_struct_layouts_match_header(tmp_path):
    from pod_compare_b200 import _cabi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(pod_dropout),sizeof(pod_conv_args),sizeof(pod_decode_args),sizeof(pod_nms_args),'
                   'offsetof(pod_conv_args,out2_pixel_stride),offsetof(pod_decode_args,out_anchor),offsetof(pod_nms_args,skip_post),'
                   'sizeof(pod_merge_args),offsetof(pod_merge_args,out_count),offsetof(pod_conv_args,map_live));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_cabi.Dropout), ctypes.sizeof(_cabi.ConvArgs), ctypes.sizeof(_cabi.DecodeArgs),
            ctypes.sizeof(_cabi.NmsArgs), _cabi.ConvArgs.out2_pixel_stride.offset, _cabi.DecodeArgs.out_anchor.offset,
            _cabi.NmsArgs.skip_post.offset, ctypes.sizeof(_cabi.MergeArgs), _cabi.MergeArgs.out_count.offset,
            _cabi.ConvArgs.map_live.offset]
    assert got == want


def test_error_codes_and_messages(lib):
    assert lib.pod_version() == 2
    rc = lib.pod_conv3x3_tc(None, None)
    assert rc < 0 and b"null args" in lib.pod_last_error()
    rc = lib.pod_sample_mean_q1(None, 0, 0, 0, None, None)
    assert rc < 0 and b"pod_sample_mean_q1" in lib.pod_last_error()
    rc = lib.pod_conv3x3_tc_set_kblock(48)
    assert rc < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu():
    from pod_compare_b200 import _cabi
    from pod_compare_b200.predictor import build_predictor
    from oracle import cases as C
    with pytest.raises(_cabi.PodError):
        build_predictor(C.build_cfg("baseline_std"))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pod_compare_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_reference_config_yaml_surface():
    """The reference's own YAML files (model + inference config, two-stage merge) load unchanged
    when present; the key set the path reads exists with the reference defaults."""
    from pod_compare_b200.config import get_cfg, setup_config
    cfg = get_cfg()
    assert cfg.PROBABILISTIC_INFERENCE.AFFINITY_THRESHOLD == 0.7
    assert cfg.PROBABILISTIC_INFERENCE.ENSEMBLES.RANDOM_SEED_NUMS == [0, 1000, 2000, 3000, 4000]
    assert cfg.MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.NUM_SAMPLES == 1000
    d = "/root/reference/src/configs"
    if os.path.isdir(d):
        c = setup_config(os.path.join(d, "BDD-Detection/retinanet/retinanet_R_50_FPN_1x_reg_cls_var_dropout.yaml"),
                         os.path.join(d, "Inference/bayes_od_mc_dropout.yaml"))
        assert c.MODEL.RETINANET.NUM_CLASSES == 7 and c.MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE == 0.2
        assert c.PROBABILISTIC_INFERENCE.INFERENCE_MODE == "bayes_od" and c.PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS == 10
        assert c.PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE == "max_score"
        assert abs(c.MODEL.ANCHOR_GENERATOR.SIZES[0][1] - 32 * 2 ** (1 / 3)) < 1e-9
        assert c.is_frozen()


def test_build_predictor_rejects_other_meta_architectures_like_the_reference():
    """probabilistic_inference.py:20-33: build_predictor raises ValueError('Invalid meta-architecture ...') for anything
    but ProbabilisticRetinaNet -- before any device is touched, so the check also holds on a CPU-only host."""
    from pod_compare_b200.config import get_cfg
    from pod_compare_b200.predictor import build_predictor
    cfg = get_cfg()
    cfg.MODEL.META_ARCHITECTURE = "GeneralizedRCNN"
    with pytest.raises(ValueError, match="Invalid meta-architecture GeneralizedRCNN"):
        build_predictor(cfg)


def test_yaml_eval_tag_is_arithmetic_only(tmp_path):
    """The reference's YAML carries `!!python/object/apply:eval` for the anchor sizes (Base-RetinaNet: a nested list
    comprehension).  The loader evaluates that construct by walking a whitelisted AST -- arithmetic on numbers and
    comprehension variables only -- so a config file shipped next to a checkpoint cannot run code: the classic
    escapes from eval() with emptied builtins must be rejected, not executed."""
    import yaml
    from pod_compare_b200.config import _safe_arith_eval, load_yaml_with_base
    got = _safe_arith_eval("[[x, x * 2**(1.0/3), x * 2**(2.0/3)] for x in [32, 64, 128, 256, 512]]")
    assert len(got) == 5 and got[0][0] == 32 and abs(got[4][2] - 512 * 2 ** (2.0 / 3)) < 1e-9
    assert _safe_arith_eval("[-(1 + 2) * 3, 7 // 2, 2 ** 10]") == [-9, 3, 1024]
    for bad in ("().__class__.__base__.__subclasses__()",          # attribute walk to every loaded class
                "__import__('os').system('true')",                  # call
                "[x for x in range(3)]",                            # call as the iterable
                "open('/etc/passwd')",
                "(lambda: 1)()",
                "[x for x in [1, 2] if x]",                         # filters are not part of the construct
                "y + 1",                                            # free name
                "[1, 2][0]",                                        # subscript
                "'a' * 3",                                          # non-numeric constant
                "2 ** 4096"):                                       # resource exhaustion
        with pytest.raises((yaml.YAMLError, SyntaxError)):
            _safe_arith_eval(bad)
    # through the loader, as the tag appears in the reference's files
    f = tmp_path / "c.yaml"
    f.write_text("MODEL:\n  ANCHOR_GENERATOR:\n    SIZES: !!python/object/apply:eval [\"[[x, 2 * x] for x in [8, 16]]\"]\n")
    assert load_yaml_with_base(str(f))["MODEL"]["ANCHOR_GENERATOR"]["SIZES"] == [[8, 16], [16, 32]]
    f.write_text("X: !!python/object/apply:eval [\"__import__('os').getcwd()\"]\n")
    with pytest.raises(yaml.YAMLError):
        load_yaml_with_base(str(f))
    f.write_text("X: !!python/object/apply:os.getcwd []\n")          # any other python tag stays unknown to the loader
    with pytest.raises(yaml.YAMLError):
        load_yaml_with_base(str(f))


def test_backbone_shapes_and_keys_cpu():
    """The upstream ResNet-50-FPN restatement (library torch ops): detectron2 key coverage and the
    P3..P7 geometry (detectron2 pads to the res5 stride 32: 100x190 -> 128x192; P6 / P7 are stride-2 3x3 convolutions)."""
    from pod_compare_b200 import backbone as BB
    sd = BB.random_state_dict(0)
    assert sorted(sd) == sorted(BB.expected_keys())
    net = BB.ResNetFPNBackbone(sd, device="cpu")
    img = torch.randint(0, 256, (3, 100, 190), dtype=torch.uint8)
    feats = net([img, img])
    assert [tuple(f.shape) for f in feats] == [(2, 256, 16, 24), (2, 256, 8, 12), (2, 256, 4, 6), (2, 256, 2, 3), (2, 256, 1, 2)]
    from pod_compare_b200 import synthetic as S
    assert [tuple(f.shape[-2:]) for f in feats] == S.level_shapes(100, 190, divisibility=32)
    # a 1280x720 frame: 736 rows (not 768), as detectron2's RetinaNet backbone produces
    assert S.level_shapes(720, 1280, divisibility=32) == [(92, 160), (46, 80), (23, 40), (12, 20), (6, 10)]
    big = BB.ResNetFPNBackbone(sd, device="cpu", size_divisibility=128)
    assert [tuple(f.shape[-2:]) for f in big([img])] == [(16, 32), (8, 16), (4, 8), (2, 4), (1, 2)]
    assert all(torch.isfinite(f).all() for f in feats)
    assert torch.equal(feats[0][0], feats[0][1])


def test_shipped_configs_cover_the_reference_plugin_surface():
    """The YAMLs shipped in pod_compare_b200/configs (5 model variants x 8 inference modes) resolve, and --
    where the reference tree is present -- to exactly the configuration the reference's own files give."""
    import glob
    from pod_compare_b200.config import setup_config
    d = os.path.join(ROOT, "pod_compare_b200", "configs")
    models = sorted(glob.glob(os.path.join(d, "BDD-Detection", "retinanet", "retinanet*.yaml")))
    modes = sorted(glob.glob(os.path.join(d, "Inference", "*.yaml")))
    assert len(models) == 4 and len(modes) == 8
    ref = "/root/reference/src/configs"
    for m in models:
        for i in modes:
            cfg = setup_config(m, i)
            assert cfg.MODEL.META_ARCHITECTURE == "ProbabilisticRetinaNet" and cfg.MODEL.RETINANET.NUM_CLASSES == 7
            if os.path.isdir(ref):
                assert cfg == setup_config(m.replace(d, ref), i.replace(d, ref)), (m, i)
