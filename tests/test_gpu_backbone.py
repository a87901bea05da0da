"""GPU parity tests of the hand-written ResNet-50-FPN backbone (pod_compare_b200/backbone_tc.py).
The oracle is the torch restatement of detectron2's backbone in pod_compare_b200/backbone.py evaluated in fp32 without
TF32 (and fp64 for the single-convolution checks); tolerance 1e-4 of each map's largest magnitude (north_star)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from pod_compare_b200 import backbone as BB
from pod_compare_b200 import backbone_tc as TC
from pod_compare_b200 import ops, synthetic as S


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    from pod_compare_b200 import _cabi
    _cabi.require_device()


def _rel(a, ref):
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("cin,cout,k,stride,hw,res,relu", [
    (64, 64, 1, 1, (24, 40), False, True),        # res2 conv1
    (64, 256, 1, 1, (24, 40), True, True),        # conv3 + residual + ReLU, 256-column block
    (256, 128, 1, 2, (24, 40), False, True),      # strided 1x1 (STRIDE_IN_1X1), 128-column block
    (128, 128, 3, 1, (23, 37), False, True),      # 3x3, ragged tiles
    (256, 512, 1, 2, (23, 37), False, False),     # strided shortcut, two column blocks, odd input size
    (512, 256, 3, 2, (23, 40), False, False),     # P6-style 3x3 / 2
    (1024, 256, 1, 1, (9, 13), False, False),     # lateral, deep K
])
def test_general_convolution_vs_fp64(cin, cout, k, stride, hw, res, relu):
    g = torch.Generator().manual_seed(cin + cout + k + stride)
    NB, (H, W) = 2, hw
    x = torch.randn((NB, cin, H, W), generator=g) * 2.0
    w = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn((cout,), generator=g) * 0.1
    cv = TC._Conv(w, None, b, "cuda")
    xs = ops.split_f32(x.permute(0, 2, 3, 1).contiguous().cuda(), scale=TC.ACT)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=pad)
    r = None
    if res:
        rr = torch.randn((NB, cout, Ho, Wo), generator=g)
        r = ops.split_f32(rr.permute(0, 2, 3, 1).contiguous().cuda(), scale=TC.ACT)
        ref = ref + rr.double()
    if relu:
        ref = torch.relu(ref)
    # split-pair output
    ohi = torch.zeros((NB, Ho, Wo, cv.rows), dtype=torch.float16, device="cuda")
    olo = torch.zeros_like(ohi)
    ops.conv_tc_general(xs[0], xs[1], TC.ACT, NB, H, W, cin, k, stride, cv.w_hi, cv.w_lo, cv.w_scale, cv.rows, cv.cout, cv.bias, relu,
                        cv.block, out_hi=ohi, out_lo=olo, out_scale=TC.ACT, out_ch_stride=cv.rows, res=r, res_scale=TC.ACT)
    torch.cuda.synchronize()
    assert ops.status() == 0
    got = ((ohi.float() + olo.float()) / TC.ACT).permute(0, 3, 1, 2)[:, :cout].cpu()
    assert _rel(got, ref) <= 1e-5, _rel(got, ref)
    if not res:
        # fp32 channels-last output
        of = torch.full((NB, Ho, Wo, cout), float("nan"), device="cuda")
        ops.conv_tc_general(xs[0], xs[1], TC.ACT, NB, H, W, cin, k, stride, cv.w_hi, cv.w_lo, cv.w_scale, cv.rows, cv.cout, cv.bias, relu,
                            cv.block, out_ch_stride=cout, out_f32=of)
        torch.cuda.synchronize()
        assert ops.status() == 0
        assert _rel(of.permute(0, 3, 1, 2).cpu(), ref) <= 1e-5


def test_stem_pool_upsample_and_split_kernels():
    g = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (2, 3, 50, 70), generator=g, dtype=torch.uint8)
    w = torch.randn((64, 3, 7, 7), generator=g) * 0.01
    b = torch.randn((64,), generator=g) * 0.1
    mean, std = (103.53, 116.28, 123.675), (1.0, 57.0, 58.0)
    H, W = 64, 96                                         # padded size
    x = (img.float() - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    x = F.pad(x, (0, W - 70, 0, H - 50))
    ref = F.max_pool2d(torch.relu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=3)), 3, 2, 1)
    Hc, Wc, Hp, Wp = 32, 48, 16, 24
    scratch = torch.empty((2, Hc, Wc, 64), device="cuda")
    ohi = torch.empty((2, Hp, Wp, 64), dtype=torch.float16, device="cuda"); olo = torch.empty_like(ohi)
    for im in (img.cuda(), img.float().cuda()):
        ops.stem_conv7_pool(im.contiguous(), H, W, mean, std, w.permute(2, 3, 1, 0).reshape(147, 64).contiguous().cuda(), b.cuda(), scratch,
                            ohi, olo, TC.ACT)
        got = ((ohi.float() + olo.float()) / TC.ACT).permute(0, 3, 1, 2).cpu()
        assert _rel(got, ref) <= 4e-6            # fp32 SIMT convolution + the 2^-19 split of its output
    d = torch.randn((2, 5, 7, 16), generator=g).cuda()
    s = torch.randn((2, 3, 4, 16), generator=g).cuda()
    want = d + F.interpolate(s.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")[:, :, :5, :7].permute(0, 2, 3, 1)
    ops.upsample2_add(d, s)
    assert torch.equal(d, want)
    v = torch.randn((3, 4, 5, 8), generator=g).cuda() * 100
    hi, lo = ops.split_f32(v, scale=2.0, relu=True)
    assert torch.allclose((hi.float() + lo.float()) / 2.0, torch.relu(v), rtol=2e-6, atol=1e-9)      # the pair holds x to 2^-19 (common.cuh POD_LO_BITS)


@pytest.mark.parametrize("hw", [(100, 190), (720, 1280)])
def test_backbone_matches_torch_fp32(hw):
    """Whole ResNet-50-FPN: 53 convolutions deep.  Against the torch fp32 backbone on the same weights."""
    sd = BB.random_state_dict(2)
    ref_net = BB.ResNetFPNBackbone(sd, device="cuda")
    net = TC.TcResNetFPNBackbone(sd, device="cuda")
    imgs = [S.make_image(0, i, hw[0], hw[1]) for i in range(2)]
    ref = ref_net(imgs)
    got = net(imgs)
    torch.cuda.synchronize()
    assert ops.status() == 0
    assert [tuple(f.shape) for f in got] == [tuple(f.shape) for f in ref]
    assert [tuple(f.shape[-2:]) for f in got] == S.level_shapes(hw[0], hw[1], divisibility=32)
    for l, (a, r) in enumerate(zip(got, ref)):
        assert _rel(a.cpu(), r.cpu()) <= 1e-4, (l, _rel(a.cpu(), r.cpu()))
    # float images == uint8 images; batch position independence
    solo = net([imgs[1].float()])
    assert all(torch.equal(s[0], g[1]) for s, g in zip(solo, got))


def test_predictor_from_raw_images_on_the_tc_backbone():
    """build_predictor(cfg)(input_im) from a raw image: hand-written backbone -> head -> detections; the channels-last
    FPN maps go into the head without a layout pass.  Equals infer_from_features on the same maps (bit for bit) and the
    torch-backbone result within the backbone tolerance."""
    from oracle import cases as C
    from pod_compare_b200.predictor import build_predictor
    name = "regclsvar_std"
    cfg = C.build_cfg(name)
    sd_head = S.make_head_state_dict(0, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True)
    sd_bb = BB.random_state_dict(1)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd_head)
    pred.load_backbone(sd_bb)                                   # impl="tc" is the default
    assert type(pred.backbone).__name__ == "TcResNetFPNBackbone"
    img = S.make_image(0, 0, 96, 160)
    inst = pred([{"image": img, "height": 96, "width": 160, "image_id": 0}])
    feats = pred.backbone([img])
    ref = pred.infer_from_features(feats, (96, 160), (96, 160), image0=0)[0]
    assert len(inst) == len(ref) and len(inst) > 0 and torch.equal(inst.pred_boxes.tensor, ref.pred_boxes.tensor)
    nchw = pred.infer_from_features([f.contiguous() for f in feats], (96, 160), (96, 160), image0=0)[0]
    assert torch.equal(nchw.scores, ref.scores) and torch.equal(nchw.pred_boxes.tensor, ref.pred_boxes.tensor)
    pred_t = build_predictor(cfg)
    pred_t.load_weight_sets(sd_head)
    pred_t.load_backbone(sd_bb, impl="torch")
    inst_t = pred_t([{"image": img, "height": 96, "width": 160, "image_id": 0}])
    assert abs(len(inst_t) - len(inst)) <= 2
    n = min(len(inst), len(inst_t), 20)
    assert torch.allclose(inst.scores[:n], inst_t.scores[:n], rtol=2e-3, atol=1e-5)


def test_checkpoints_found_like_the_reference_finds_them(tmp_path):
    """a16: `build_predictor(cfg)` loads <OUTPUT_DIR>/model_final.pth -- head AND backbone keys -- so that
    `build_predictor(cfg)(input_im)` works from a raw frame with no further calls (reference
    probabilistic_inference.py:72-84); 'ensembles' reads the sibling random_seed_<s> directories (:58-70), one full
    model (own backbone) per member."""
    from oracle import cases as C
    from pod_compare_b200.predictor import build_predictor
    img = S.make_image(0, 3, 64, 96)
    inp = [{"image": img, "height": 64, "width": 96, "image_id": 3}]
    # single model
    cfg = C.build_cfg("regclsvar_std")
    cfg.defrost()
    cfg.OUTPUT_DIR = str(tmp_path / "single" / "random_seed_0")
    cfg.freeze()
    sd = dict(S.make_head_state_dict(0, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True))
    sd.update(BB.random_state_dict(1))
    (tmp_path / "single" / "random_seed_0").mkdir(parents=True)
    torch.save({"model": sd}, str(tmp_path / "single" / "random_seed_0" / "model_final.pth"))
    pred = build_predictor(cfg)
    assert pred.backbone is not None and len(pred.weight_sets) == 1
    got = pred(inp)
    manual = build_predictor(C.build_cfg("regclsvar_std"))
    manual.load_weight_sets(sd)
    manual.load_backbone(sd)
    want = manual(inp)
    assert len(got) == len(want) and len(got) > 0 and torch.equal(got.pred_boxes.tensor, want.pred_boxes.tensor)
    # ensemble: three sibling directories, three different full models
    cfg = C.build_cfg("ensembles_e3")
    cfg.defrost()
    cfg.OUTPUT_DIR = str(tmp_path / "ens" / "random_seed_0")
    cfg.freeze()
    sds = []
    for e, seed in enumerate((0, 1000, 2000)):
        d = tmp_path / "ens" / ("random_seed_%d" % seed)
        d.mkdir(parents=True)
        m = dict(S.make_member_state_dicts(3, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True)[e])
        m.update(BB.random_state_dict(10 + e))
        sds.append(m)
        torch.save({"model": m}, str(d / "model_final.pth"))
    pred = build_predictor(cfg)
    assert len(pred.weight_sets) == 3 and pred.member_backbones is not None and len(pred.member_backbones) == 3
    got = pred(inp)
    feats = [net([img]) for net in pred.member_backbones]
    assert not torch.equal(feats[0][0], feats[1][0])                         # every member has its own feature maps
    want = pred.infer_from_features(feats, (64, 96), (64, 96), image0=3)[0]
    assert len(got) == len(want) and torch.equal(got.scores, want.scores)


def test_cuda_graph_replay_equals_eager_call():
    """predictor.capture(): the whole single-forward step (backbone + head + post-processing) as one CUDA graph; replays on
    new frames give bit-identical results to the eager call, and device-side errors still surface after a replay."""
    from oracle import cases as C
    from pod_compare_b200.predictor import build_predictor
    cfg = C.build_cfg("regclsvar_std")
    pred = build_predictor(cfg)
    pred.load_weight_sets(S.make_head_state_dict(0, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True))
    pred.load_backbone(BB.random_state_dict(1))
    frames = [torch.stack([S.make_image(0, 10 * k + i, 96, 160) for i in range(3)]) for k in range(3)]
    run = pred.capture(frames[0].cuda(), image0=5)
    assert run.launches > 50
    for k in (1, 2, 0):
        insts, det = run(frames[k].cuda() if k else frames[k].pin_memory())
        eager = pred.infer_from_images(frames[k].cuda(), image0=5)
        assert len(insts) == 3
        for a, b in zip(insts, eager):
            assert len(a) == len(b) and len(a) > 0
            assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor) and torch.equal(a.scores, b.scores)
            assert torch.equal(a.pred_boxes_covariance, b.pred_boxes_covariance)
    # features-in capture
    feats = [f.contiguous() for f in pred.backbone(frames[1].cuda())]
    run_f = pred.capture(feats, out_hw=(96, 160), image0=5)
    insts, _ = run_f(feats)
    eager = pred.infer_from_features(feats, (96, 160), (96, 160), image0=5)
    assert all(torch.equal(a.scores, b.scores) for a, b in zip(insts, eager))
