"""GPU parity tests, kernel by kernel, through the C ABI (pod_compare_b200.ops).

Tolerances (stated per SURVEY 8c / BASELINE north_star "within 1e-4 rel fp32, survivor indices
bit-exact"):
  * integer / index / mask outputs ........ bit-exact
  * Philox normals ........................ |d| <= 2e-6 (fp32 log/sincos a few ulp from the fp64-rounded oracle)
  * convolutions .......................... max|err| <= 1e-5 * max|ref| against an fp64 convolution
    (the fp32 CPU convolution the reference runs is itself ~1e-6 from fp64)
  * scores / boxes ........................ rtol 1e-4 (observed ~1e-6)
  * covariances ........................... max|err| <= 1e-4 * max|cov| per matrix
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import philox
from oracle import podref as O
from pod_compare_b200 import engine, ops, synthetic as S
from tests import gpu_util as G


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    from pod_compare_b200 import _cabi
    _cabi.require_device()      # raises (test error) rather than skipping: no silent fallback


# ------------------------------------------------------------------------------------------ RNG
def test_philox_dropout_mask_bit_exact():
    for (H, W, C, p) in ((6, 10, 256, 0.2), (8, 16, 64, 0.5), (3, 5, 256, 0.05)):
        ref = philox.dropout_keep_mask(77, 3, 5, 1, 1, 2, 4, H, W, C, p)
        got = ops.philox_dropout_mask(H, W, C, 77, 3, 5, 1, 1, 2, 4, p).cpu().numpy().astype(bool)
        assert np.array_equal(ref, got)


def test_philox_normals_close():
    ref = philox.logit_normals(5, 2, 3, 10, 333, 7)
    got = ops.philox_logit_normals(10, 333, 7, 5, 2, 3).cpu().numpy()
    assert np.abs(ref - got).max() <= 2e-6
    ids = np.array([0, 5, 17, 184139, 99999], dtype=np.int64)
    ref = philox.box_normals((1 << 40) + 9, 6, ids, 100)
    got = ops.philox_box_normals(torch.from_numpy(ids).cuda(), 100, (1 << 40) + 9, 6).cpu().numpy()
    assert np.abs(ref - got).max() <= 2e-6
    assert abs(float(got.mean())) < 0.2 and 0.8 < float(got.std()) < 1.2


# ------------------------------------------------------------------------------------------ prep
def test_layout_and_split_roundtrip():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 256, 5, 7), generator=g) * 3
    hi, lo = ops.nchw_to_nhwc_split(x.cuda(), 16.0)
    rec = ((hi.float() + lo.float()) / 16.0).permute(0, 3, 1, 2).cpu()
    assert G.rel_err(rec, x) < 1e-6
    f32 = ops.nchw_to_nhwc_f32(x.cuda()).permute(0, 3, 1, 2).cpu()
    assert torch.equal(f32, x)


def test_sample_mean_q1_bit_exact():
    g = torch.Generator().manual_seed(1)
    for n in (1001, 1004):          # scalar path / 16-byte vector path (n % 4 == 0)
        for S_ in (1, 2, 5, 9, 10, 18, 30):
            x = torch.randn((3, S_, n), generator=g)
            ref = O.quirk_mean([[x[:, s]] for s in range(S_)])[0] if S_ > 1 else x[:, 0]
            got = ops.sample_mean_q1(x.cuda()).cpu()
            assert torch.equal(got, ref), (n, S_)


def test_mask_expand_matches_oracle():
    g = torch.Generator().manual_seed(2)
    H, W, C, N = 4, 6, 256, 3
    x = torch.relu(torch.randn((2, H * W, C), generator=g))
    d = ops.make_dropout(0.2, 9, 4, N, 2, 0, 1, 0, 2)
    hi, lo = ops.mask_expand_split(x.cuda(), d, 16.0)
    rec = ((hi.float() + lo.float()) / 16.0).cpu().view(2, N, 2, H, W, C)
    drop = O.DropoutSource("philox", 0.2, 9, 0)
    for b in range(2):
        for s in range(N):
            for ps in range(2):
                drop.image = 4 + b
                ref = drop(x[b].view(1, H, W, C).permute(0, 3, 1, 2), 2, s, ps, 1, 0)[0].permute(1, 2, 0)
                assert G.rel_err(rec[b, s, ps], ref) < 1e-6


# ------------------------------------------------------------------------------------------ convolutions
CONV_SHAPES = [  # (NB, Cin, H, W, Cout)
    (1, 256, 8, 16, 256),      # exactly one tile (odd tile count: the pair kernel's tail path)
    (3, 256, 20, 40, 256),     # 27 tiles per launch, ragged in x and y, odd count
    (5, 64, 8, 32, 256),       # 10 tiles, Cin=64
    (2, 256, 6, 10, 63),       # P7-like, partial tile, N=64
    (3, 256, 12, 20, 36),      # P6-like, N=48
    (1, 256, 24, 40, 256),     # P5-like, ragged tiles in x
    (2, 128, 9, 17, 90),       # Cin=128, N=96, one pixel past a tile in both directions
    (1, 64, 17, 33, 72),       # Cin=64, N=80
    (1, 256, 16, 32, 126),     # N=128
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_simt_conv_vs_fp64(shape):
    NB, Cin, H, W, Cout = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn((NB, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn((Cout,), generator=g) * 0.1
    cpad = (Cout + 63) // 64 * 64
    wk = ops.pack_conv_weight_f32(w.cuda(), cpad)
    bias = torch.zeros(cpad, device="cuda")
    bias[:Cout] = b.cuda()
    out = ops.conv3x3_simt(ops.nchw_to_nhwc_f32(x.cuda()), wk, bias, Cout, cpad, True)
    got = out.view(NB, H, W, Cout).permute(0, 3, 1, 2).cpu()
    assert G.rel_err(got, G.conv_ref64(x, w, b, True)) < 1e-5


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("staging", ["halo64", "tap64", "tap32"])
@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_tc_conv_raw_vs_fp64(shape, staging, pair):
    """Every operand-staging variant of the tcgen05 kernel: row-halo boxes (one 10-row box serves three taps),
    one box per tap with K-block 64 (default), one box per tap with K-block 32."""
    NB, Cin, H, W, Cout = shape
    if pair == 0 and Cout != 256:
        pytest.skip("single-CTA / paired choice only exists for 256 output channels")
    kblock = 32 if staging == "tap32" else 64
    ops.set_conv_kblock(kblock)
    ops.set_conv_halo(1 if staging == "halo64" else 0)
    ops.set_conv_pair(pair)
    try:
        g = torch.Generator().manual_seed(4)
        x = torch.randn((NB, Cin, H, W), generator=g) * 2.0
        w = torch.randn((Cout, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5
        b = torch.randn((Cout,), generator=g) * 0.1
        got = G.tc_conv_raw(x, w, b, False)
        ref = G.conv_ref64(x, w, b, False)
        err = G.rel_err(got, ref)
        cpu32 = G.rel_err(torch.nn.functional.conv2d(x, w, b, padding=1), ref)
        print("tc conv %s kblock=%d: rel err vs fp64 %.3e (fp32 CPU conv: %.3e)" % (shape, kblock, err, cpu32))
        assert not torch.isnan(got).any()
        assert err < 1e-5
    finally:
        ops.set_conv_kblock(64)
        ops.set_conv_halo(DEFAULT_HALO)
        ops.set_conv_pair(1)


DEFAULT_HALO = 3     # library default of pod_conv3x3_tc_set_halo (bit 1 = weights-as-A kernel)

WT_SHAPES = [  # (NB, Cin, H, W, Cout): output convolutions of <= 64 channels run weights-as-A (16x16 pixel tiles)
    (2, 256, 6, 10, 63),       # P7-like: one partial tile
    (3, 256, 20, 40, 63),      # ragged in x and y, several maps
    (1, 64, 17, 33, 40),       # one pixel past a tile in both directions, Cin=64
    (2, 128, 9, 17, 64),       # all 64 rows used
    (1, 256, 48, 80, 63),      # P4-like: 15 full tiles
    (5, 256, 16, 16, 7),       # exactly one tile per map, few channels
]


@pytest.mark.parametrize("halo", [2, 0])
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("shape", WT_SHAPES)
def test_tc_conv_weights_as_a_vs_fp64(shape, relu, halo):
    """k_conv3x3_wt (stacked [w_hi; w_lo] as the A operand, 256 pixels as N) against an fp64 convolution and
    against the pixels-as-M kernel on the same inputs."""
    NB, Cin, H, W, Cout = shape
    g = torch.Generator().manual_seed(40 + Cout)
    x = torch.randn((NB, Cin, H, W), generator=g) * 2.0
    w = torch.randn((Cout, Cin, 3, 3), generator=g) * (2.0 / (9 * Cin)) ** 0.5
    b = torch.randn((Cout,), generator=g) * 0.1
    ref = G.conv_ref64(x, w, b, relu)
    try:
        ops.set_conv_wt(1)
        ops.set_conv_halo(halo)                 # 2: 18-row pixel boxes shared by three taps; 0: one box per tap
        got = G.tc_conv_raw(x, w, b, relu, cout_pad=64)
        ops.set_conv_wt(0)
        ops.set_conv_halo(0)
        other = G.tc_conv_raw(x, w, b, relu, cout_pad=64)
    finally:
        ops.set_conv_wt(1)
        ops.set_conv_halo(DEFAULT_HALO)
    assert not torch.isnan(got).any()          # the output buffer is NaN-filled: every valid element was written
    err, err_other = G.rel_err(got, ref), G.rel_err(other, ref)
    print("weights-as-A %s relu=%s: rel err vs fp64 %.3e (pixels-as-M: %.3e)" % (shape, relu, err, err_other))
    assert err < 1e-5 and err_other < 1e-5
    assert G.rel_err(got, other.double()) < 4e-6


def test_tc_conv_chunked_accumulation_is_more_accurate():
    """tcgen05 accumulates with truncation; summing one tap per TMEM chain and adding the nine partial
    sums in fp32 RN (the default) must beat the single 2304-long chain and stay within 3e-6."""
    g = torch.Generator().manual_seed(8)
    x = torch.relu(torch.randn((1, 256, 24, 40), generator=g)) * 1.25
    w = torch.randn((256, 256, 3, 3), generator=g) * (2.0 / 2304) ** 0.5
    b = torch.randn((256,), generator=g) * 0.05
    ref = G.conv_ref64(x, w, b, False)
    err = {}
    try:
        ops.set_conv_chunk_kblocks(0)
        for taps in (9, 1):
            ops.set_conv_chunk_taps(taps)
            err[taps] = G.rel_err(G.tc_conv_raw(x, w, b, False), ref)
    finally:
        ops.set_conv_chunk_taps(1)
        ops.set_conv_chunk_kblocks(12)
    print("conv rel err vs fp64: single chain %.3e, per-tap chunks %.3e" % (err[9], err[1]))
    assert err[1] < 3e-6
    assert err[1] < err[9]


def test_truncation_compensation_removes_the_accumulation_bias():
    """tcgen05 adds every MMA into the fp32 accumulator with truncation toward zero, so a TMEM chain of n
    accumulations comes out ~0.27*n*2^-24 too small (measured: profiles/r1e_trunc_comp_robustness.txt).  The
    default epilogue scale compensates the expected loss: the signed bias against an fp64 convolution must vanish
    and the rms error must drop, for signed and for ReLU inputs."""
    g = torch.Generator().manual_seed(21)
    w = torch.randn((256, 256, 3, 3), generator=g) * (2.0 / 2304) ** 0.5
    b = torch.zeros(256)
    for x in (torch.relu(torch.randn((1, 256, 24, 40), generator=g)) * 1.25, torch.randn((1, 256, 24, 40), generator=g)):
        ref = G.conv_ref64(x, w, b, False)
        out = {}
        try:
            for comp in (0.0, 0.27):
                ops.set_conv_trunc_comp(comp)
                d = G.tc_conv_raw(x, w, b, False).double() - ref
                out[comp] = (float((d * ref.sign()).mean() / ref.abs().mean()), float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))
        finally:
            ops.set_conv_trunc_comp(0.27)
        print("bias / rms vs fp64: uncompensated %+.2e / %.2e, compensated %+.2e / %.2e" % (out[0.0] + out[0.27]))
        assert out[0.0][0] < -1.5e-6                  # the raw chain (12 K-blocks = 144 accumulations) is biased low
        assert abs(out[0.27][0]) < 3e-7               # compensated: no bias left
        assert out[0.27][1] < 0.7 * out[0.0][1] and out[0.27][1] < 2e-6


@pytest.mark.parametrize("halo", [1, 0])
def test_tc_conv_hidden_dropout_and_strided_input(halo):
    ops.set_conv_halo(halo)
    try:
        _hidden_dropout_and_strided_input()
    finally:
        ops.set_conv_halo(DEFAULT_HALO)


def _hidden_dropout_and_strided_input():
    g = torch.Generator().manual_seed(5)
    NB, H, W = 4, 10, 18
    x = torch.randn((NB, 256, H, W), generator=g)
    w = torch.randn((256, 256, 3, 3), generator=g) * (2.0 / 2304) ** 0.5
    b = torch.randn((256,), generator=g) * 0.05
    # maps decode as image = 7 + n // (2*1), sample = n % 2, pass = 1
    d = ops.make_dropout(0.2, 123, 7, 2, 1, 1, 0, 3, 1)
    got = G.tc_conv_hidden(x, w, b, d)
    ref = G.conv_ref64(x, w, b, True)
    for n in range(NB):
        keep = philox.dropout_keep_mask(123, 7 + n // 2, n % 2, 1, 0, 3, 1, H, W, 256, 0.2)
        keep = torch.from_numpy(np.ascontiguousarray(keep.transpose(2, 0, 1)))
        r = ref[n] * keep * 1.25
        assert G.rel_err(got[n], r) < 1e-5, n
        # dropped elements are exactly zero
        assert float(got[n][~keep].abs().max()) == 0.0
    # every 2nd map through in_map_stride / in_offset (how the variance heads read pass-1 maps)
    hi, lo = ops.nchw_to_nhwc_split(x.cuda(), 16.0)
    pcv = engine.pack_conv(w[:63], b[:63], "cuda")
    out = torch.zeros((2, H * W, 63), device="cuda")
    ops.conv3x3_tc(hi, lo, 16.0, 2, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, 63, 64, G.POD_OUT_RAW, False,
                   out_f32=out, out_map_stride=H * W * 63, out_pixel_stride=63, in_map_stride=2 * H * W * 256,
                   in_offset=H * W * 256)
    torch.cuda.synchronize()
    got2 = out.view(2, H, W, 63).permute(0, 3, 1, 2).cpu()
    ref2 = G.conv_ref64(x[1::2], w[:63], b[:63], False)
    assert G.rel_err(got2, ref2) < 1e-5


TILE_WIDTH_SHAPES = [  # (NB, H, W): 256 -> 256 channels on the CTA-pair row-halo kernel
    (1, 4, 32),        # exactly one 32 x 4 tile (odd tile count: the pair's idle CTA)
    (2, 17, 9),        # narrower than either tile, one row past two 16 x 8 tiles
    (2, 92, 160),      # P3 of a 1280 x 720 frame: 115 tiles of 32 x 4 against 120 of 16 x 8 (the shape the choice exists for)
    (3, 13, 21),       # ragged in both directions for both geometries
    (2, 5, 37),        # one row / five columns past a tile
    (5, 1, 1),         # single pixel
    (1, 23, 40),       # P5
]


@pytest.mark.parametrize("shape", TILE_WIDTH_SHAPES)
def test_tc_conv_tile_widths_are_bit_identical(shape):
    """pod_conv3x3_tc_set_tile_width: the CTA-pair row-halo kernel covers a map with 16 x 8 or 32 x 4 pixel tiles.  Every
    output pixel sees the same K order of the same MMAs either way, so raw outputs, hidden activations (hi / lo pairs)
    and the in-epilogue dropout must be bit-identical; the default (0) picks per map shape."""
    NB, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn((NB, 256, H, W), generator=g) * 1.5
    w = torch.randn((256, 256, 3, 3), generator=g) * (2.0 / 2304) ** 0.5
    b = torch.randn((256,), generator=g) * 0.05
    d = ops.make_dropout(0.1, 77, 3, 1, 1, 2, 1, 2, 0)
    out = {}
    try:
        for tw in (16, 32, 0):
            ops.set_conv_tile_width(tw)
            raw = G.tc_conv_raw(x, w, b, False)
            hid = G.tc_conv_hidden(x, w, b, d)
            assert ops.status() == 0
            out[tw] = (raw, hid)
    finally:
        ops.set_conv_tile_width(0)
    assert G.rel_err(out[16][0], G.conv_ref64(x, w, b, False)) < 1e-5
    for tw in (32, 0):
        assert torch.equal(out[tw][0], out[16][0]), tw
        assert torch.equal(out[tw][1], out[16][1]), tw


# ------------------------------------------------------------------------------------------ scores / top-k
def _level_off(level_hw, A=9):
    off = [0]
    for h, w in level_hw:
        off.append(off[-1] + h * w * A)
    return off


def test_scores_and_topk_vs_oracle():
    g = torch.Generator().manual_seed(6)
    level_hw = [(12, 20), (6, 10), (3, 5), (2, 3), (1, 2)]
    off = _level_off(level_hw)
    R, K, B = off[-1], 7, 2
    logits = torch.randn((B, R, K), generator=g) * 1.2 - 3.0
    logvar = torch.randn((B, R, K), generator=g) * 0.5 - 2.0
    for use_var in (True, False):
        probs, score, cls = ops.scores(logits.cuda(), logvar.cuda() if use_var else None, off, 10, 31, 5)
        cand_idx, cand_cnt, seg = ops.topk_levels(score, off, 100, 0.05)
        probs, score, cls, cand_idx, cand_cnt = [t.cpu() for t in (probs, score, cls, cand_idx, cand_cnt)]
        for b in range(B):
            for l in range(len(level_hw)):
                mu = logits[b, off[l]:off[l + 1]]
                if use_var:
                    eps = torch.from_numpy(philox.logit_normals(31, 5 + b, l, 10, mu.shape[0], K))
                    ref = torch.mean((mu + eps * torch.sqrt(torch.exp(logvar[b, off[l]:off[l + 1]]))).sigmoid_(), 0)
                else:
                    ref = mu.clone().sigmoid_()
                got = probs[b, off[l]:off[l + 1]]
                assert torch.allclose(got, ref, rtol=1e-4, atol=1e-7)
                # selection is checked on the GPU's own scores (bit-exact rule: top-k, stable, > thresh)
                sc = score[b, off[l]:off[l + 1]]
                assert torch.equal(sc, got.max(1)[0])
                assert torch.equal(cls[b, off[l]:off[l + 1]].long(), got.max(1)[1])
                k = min(100, sc.shape[0])
                order = torch.sort(sc, descending=True, stable=True)[1][:k]
                order = order[sc[order] > 0.05]
                n = int(cand_cnt[b, l])
                assert n == order.numel()
                assert torch.equal(cand_idx[b, seg[l]:seg[l] + n].long(), order + off[l])


def test_topk_ties_and_full_range():
    # many equal scores: ties resolve to the lower anchor index; k larger than the level
    off = [0, 5000, 5040]
    score = torch.full((1, 5040), 0.5)
    score[0, 100:200] = 0.75
    score[0, 4000:4100] = 0.25
    score[0, 5000:5040] = torch.linspace(0.01, 0.4, 40)
    cand_idx, cand_cnt, seg = ops.topk_levels(score.cuda(), off, 1000, 0.05)
    cand_idx, cand_cnt = cand_idx.cpu(), cand_cnt.cpu()
    for l in range(2):
        sc = score[0, off[l]:off[l + 1]]
        k = min(1000, sc.shape[0])
        order = torch.sort(sc, descending=True, stable=True)[1][:k]
        order = order[sc[order] > 0.05]
        assert int(cand_cnt[0, l]) == order.numel()
        assert torch.equal(cand_idx[0, seg[l]:seg[l] + order.numel()].long(), order + off[l])


# ------------------------------------------------------------------------------------------ randomized selection tests
def _random_candidates(seed, M, K=7, tie_scores=False, img=(720, 1280)):
    g = torch.Generator().manual_seed(seed)
    n_gt = max(1, M // 12)
    centers = torch.rand((n_gt, 2), generator=g) * torch.tensor([img[1] * 0.8, img[0] * 0.8])
    sizes = torch.rand((n_gt, 2), generator=g) * 250 + 20
    which = torch.randint(0, n_gt, (M,), generator=g)
    jit = torch.randn((M, 4), generator=g) * torch.rand((M, 1), generator=g) * 12
    x1y1 = centers[which] + jit[:, :2]
    boxes = torch.cat([x1y1, x1y1 + sizes[which] + jit[:, 2:]], 1).float()
    boxes[:: max(M // 7, 1), 2:] = boxes[:: max(M // 7, 1), :2]              # some zero-area boxes
    classes = torch.randint(0, K, (M,), generator=g)
    scores = torch.rand((M,), generator=g) * 0.9 + 0.05
    if tie_scores:
        scores = (scores * 20).round() / 20 + 0.01                           # many exactly equal scores
    probs = torch.rand((M, K), generator=g) * 0.04
    probs[torch.arange(M), classes] = scores
    A = torch.randn((M, 4, 4), generator=g)
    cov = A @ A.transpose(1, 2) + torch.eye(4)[None]
    return O.Candidates(boxes, cov.float(), scores.float(), classes, probs.float(), np.arange(M), [M])


@pytest.mark.parametrize("M", [1, 2, 33, 257, 999, 1000, 1001, 2500, 5000])
@pytest.mark.parametrize("ties", [False, True])
def test_nms_survivors_bit_exact_random(M, ties):
    """NMS survivor indices against the scalar restatement of torchvision's CPU loop (stable order, lower
    index first on equal scores) and against the installed torchvision op when scores are distinct; both
    batched_nms variants; zero-area boxes included."""
    c = _random_candidates(1000 + M, M, tie_scores=ties)
    pp = O.PathParams()
    cd = G.cand_to_dict(c)
    for variant in (ops.NMS_AUTO, ops.NMS_VANILLA, ops.NMS_TRICK):
        det = ops.nms_fuse(cd, 0, 0.5, 0.9, 100, (720, 1280), (720, 1280), nms_variant=variant)
        n = int(det["keep_count"][0])
        got = det["keep"][0, :n].cpu().numpy().astype(np.int64)
        auto_is_vanilla = 4 * M > 4000
        if variant == ops.NMS_AUTO or (variant == ops.NMS_VANILLA) == auto_is_vanilla:
            ref = O.standard_nms_post(c, pp, (720, 1280), nms_impl="loop").keep.numpy()
            assert np.array_equal(got, ref), (M, ties, variant)
            if not ties:
                ref_tv = O.standard_nms_post(c, pp, (720, 1280), nms_impl="torchvision").keep.numpy()
                assert np.array_equal(got, ref_tv), (M, variant)
        else:
            # the other torchvision variant on the same boxes: restate it directly
            b = c.boxes.float()
            if variant == ops.NMS_TRICK:
                off = c.classes.to(b) * (b.max() + torch.tensor(1).to(b))
                keep = O.nms_loop(b + off[:, None], c.scores, 0.5)[:100].numpy()
            else:
                mask = torch.zeros_like(c.scores, dtype=torch.bool)
                for k in torch.unique(c.classes):
                    cur = torch.where(c.classes == k)[0]
                    mask[cur[O.nms_loop(b[cur], c.scores[cur], 0.5)]] = True
                kept = torch.where(mask)[0]
                keep = kept[torch.sort(c.scores[kept], descending=True, stable=True)[1]][:100].numpy()
            assert np.array_equal(got, keep), (M, ties, variant)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_topk_random_levels_with_ties(seed):
    g = torch.Generator().manual_seed(seed)
    sizes = [int(x) for x in torch.randint(1, 6000, (5,), generator=g)]
    off = [0]
    for s_ in sizes:
        off.append(off[-1] + s_)
    B = 3
    score = torch.rand((B, off[-1]), generator=g)
    score = torch.where(torch.rand(score.shape, generator=g) < 0.5, (score * 50).round() / 50, score)   # heavy ties
    for topk, thr in ((1000, 0.05), (7, 0.5), (1024, 0.0)):
        cand_idx, cand_cnt, seg = ops.topk_levels(score.cuda(), off, topk, thr)
        cand_idx, cand_cnt = cand_idx.cpu(), cand_cnt.cpu()
        for b in range(B):
            for l in range(5):
                sc = score[b, off[l]:off[l + 1]]
                order = torch.sort(sc, descending=True, stable=True)[1][: min(topk, sc.shape[0])]
                order = order[sc[order] > thr]
                assert int(cand_cnt[b, l]) == order.numel(), (seed, topk, b, l)
                assert torch.equal(cand_idx[b, seg[l]:seg[l] + order.numel()].long(), order + off[l])


def test_tc_conv_dual_destination_raw():
    """One launch, two heads: channels [0,63) -> first buffer, [63,126) -> second (eval-mode cls_score|cls_var)."""
    g = torch.Generator().manual_seed(21)
    NB, H, W = 2, 9, 13
    x = torch.randn((NB, 256, H, W), generator=g)
    w = torch.randn((126, 256, 3, 3), generator=g) * (2.0 / 2304) ** 0.5
    b = torch.randn((126,), generator=g) * 0.1
    hi, lo = ops.nchw_to_nhwc_split(x.cuda(), 16.0)
    pcv = engine.pack_conv(w, b, "cuda")
    o1 = torch.full((NB, H * W, 63), float("nan"), device="cuda")
    o2 = torch.full((NB, H * W, 63), float("nan"), device="cuda")
    ops.conv3x3_tc(hi, lo, 16.0, NB, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, 126, pcv.cout_pad, G.POD_OUT_RAW,
                   False, out_f32=o1, out_map_stride=H * W * 63, out_pixel_stride=63, out2_f32=o2, split_col=63,
                   out2_map_stride=H * W * 63, out2_pixel_stride=63)
    torch.cuda.synchronize()
    ref = G.conv_ref64(x, w, b, False)
    got1 = o1.view(NB, H, W, 63).permute(0, 3, 1, 2).cpu()
    got2 = o2.view(NB, H, W, 63).permute(0, 3, 1, 2).cpu()
    assert G.rel_err(got1, ref[:, :63]) < 1e-5 and G.rel_err(got2, ref[:, 63:]) < 1e-5
