"""Tensor-level wrappers of the C ABI (one function per `pod_*` entry point).

PyTorch is only the owner of device memory and streams here: each wrapper checks device / dtype /
contiguity, allocates outputs with torch, and launches the hand-written kernel on the current
stream.  Reference lines each op replaces are cited in include/podb200.h.
"""
import ctypes as C
import math

import torch

from . import _cabi
from ._cabi import (POD_OUT_HIDDEN, POD_OUT_RAW, ConvArgs, ConvGArgs, DecodeArgs, Dropout, MergeArgs, NmsArgs, WireArgs, check, int_array,
                    ptr, stream_ptr)

_launches = 0
PROFILE = None      # bench.py sets this to a list: (start_event, end_event, algorithmic FLOPs, tag) per conv launch


def launch_count():
    """Number of libpodb200 kernel launches issued by this process (bench.py's gpu_launches)."""
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _chk(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _cabi.PodError("%s must be a contiguous CUDA tensor of dtype %s" % (name, dtype))
    return t


def make_dropout(p=0.0, seed=0, image0=0, samples=1, passes=1, pass0=0, tower=0, layer=0, level=0):
    return Dropout(float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, image0, samples, passes, pass0, tower, layer, level)


def pow2_scale(max_abs, target=1024.0):
    """Largest power of two s with max_abs * s <= target (exact rescaling of fp32 values)."""
    if not (max_abs > 0.0) or not math.isfinite(max_abs):
        return 1.0
    return float(2.0 ** math.floor(math.log2(target / max_abs)))


# ---------------------------------------------------------------------------------------------------
def philox_dropout_mask(H, W, Cn, seed, image, sample, pass_, tower, layer, level, p):
    lib = _cabi.require_device()
    out = torch.empty((H, W, Cn), dtype=torch.uint8, device="cuda")
    check(lib.pod_philox_dropout_mask(ptr(out), H, W, Cn, seed, image, sample, pass_, tower, layer, level, float(p),
                                      stream_ptr()), "pod_philox_dropout_mask")
    _count()
    return out


def philox_logit_normals(draws, n_anchor, K, seed, image, level):
    lib = _cabi.require_device()
    out = torch.empty((draws, n_anchor, K), dtype=torch.float32, device="cuda")
    check(lib.pod_philox_logit_normals(ptr(out), draws, n_anchor, K, seed, image, level, stream_ptr()),
          "pod_philox_logit_normals")
    _count()
    return out


def philox_box_normals(anchor_ids, draws, seed, image):
    lib = _cabi.require_device()
    ids = _chk(anchor_ids, torch.int64, "anchor_ids")
    out = torch.empty((draws, ids.numel(), 4), dtype=torch.float32, device="cuda")
    check(lib.pod_philox_box_normals(ptr(out), ptr(ids), ids.numel(), draws, seed, image, stream_ptr()),
          "pod_philox_box_normals")
    _count()
    return out


# ---------------------------------------------------------------------------------------------------
def nchw_to_nhwc_split(x, scale):
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    NB, Cn, H, W = x.shape
    hi = torch.empty((NB, H, W, Cn), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    check(lib.pod_nchw_to_nhwc_split(ptr(x), NB, Cn, H, W, scale, ptr(hi), ptr(lo), stream_ptr()), "pod_nchw_to_nhwc_split")
    _count()
    return hi, lo


def feature_scale_dev(feats, max_scale, target=32768.0):
    """fp16 split scale of the input maps, computed on the device: min(max_scale, largest power of two s with
    max|x| * s <= target).  Returns a 1-element fp32 CUDA tensor (no host synchronisation); non-finite inputs set
    the library status word (check_status raises)."""
    lib = _cabi.require_device()
    amax = torch.zeros((1,), dtype=torch.int32, device=feats[0].device)
    scale = torch.empty((1,), dtype=torch.float32, device=feats[0].device)
    for f in feats:
        _chk(f, torch.float32, "feature map")
        check(lib.pod_absmax_accumulate(ptr(f), f.numel(), ptr(amax), stream_ptr()), "pod_absmax_accumulate")
    check(lib.pod_pow2_scale_from_absmax(ptr(amax), float(max_scale), float(target), ptr(scale), stream_ptr()),
          "pod_pow2_scale_from_absmax")
    _count(len(feats) + 1)
    return scale


def nchw_to_nhwc_split_dev(x, scale_dev):
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    _chk(scale_dev, torch.float32, "scale_dev")
    NB, Cn, H, W = x.shape
    hi = torch.empty((NB, H, W, Cn), dtype=torch.float16, device=x.device)
    lo = torch.empty_like(hi)
    check(lib.pod_nchw_to_nhwc_split_dev(ptr(x), NB, Cn, H, W, ptr(scale_dev), ptr(hi), ptr(lo), stream_ptr()),
          "pod_nchw_to_nhwc_split_dev")
    _count()
    return hi, lo


def nchw_to_nhwc_f32(x):
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    NB, Cn, H, W = x.shape
    out = torch.empty((NB, H, W, Cn), dtype=torch.float32, device=x.device)
    check(lib.pod_nchw_to_nhwc_f32(ptr(x), NB, Cn, H, W, ptr(out), stream_ptr()), "pod_nchw_to_nhwc_f32")
    _count()
    return out


def pack_conv_weight(w, cout_pad, scale):
    lib = _cabi.require_device()
    _chk(w, torch.float32, "w")
    cout, cin = w.shape[0], w.shape[1]
    assert tuple(w.shape[2:]) == (3, 3)
    # one buffer, hi rows then lo rows: the weights-as-A kernel fetches the stacked [hi; lo] tile with one TMA box
    both = torch.empty((2 * cout_pad, 9 * cin), dtype=torch.float16, device=w.device)
    hi, lo = both[:cout_pad], both[cout_pad:]
    check(lib.pod_pack_conv_weight(ptr(w), cout, cin, cout_pad, scale, ptr(hi), ptr(lo), stream_ptr()), "pod_pack_conv_weight")
    _count()
    return hi, lo


def pack_conv_weight_f32(w, cout_pad):
    lib = _cabi.require_device()
    _chk(w, torch.float32, "w")
    cout, cin = w.shape[0], w.shape[1]
    out = torch.empty((9 * cin, cout_pad), dtype=torch.float32, device=w.device)
    check(lib.pod_pack_conv_weight_f32(ptr(w), cout, cin, cout_pad, ptr(out), stream_ptr()), "pod_pack_conv_weight_f32")
    _count()
    return out


def mask_expand_split(x, drop, scale, out_hi=None, out_lo=None, live_reps=0, scale_dev=None):
    """x (NB, HW, C) fp32 -> (NB*samples*passes, HW, C) fp16 split pair.  live_reps > 0: only the first
    live_reps copies of every image are written (the others are never read, see conv3x3_tc map_live)."""
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    NB, HW, Cn = x.shape
    reps = drop.samples * drop.passes
    if out_hi is None:
        out_hi = torch.empty((NB * reps, HW, Cn), dtype=torch.float16, device=x.device)
        out_lo = torch.empty_like(out_hi)
    assert out_hi.numel() >= NB * reps * HW * Cn and out_lo.numel() >= NB * reps * HW * Cn
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.pod_mask_expand_split(ptr(x), NB, HW, Cn, C.byref(drop), scale, ptr(out_hi), ptr(out_lo), int(live_reps),
                                    ptr(scale_dev), stream_ptr()), "pod_mask_expand_split")
    if PROFILE is not None:
        e1.record()
        PROFILE.append((e0, e1, 4.0 * NB * HW * Cn * (1 + (live_reps or reps)), "mask_expand"))   # 1 read + live writes
    _count()
    return out_hi, out_lo


def conv3x3_tc(in_hi, in_lo, in_scale, NB, H, W, Cin, w_hi, w_lo, w_scale, bias, Cout, Cout_pad, mode, relu,
               out_hi=None, out_lo=None, out_scale=1.0, out_f32=None, out_map_stride=0, out_pixel_stride=0,
               drop=None, in_map_stride=None, in_offset=0, out_offset=0, out2_f32=None, out2_offset=0, split_col=0,
               out2_map_stride=0, out2_pixel_stride=0, map_group=0, map_live=0, in_scale_dev=None, out_scale_dev=None,
               q1=None, tag=None, mask_in=None, drop_scale_only=False):
    """Raw-pointer launch of the tcgen05 convolution. `in_offset`/`out_offset` are ELEMENT offsets
    into in_hi/in_lo and out_f32.  in_scale_dev: 1-element fp32 CUDA tensor replacing in_scale."""
    lib = _cabi.require_device()
    a = ConvArgs()
    esz = 2
    a.in_hi = in_hi.data_ptr() + in_offset * esz
    a.in_lo = in_lo.data_ptr() + in_offset * esz
    a.in_map_stride = in_map_stride if in_map_stride is not None else H * W * Cin
    a.in_scale = in_scale
    a.NB, a.H, a.W, a.Cin = NB, H, W, Cin
    a.w_hi, a.w_lo, a.w_scale = w_hi.data_ptr(), w_lo.data_ptr(), w_scale
    a.bias = bias.data_ptr()
    a.Cout, a.Cout_pad, a.mode, a.relu = Cout, Cout_pad, mode, int(relu)
    a.out_hi = out_hi.data_ptr() if out_hi is not None else None
    a.out_lo = out_lo.data_ptr() if out_lo is not None else None
    a.out_scale = out_scale
    a.out_f32 = (out_f32.data_ptr() + out_offset * 4) if out_f32 is not None else None
    a.out_map_stride, a.out_pixel_stride = out_map_stride, out_pixel_stride
    a.drop = drop if drop is not None else make_dropout()
    if out2_f32 is not None:
        a.out2_f32 = out2_f32.data_ptr() + out2_offset * 4
        a.split_col, a.out2_map_stride, a.out2_pixel_stride = split_col, out2_map_stride, out2_pixel_stride
    a.map_group, a.map_live = int(map_group), int(map_live)
    a.in_scale_dev = in_scale_dev.data_ptr() if in_scale_dev is not None else None
    a.out_scale_dev = out_scale_dev.data_ptr() if out_scale_dev is not None else None
    if mask_in is not None:
        a.mask_in, a.mask_in_layer = 1, int(mask_in)       # in-kernel dropout of the input (include/podb200.h)
    a.drop_scale_only = int(bool(drop_scale_only))
    if q1 is not None:
        # Q1 sample accumulation of the last tower layer (include/podb200.h, pod_conv_args.q1_acc)
        a.q1_acc = q1["acc"].data_ptr()
        a.q1_samples, a.q1_passes = int(q1["samples"]), int(q1["passes"])
        a.q1_live[0], a.q1_live[1] = int(q1["live"][0]), int(q1["live"][1] if len(q1["live"]) > 1 else 0)
        a.q1_acc_mask, a.q1_group = int(q1["mask"]), int(q1["group"])
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.pod_conv3x3_tc(C.byref(a), stream_ptr()), "pod_conv3x3_tc")
    if PROFILE is not None:
        e1.record()
        user_tag = tag
        tag = "tower256" if (Cout_pad == 256 and mode == POD_OUT_HIDDEN) else ("conv1" if Cout_pad == 256 else "out")
        tag = user_tag or tag
        if q1 is not None:
            tag = "tower256_q1"
        maps = NB if not map_group else NB // map_group * map_live          # maps actually evaluated
        if q1 is not None:
            maps = NB // (a.q1_samples * a.q1_passes) * sum(int(v) for v in q1["live"][:a.q1_passes])
        PROFILE.append((e0, e1, 2.0 * 9 * Cin * Cout * maps * H * W, tag))
    _count()


def conv3x3_tc_status():
    lib = _cabi.require_device()
    v = C.c_int(0)
    check(lib.pod_conv3x3_tc_status(C.byref(v)), "pod_conv3x3_tc_status")
    return v.value


STATUS_TEXT = {100: "a hidden tower activation left the fp16 split range (|x| * 16 > 65504): the checkpoint / input "
                    "produces activations this path cannot represent",
               101: "a first-layer tower activation left the fp16 split range (|x| * 16 / (1-p) > 65504)",
               102: "non-finite values in the input feature maps",
               103: "a backbone feature map left the fp16 split range"}


def status():
    """Device-side error word of the library, read and cleared (synchronises the device)."""
    lib = _cabi.require_device()
    v = C.c_int(0)
    check(lib.pod_status(C.byref(v)), "pod_status")
    return v.value


def check_status():
    """Raise PodError if any kernel since the last check reported a device-side error (expired bounded barrier
    wait, activation outside the fp16 split range, non-finite input).  Called once per inference call."""
    v = status()
    if v != 0:
        raise _cabi.PodError("libpodb200 device-side error %d: %s -- the results of this call are invalid"
                             % (v, STATUS_TEXT.get(v, "a bounded mbarrier wait of the tcgen05 convolution expired "
                                                      "(wait code %d)" % v)))


def set_conv_wait_limit(cycles):
    check(_cabi.require_device().pod_conv3x3_tc_set_wait_limit(int(cycles)), "pod_conv3x3_tc_set_wait_limit")


def set_conv_debug_fault(on):
    check(_cabi.load_library().pod_conv3x3_tc_debug_fault(int(bool(on))), "pod_conv3x3_tc_debug_fault")


def conv_debug_clock(enable=None):
    """enable=True/False switches the in-kernel clock probe of the CTA-pair kernel; enable=None returns
    (cycles, ns, MHz) of the last CTA-pair launch."""
    lib = _cabi.require_device()
    if enable is not None:
        check(lib.pod_conv3x3_tc_debug_clock(int(bool(enable)), None), "pod_conv3x3_tc_debug_clock")
        return None
    out = (C.c_longlong * 2)()
    check(lib.pod_conv3x3_tc_debug_clock(0, out), "pod_conv3x3_tc_debug_clock")
    return out[0], out[1], (1e3 * out[0] / out[1]) if out[1] else float("nan")


def set_conv_kblock(bk):
    check(_cabi.load_library().pod_conv3x3_tc_set_kblock(int(bk)), "pod_conv3x3_tc_set_kblock")


def set_conv_pair(on):
    check(_cabi.load_library().pod_conv3x3_tc_set_pair(int(bool(on))), "pod_conv3x3_tc_set_pair")


def set_conv_tile_width(tw):
    """0 = per map shape (default), 16 or 32: pixel-tile width of the CTA-pair row-halo kernel (bit-identical results)."""
    check(_cabi.load_library().pod_conv3x3_tc_set_tile_width(int(tw)), "pod_conv3x3_tc_set_tile_width")


def set_conv_trunc_comp(ulps_per_mma):
    check(_cabi.load_library().pod_conv3x3_tc_set_trunc_comp(float(ulps_per_mma)), "pod_conv3x3_tc_set_trunc_comp")


def set_conv_wt(on):
    check(_cabi.load_library().pod_conv3x3_tc_set_wt(int(bool(on))), "pod_conv3x3_tc_set_wt")


def set_conv_halo(mode):
    """Row-halo operand staging: bit 0 = pixels-as-M kernels, bit 1 = weights-as-A kernel."""
    check(_cabi.load_library().pod_conv3x3_tc_set_halo(int(mode)), "pod_conv3x3_tc_set_halo")


def set_conv_chunk_kblocks(kb):
    check(_cabi.load_library().pod_conv3x3_tc_set_chunk_kblocks(int(kb)), "pod_conv3x3_tc_set_chunk_kblocks")


def set_conv_chunk_taps(taps):
    check(_cabi.load_library().pod_conv3x3_tc_set_chunk_taps(int(taps)), "pod_conv3x3_tc_set_chunk_taps")


def conv3x3_simt(x, w_kc, bias, Cout, Cout_pad, relu, drop=None, out=None, out_map_stride=None, out_pixel_stride=None,
                 out_offset=0):
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    NB, H, W, Cin = x.shape
    if out is None:
        out = torch.empty((NB, H * W, Cout), dtype=torch.float32, device=x.device)
        out_map_stride, out_pixel_stride = H * W * Cout, Cout
    check(lib.pod_conv3x3_simt(ptr(x), NB, H, W, Cin, ptr(w_kc), ptr(bias), Cout, Cout_pad, int(relu),
                               C.byref(drop) if drop is not None else None,
                               C.c_void_p(out.data_ptr() + out_offset * 4), out_map_stride, out_pixel_stride, stream_ptr()),
          "pod_conv3x3_simt")
    _count()
    return out


def sample_mean_q1(x):
    """x (B, S, ...) -> (B, ...) with the reference's Q1 weighting."""
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    B, S = x.shape[0], x.shape[1]
    n = x[0, 0].numel()
    out = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.pod_sample_mean_q1(ptr(x), B, S, n, ptr(out), stream_ptr()), "pod_sample_mean_q1")
    if PROFILE is not None:
        e1.record()
        PROFILE.append((e0, e1, 4.0 * B * n * (max(S - 1, 1) + 1), "sample_mean"))   # reads S-1 samples (Q1), writes 1
    _count()
    return out


def scores(logits, logvar, level_off, draws, seed, image0, runs=1):
    """logits/logvar (B,R,K) -> probs (B,R,K), score (B,R), cls (B,R) int32.  With runs > 1 row b is run
    b % runs of image image0 + b // runs (own noise draws per run)."""
    lib = _cabi.require_device()
    _chk(logits, torch.float32, "logits")
    if logvar is not None:
        _chk(logvar, torch.float32, "logvar")
    B, R, K = logits.shape
    probs = torch.empty_like(logits)
    score = torch.empty((B, R), dtype=torch.float32, device=logits.device)
    cls = torch.empty((B, R), dtype=torch.int32, device=logits.device)
    check(lib.pod_scores(ptr(logits), ptr(logvar), B, R, K, len(level_off) - 1, int_array(level_off), draws, seed, image0,
                         runs, ptr(probs), ptr(score), ptr(cls), stream_ptr()), "pod_scores")
    _count()
    return probs, score, cls


def seg_offsets(level_off, topk):
    seg = [0]
    for l in range(len(level_off) - 1):
        seg.append(seg[-1] + min(topk, level_off[l + 1] - level_off[l]))
    return seg


def topk_levels(score, level_off, topk, thresh):
    lib = _cabi.require_device()
    _chk(score, torch.float32, "score")
    B, R = score.shape
    seg = seg_offsets(level_off, topk)
    cap = seg[-1]
    cand_idx = torch.zeros((B, cap), dtype=torch.int32, device=score.device)
    cand_cnt = torch.zeros((B, len(level_off) - 1), dtype=torch.int32, device=score.device)
    check(lib.pod_topk_levels(ptr(score), B, R, len(level_off) - 1, int_array(level_off), int_array(seg), topk, thresh,
                              ptr(cand_idx), ptr(cand_cnt), stream_ptr()), "pod_topk_levels")
    _count()
    return cand_idx, cand_cnt, seg


def decode_cov(mean_delta, mean_regvar, sample_delta, anchors, probs, score, cls, cand_idx, cand_cnt, seg, box_draws,
               seed, image0, reg_weights, runs=1, sample_reg_weights=None):
    lib = _cabi.require_device()
    B, R, K = probs.shape
    cap = seg[-1]
    dev = probs.device
    out = {
        "boxes": torch.zeros((B, cap, 4), dtype=torch.float32, device=dev),
        "cov": torch.zeros((B, cap, 4, 4), dtype=torch.float32, device=dev),
        "scores": torch.zeros((B, cap), dtype=torch.float32, device=dev),
        "classes": torch.zeros((B, cap), dtype=torch.int32, device=dev),
        "probs": torch.zeros((B, cap, K), dtype=torch.float32, device=dev),
        "count": torch.zeros((B,), dtype=torch.int32, device=dev),
        "anchor": torch.zeros((B, cap), dtype=torch.int32, device=dev),
    }
    a = DecodeArgs()
    a.mean_delta = _chk(mean_delta, torch.float32, "mean_delta").data_ptr()
    a.mean_regvar = _chk(mean_regvar, torch.float32, "mean_regvar").data_ptr() if mean_regvar is not None else None
    a.cov_dims = mean_regvar.shape[-1] if mean_regvar is not None else 4
    a.sample_delta = _chk(sample_delta, torch.float32, "sample_delta").data_ptr() if sample_delta is not None else None
    a.S = sample_delta.shape[1] if sample_delta is not None else 1
    a.anchors = _chk(anchors, torch.float32, "anchors").data_ptr()
    a.probs, a.score, a.cls = probs.data_ptr(), score.data_ptr(), cls.data_ptr()
    a.cand_idx, a.cand_cnt = cand_idx.data_ptr(), cand_cnt.data_ptr()
    a.B, a.R, a.K, a.n_levels, a.cap = B, R, K, len(seg) - 1, cap
    seg_arr = int_array(seg)
    a.seg_off_host = seg_arr
    a.box_draws, a.seed, a.image0 = box_draws, int(seed) & 0xFFFFFFFFFFFFFFFF, image0
    a.runs = runs
    a.wx, a.wy, a.ww, a.wh = [float(v) for v in reg_weights]
    a.swx, a.swy, a.sww, a.swh = [float(v) for v in (sample_reg_weights or reg_weights)]
    a.out_boxes, a.out_cov = out["boxes"].data_ptr(), out["cov"].data_ptr()
    a.out_scores, a.out_classes = out["scores"].data_ptr(), out["classes"].data_ptr()
    a.out_probs, a.out_count, a.out_anchor = out["probs"].data_ptr(), out["count"].data_ptr(), out["anchor"].data_ptr()
    check(lib.pod_decode_cov(C.byref(a), stream_ptr()), "pod_decode_cov")
    _count()
    out["has_cov"] = (mean_regvar is not None) or (sample_delta is not None and a.S > 1)
    return out


NMS_VANILLA, NMS_TRICK, NMS_AUTO = 0, 1, 2


def nms_fuse(cand, mode, nms_thresh, affinity, max_dets, in_hw, out_hw, nms_variant=NMS_AUTO, box_merge=0, cls_merge=0,
             skip_post=False):
    """cand: dict with boxes (B,cap,4), cov (B,cap,4,4), scores, classes (int32), probs, count, has_cov."""
    lib = _cabi.require_device()
    B, cap, K = cand["probs"].shape
    dev = cand["probs"].device
    out = {
        "boxes": torch.zeros((B, max_dets, 4), dtype=torch.float32, device=dev),
        "cov": torch.zeros((B, max_dets, 4, 4), dtype=torch.float32, device=dev),
        "scores": torch.zeros((B, max_dets), dtype=torch.float32, device=dev),
        "classes": torch.zeros((B, max_dets), dtype=torch.int32, device=dev),
        "probs": torch.zeros((B, max_dets, K), dtype=torch.float32, device=dev),
        "count": torch.zeros((B,), dtype=torch.int32, device=dev),
        "keep": torch.zeros((B, max_dets), dtype=torch.int32, device=dev),
        "keep_count": torch.zeros((B,), dtype=torch.int32, device=dev),
        "src": torch.zeros((B, max_dets), dtype=torch.int32, device=dev),
    }
    a = NmsArgs()
    a.boxes = _chk(cand["boxes"], torch.float32, "boxes").data_ptr()
    a.cov = _chk(cand["cov"], torch.float32, "cov").data_ptr()
    a.scores = _chk(cand["scores"], torch.float32, "scores").data_ptr()
    a.classes = _chk(cand["classes"], torch.int32, "classes").data_ptr()
    a.probs = _chk(cand["probs"], torch.float32, "probs").data_ptr()
    a.count = _chk(cand["count"], torch.int32, "count").data_ptr()
    a.B, a.cap, a.K = B, cap, K
    a.has_cov = int(bool(cand.get("has_cov", True)))
    a.mode, a.nms_variant, a.box_merge, a.cls_merge = mode, nms_variant, box_merge, cls_merge
    a.nms_thresh, a.affinity, a.max_dets = float(nms_thresh), float(affinity), max_dets
    a.in_h, a.in_w, a.out_h, a.out_w = int(in_hw[0]), int(in_hw[1]), int(out_hw[0]), int(out_hw[1])
    a.det_boxes, a.det_cov = out["boxes"].data_ptr(), out["cov"].data_ptr()
    a.det_scores, a.det_classes = out["scores"].data_ptr(), out["classes"].data_ptr()
    a.det_probs, a.det_count = out["probs"].data_ptr(), out["count"].data_ptr()
    a.keep, a.keep_count = out["keep"].data_ptr(), out["keep_count"].data_ptr()
    a.det_src = out["src"].data_ptr()
    a.skip_post = int(bool(skip_post))
    check(lib.pod_nms_fuse(C.byref(a), stream_ptr()), "pod_nms_fuse")
    _count()
    return out


def cluster_merge(det, runs, affinity):
    """Per-run detections (B*runs rows, from nms_fuse(skip_post=True)) -> clustered candidate set per image
    (reference general_black_box_ensembles_post_processing up to its final NMS)."""
    lib = _cabi.require_device()
    BR, D, K = det["probs"].shape
    assert BR % runs == 0
    B = BR // runs
    cap = runs * D
    dev = det["probs"].device
    out = {
        "boxes": torch.zeros((B, cap, 4), dtype=torch.float32, device=dev),
        "cov": torch.zeros((B, cap, 4, 4), dtype=torch.float32, device=dev),
        "scores": torch.zeros((B, cap), dtype=torch.float32, device=dev),
        "classes": torch.zeros((B, cap), dtype=torch.int32, device=dev),
        "probs": torch.zeros((B, cap, K), dtype=torch.float32, device=dev),
        "count": torch.zeros((B,), dtype=torch.int32, device=dev),
        "has_cov": True,
    }
    a = MergeArgs()
    a.det_boxes, a.det_cov = det["boxes"].data_ptr(), det["cov"].data_ptr()
    a.det_probs, a.det_classes, a.det_count = det["probs"].data_ptr(), det["classes"].data_ptr(), det["count"].data_ptr()
    a.B, a.runs, a.max_dets, a.K = B, runs, D, K
    a.affinity = float(affinity)
    a.out_boxes, a.out_cov, a.out_scores = out["boxes"].data_ptr(), out["cov"].data_ptr(), out["scores"].data_ptr()
    a.out_classes, a.out_probs, a.out_count = out["classes"].data_ptr(), out["probs"].data_ptr(), out["count"].data_ptr()
    seeds = torch.empty((B, cap), dtype=torch.int32, device=dev)
    a.seed_scratch = seeds.data_ptr()
    check(lib.pod_cluster_merge(C.byref(a), stream_ptr()), "pod_cluster_merge")
    _count(2)
    return out


def record_width(max_dets, K):
    return 1 + max_dets * (4 + 1 + 1 + K + 16)


def wire_records(det, xywh=False, cat_map=None, out=None):
    """Detections of a batch (dict of nms_fuse) -> (B, 1 + max_dets*(22+K)) fp32 records in one launch.
    xywh=False: the all-gather record (xyxy boxes, covariance as is); xywh=True: the reference's JSON layout
    (XYWH boxes, T Sigma T^T).  cat_map: int32 CUDA tensor (K,) contiguous class -> dataset category id (-1 = none)."""
    lib = _cabi.require_device()
    B, D, K = det["probs"].shape
    if out is None:
        out = torch.empty((B, record_width(D, K)), dtype=torch.float32, device=det["probs"].device)
    a = WireArgs()
    a.det_boxes = _chk(det["boxes"], torch.float32, "boxes").data_ptr()
    a.det_cov = _chk(det["cov"], torch.float32, "cov").data_ptr()
    a.det_scores = _chk(det["scores"], torch.float32, "scores").data_ptr()
    a.det_classes = _chk(det["classes"], torch.int32, "classes").data_ptr()
    a.det_probs = _chk(det["probs"], torch.float32, "probs").data_ptr()
    a.det_count = _chk(det["count"], torch.int32, "count").data_ptr()
    a.B, a.max_dets, a.K, a.xywh = B, D, K, int(bool(xywh))
    a.cat_map = _chk(cat_map, torch.int32, "cat_map").data_ptr() if cat_map is not None else None
    if cat_map is not None and cat_map.numel() != K:
        raise _cabi.PodError("cat_map must have one entry per class")
    a.records = _chk(out, torch.float32, "records").data_ptr()
    check(lib.pod_wire_records(C.byref(a), stream_ptr()), "pod_wire_records")
    _count()
    return out


def q1_finish(acc, n_maps, groups, n, samples, scale_dev, out_hi, out_lo):
    """Partial sample sums of the last tower layer -> mean activation maps as fp16 split pair (pod_q1_finish)."""
    lib = _cabi.require_device()
    _chk(acc, torch.float32, "acc")
    assert acc.numel() >= n_maps * groups * n and out_hi.numel() >= n_maps * n and out_lo.numel() >= n_maps * n
    check(lib.pod_q1_finish(ptr(acc), int(n_maps), int(groups), int(n), int(samples), 1.0, ptr(scale_dev), ptr(out_hi), ptr(out_lo),
                            stream_ptr()), "pod_q1_finish")
    _count()


# --------------------------------------------------------------------------------------------------- backbone ops
def pack_conv_weight_k(w, cout_pad, scale):
    """(Cout, Cin, k, k) fp32 -> [cout_pad][k*k*Cin] split pair (hi rows then lo rows in one buffer)."""
    lib = _cabi.require_device()
    _chk(w, torch.float32, "w")
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    both = torch.empty((2 * cout_pad, k * k * cin), dtype=torch.float16, device=w.device)
    hi, lo = both[:cout_pad], both[cout_pad:]
    check(lib.pod_pack_conv_weight_k(ptr(w), cout, cin, k, cout_pad, scale, ptr(hi), ptr(lo), stream_ptr()), "pod_pack_conv_weight_k")
    _count()
    return hi, lo


def conv_tc_general(x_hi, x_lo, in_scale, NB, Hin, Win, Cin, ksize, stride, w_hi, w_lo, w_scale, cout_rows, cout, bias, relu,
                    block_cols, out_hi=None, out_lo=None, out_scale=1.0, out_ch_stride=None, res=None, res_scale=1.0, out_f32=None):
    """One convolution of the backbone = one tcgen05 launch per column block of `block_cols` output channels."""
    lib = _cabi.require_device()
    pad = ksize // 2
    Hout, Wout = (Hin + 2 * pad - ksize) // stride + 1, (Win + 2 * pad - ksize) // stride + 1
    a = ConvGArgs()
    a.in_hi, a.in_lo, a.in_scale = x_hi.data_ptr(), x_lo.data_ptr(), float(in_scale)
    a.NB, a.Hin, a.Win, a.Cin, a.ksize, a.stride, a.Hout, a.Wout = NB, Hin, Win, Cin, ksize, stride, Hout, Wout
    a.w_hi, a.w_lo, a.w_scale = w_hi.data_ptr(), w_lo.data_ptr(), float(w_scale)
    a.Cout_rows, a.Cout, a.block_cols = cout_rows, cout, block_cols
    a.bias, a.relu = bias.data_ptr(), int(bool(relu))
    a.out_hi = out_hi.data_ptr() if out_hi is not None else None
    a.out_lo = out_lo.data_ptr() if out_lo is not None else None
    a.out_scale = float(out_scale)
    a.out_ch_stride = int(out_ch_stride if out_ch_stride is not None else cout_rows)
    if res is not None:
        a.res_hi, a.res_lo, a.res_scale = res[0].data_ptr(), res[1].data_ptr(), float(res_scale)
    a.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for col0 in range(0, cout_rows, block_cols):
        if col0 >= cout:
            break
        a.col0 = col0
        check(lib.pod_conv_tc_general(C.byref(a), stream_ptr()), "pod_conv_tc_general")
        _count()
    if PROFILE is not None:
        e1.record()
        PROFILE.append((e0, e1, 2.0 * ksize * ksize * Cin * cout * NB * Hout * Wout, "backbone_conv"))
    return Hout, Wout


def stem_conv7_pool(images, H, W, mean, std, w147x64, bias, scratch, out_hi, out_lo, out_scale):
    lib = _cabi.require_device()
    NB, _, Himg, Wimg = images.shape
    assert images.is_cuda and images.is_contiguous() and images.dtype in (torch.uint8, torch.float32)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    sd = (C.c_float * 3)(*[float(v) for v in std])
    check(lib.pod_stem_conv7_pool(ptr(images), int(images.dtype == torch.uint8), NB, Himg, Wimg, H, W, m, sd, ptr(w147x64), ptr(bias),
                                  ptr(scratch), ptr(out_hi), ptr(out_lo), float(out_scale), stream_ptr()), "pod_stem_conv7_pool")
    _count(2)


def upsample2_add(dst, src):
    """dst (NB,H,W,C) fp32 += nearest x2 upsample of src (NB,ceil(H/2),ceil(W/2),C)."""
    lib = _cabi.require_device()
    NB, H, W, Cn = dst.shape
    assert tuple(src.shape) == (NB, (H + 1) // 2, (W + 1) // 2, Cn)
    check(lib.pod_upsample2_add(ptr(_chk(dst, torch.float32, "dst")), ptr(_chk(src, torch.float32, "src")), NB, H, W, Cn, stream_ptr()),
          "pod_upsample2_add")
    _count()


def split_f32(x, scale=1.0, scale_dev=None, relu=False, out_hi=None, out_lo=None):
    """fp32 tensor (any shape, numel % 8 == 0) -> fp16 split pair of x * scale (or * *scale_dev)."""
    lib = _cabi.require_device()
    _chk(x, torch.float32, "x")
    if out_hi is None:
        out_hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        out_lo = torch.empty_like(out_hi)
    check(lib.pod_split_f32(ptr(x), x.numel(), float(scale), ptr(scale_dev), int(bool(relu)), ptr(out_hi), ptr(out_lo), stream_ptr()),
          "pod_split_f32")
    _count()
    return out_hi, out_lo


def q1_mean_act(act_hi, act_lo, images, samples, passes, mask, live, n, scale_dev, out_hi, out_lo):
    """Per-sample split-pair maps of the last tower layer -> Q1-weighted mean maps of the passes in `mask` (pod_q1_mean_act)."""
    lib = _cabi.require_device()
    n_acc = bin(mask).count("1")
    assert act_hi.numel() >= images * samples * passes * n and out_hi.numel() >= images * n_acc * n
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.pod_q1_mean_act(ptr(act_hi), ptr(act_lo), int(images), int(samples), int(passes), int(mask), int_array(list(live) + [0]),
                              int(n), 1.0, ptr(scale_dev), ptr(out_hi), ptr(out_lo), stream_ptr()), "pod_q1_mean_act")
    if PROFILE is not None:
        e1.record()
        reads = sum(int(live[p]) for p in range(passes) if (mask >> p) & 1)
        PROFILE.append((e0, e1, 4.0 * images * n * (reads + n_acc), "act_mean"))       # bytes: split pairs read + written
    _count()
