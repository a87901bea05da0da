"""ctypes binding of libpodb200.so (include/podb200.h).  PyTorch tensors cross the boundary as
data_ptr() + shapes + the current CUDA stream; every call checks the int return code and raises.
There is no CPU fallback: a missing library or a non-Blackwell device is an error."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpodb200.so")

EXPORTS = [
    "pod_last_error", "pod_version", "pod_device_ok", "pod_status",
    "pod_absmax_accumulate", "pod_pow2_scale_from_absmax", "pod_nchw_to_nhwc_split_dev",
    "pod_conv3x3_tc_set_wait_limit", "pod_conv3x3_tc_debug_fault", "pod_conv3x3_tc_debug_clock",
    "pod_philox_dropout_mask", "pod_philox_logit_normals", "pod_philox_box_normals",
    "pod_nchw_to_nhwc_split", "pod_nchw_to_nhwc_f32", "pod_pack_conv_weight", "pod_pack_conv_weight_f32",
    "pod_mask_expand_split", "pod_conv3x3_tc", "pod_conv3x3_tc_set_kblock", "pod_conv3x3_tc_set_chunk_taps", "pod_conv3x3_tc_set_chunk_kblocks", "pod_conv3x3_tc_set_pair", "pod_conv3x3_tc_set_tile_width", "pod_conv3x3_tc_set_halo", "pod_conv3x3_tc_set_wt", "pod_conv3x3_tc_set_trunc_comp", "pod_conv3x3_tc_status",
    "pod_conv3x3_simt", "pod_sample_mean_q1", "pod_scores", "pod_topk_levels", "pod_decode_cov", "pod_nms_fuse",
    "pod_cluster_merge", "pod_wire_records", "pod_q1_finish", "pod_q1_mean_act",
    "pod_conv_tc_general", "pod_pack_conv_weight_k", "pod_stem_conv7_pool", "pod_upsample2_add", "pod_split_f32",
]

POD_OUT_HIDDEN, POD_OUT_RAW = 0, 1


class PodError(RuntimeError):
    pass


class Dropout(C.Structure):
    _fields_ = [("p", C.c_double), ("seed", C.c_uint64), ("image0", C.c_int), ("samples", C.c_int),
                ("passes", C.c_int), ("pass0", C.c_int), ("tower", C.c_int), ("layer", C.c_int), ("level", C.c_int)]


class ConvArgs(C.Structure):
    _fields_ = [("in_hi", C.c_void_p), ("in_lo", C.c_void_p), ("in_map_stride", C.c_int64), ("in_scale", C.c_float),
                ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cin", C.c_int),
                ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_scale", C.c_float), ("bias", C.c_void_p),
                ("Cout", C.c_int), ("Cout_pad", C.c_int), ("mode", C.c_int), ("relu", C.c_int),
                ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_scale", C.c_float), ("out_f32", C.c_void_p),
                ("out_map_stride", C.c_int64), ("out_pixel_stride", C.c_int64), ("drop", Dropout),
                ("out2_f32", C.c_void_p), ("split_col", C.c_int), ("out2_map_stride", C.c_int64),
                ("out2_pixel_stride", C.c_int64), ("map_group", C.c_int), ("map_live", C.c_int),
                ("in_scale_dev", C.c_void_p), ("out_scale_dev", C.c_void_p),
                ("q1_acc", C.c_void_p), ("q1_samples", C.c_int), ("q1_passes", C.c_int), ("q1_live", C.c_int * 2),
                ("q1_acc_mask", C.c_int), ("q1_group", C.c_int), ("mask_in", C.c_int), ("mask_in_layer", C.c_int),
                ("drop_scale_only", C.c_int)]


class DecodeArgs(C.Structure):
    _fields_ = [("mean_delta", C.c_void_p), ("mean_regvar", C.c_void_p), ("cov_dims", C.c_int),
                ("sample_delta", C.c_void_p), ("S", C.c_int), ("anchors", C.c_void_p), ("probs", C.c_void_p),
                ("score", C.c_void_p), ("cls", C.c_void_p), ("cand_idx", C.c_void_p), ("cand_cnt", C.c_void_p),
                ("B", C.c_int), ("R", C.c_int), ("K", C.c_int), ("n_levels", C.c_int), ("cap", C.c_int),
                ("seg_off_host", C.POINTER(C.c_int)), ("box_draws", C.c_int), ("seed", C.c_uint64),
                ("image0", C.c_int), ("runs", C.c_int), ("wx", C.c_float), ("wy", C.c_float), ("ww", C.c_float), ("wh", C.c_float),
                ("out_boxes", C.c_void_p), ("out_cov", C.c_void_p), ("out_scores", C.c_void_p),
                ("out_classes", C.c_void_p), ("out_probs", C.c_void_p), ("out_count", C.c_void_p),
                ("out_anchor", C.c_void_p), ("swx", C.c_float), ("swy", C.c_float), ("sww", C.c_float), ("swh", C.c_float)]


class NmsArgs(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("cov", C.c_void_p), ("scores", C.c_void_p), ("classes", C.c_void_p),
                ("probs", C.c_void_p), ("count", C.c_void_p), ("B", C.c_int), ("cap", C.c_int), ("K", C.c_int),
                ("has_cov", C.c_int), ("mode", C.c_int), ("nms_variant", C.c_int), ("box_merge", C.c_int),
                ("cls_merge", C.c_int), ("nms_thresh", C.c_double), ("affinity", C.c_double), ("max_dets", C.c_int),
                ("in_h", C.c_int), ("in_w", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("det_boxes", C.c_void_p), ("det_cov", C.c_void_p), ("det_scores", C.c_void_p),
                ("det_classes", C.c_void_p), ("det_probs", C.c_void_p), ("det_count", C.c_void_p),
                ("keep", C.c_void_p), ("keep_count", C.c_void_p), ("det_src", C.c_void_p), ("skip_post", C.c_int)]


class MergeArgs(C.Structure):
    _fields_ = [("det_boxes", C.c_void_p), ("det_cov", C.c_void_p), ("det_probs", C.c_void_p), ("det_classes", C.c_void_p),
                ("det_count", C.c_void_p), ("B", C.c_int), ("runs", C.c_int), ("max_dets", C.c_int), ("K", C.c_int),
                ("affinity", C.c_double), ("out_boxes", C.c_void_p), ("out_cov", C.c_void_p), ("out_scores", C.c_void_p),
                ("out_classes", C.c_void_p), ("out_probs", C.c_void_p), ("out_count", C.c_void_p), ("seed_scratch", C.c_void_p)]


class ConvGArgs(C.Structure):
    _fields_ = [("in_hi", C.c_void_p), ("in_lo", C.c_void_p), ("in_scale", C.c_float),
                ("NB", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Cin", C.c_int), ("ksize", C.c_int), ("stride", C.c_int),
                ("Hout", C.c_int), ("Wout", C.c_int), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p), ("w_scale", C.c_float),
                ("Cout_rows", C.c_int), ("Cout", C.c_int), ("col0", C.c_int), ("block_cols", C.c_int), ("bias", C.c_void_p),
                ("relu", C.c_int), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_scale", C.c_float),
                ("out_ch_stride", C.c_int), ("res_hi", C.c_void_p), ("res_lo", C.c_void_p), ("res_scale", C.c_float),
                ("out_f32", C.c_void_p)]


class WireArgs(C.Structure):
    _fields_ = [("det_boxes", C.c_void_p), ("det_cov", C.c_void_p), ("det_scores", C.c_void_p), ("det_classes", C.c_void_p),
                ("det_probs", C.c_void_p), ("det_count", C.c_void_p), ("B", C.c_int), ("max_dets", C.c_int), ("K", C.c_int),
                ("xywh", C.c_int), ("cat_map", C.c_void_p), ("records", C.c_void_p)]


_lib = None


def load_library():
    """dlopen the in-tree library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build in-tree when the sources changed or the library is absent and nvcc is available; the
    # check is a hash of csrc/ and costs nothing when the library is current
    try:
        from . import build as _build
        _build.build()
    except Exception as e:  # noqa: BLE001
        if not os.path.exists(LIB_PATH):
            raise PodError("libpodb200.so is missing (%s) and could not be built (%s): run `python -m "
                           "pod_compare_b200.build` -- there is no CPU or PyTorch fallback for this path" % (LIB_PATH, e))
    if not os.path.exists(LIB_PATH):
        raise PodError("libpodb200.so is missing (%s): run `python -m pod_compare_b200.build` -- there is no "
                       "CPU or PyTorch fallback for this path" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.pod_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise PodError("libpodb200.so does not export %s" % name)
    lib.pod_philox_dropout_mask.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
    lib.pod_philox_logit_normals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    lib.pod_philox_box_normals.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
    lib.pod_nchw_to_nhwc_split.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_nchw_to_nhwc_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.pod_pack_conv_weight.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_pack_conv_weight_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.pod_mask_expand_split.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Dropout), C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.pod_conv3x3_tc.argtypes = [C.POINTER(ConvArgs), C.c_void_p]
    lib.pod_conv3x3_tc_set_kblock.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_chunk_taps.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_chunk_kblocks.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_pair.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_tile_width.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_halo.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_wt.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_set_trunc_comp.argtypes = [C.c_float]
    lib.pod_conv3x3_tc_status.argtypes = [C.POINTER(C.c_int)]
    lib.pod_status.argtypes = [C.POINTER(C.c_int)]
    lib.pod_conv3x3_tc_set_wait_limit.argtypes = [C.c_longlong]
    lib.pod_conv3x3_tc_debug_fault.argtypes = [C.c_int]
    lib.pod_conv3x3_tc_debug_clock.argtypes = [C.c_int, C.POINTER(C.c_longlong)]
    lib.pod_absmax_accumulate.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.pod_pow2_scale_from_absmax.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.pod_nchw_to_nhwc_split_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_conv3x3_simt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(Dropout), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.pod_sample_mean_q1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    lib.pod_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int,
                               C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_topk_levels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                    C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_decode_cov.argtypes = [C.POINTER(DecodeArgs), C.c_void_p]
    lib.pod_nms_fuse.argtypes = [C.POINTER(NmsArgs), C.c_void_p]
    lib.pod_cluster_merge.argtypes = [C.POINTER(MergeArgs), C.c_void_p]
    lib.pod_wire_records.argtypes = [C.POINTER(WireArgs), C.c_void_p]
    lib.pod_conv_tc_general.argtypes = [C.POINTER(ConvGArgs), C.c_void_p]
    lib.pod_pack_conv_weight_k.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_stem_conv7_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                        C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_void_p]
    lib.pod_upsample2_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.pod_split_f32.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_q1_mean_act.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int64, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pod_q1_finish.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise PodError("%s failed (rc=%d): %s" % (what, rc, load_library().pod_last_error().decode()))


def require_device():
    """The product path runs only on a Blackwell GPU with the native library loaded."""
    lib = load_library()
    if not torch.cuda.is_available():
        raise PodError("no CUDA device: the probabilistic-inference path has no CPU fallback")
    if not lib.pod_device_ok():
        raise PodError("device is not compute capability 10.x (sm_100a kernels only)")
    return lib


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def int_array(values):
    return (C.c_int * len(values))(*[int(v) for v in values])
