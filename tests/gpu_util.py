"""Helpers shared by the GPU parity tests (they call the product through the C ABI wrappers in
pod_compare_b200.ops and check against oracle/podref.py and tests/golden)."""
import numpy as np
import torch

from pod_compare_b200 import engine, ops
from pod_compare_b200._cabi import POD_OUT_HIDDEN, POD_OUT_RAW

ACT = engine.ACT_SCALE


def conv_ref64(x, w, b, relu):
    y = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    return torch.relu(y) if relu else y


def tc_conv_raw(x_nchw, w, b, relu, cout_pad=None):
    """x (NB,Cin,H,W) fp32 CPU -> (NB,Cout,H,W) fp32 CPU through the tcgen05 kernel (RAW mode)."""
    NB, Cin, H, W = x_nchw.shape
    hi, lo = ops.nchw_to_nhwc_split(x_nchw.cuda().contiguous(), ACT)
    pcv = engine.pack_conv(w, b, "cuda", cout_pad=cout_pad)
    out = torch.full((NB, H * W, pcv.cout), float("nan"), dtype=torch.float32, device="cuda")
    ops.conv3x3_tc(hi, lo, ACT, NB, H, W, Cin, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, pcv.cout, pcv.cout_pad,
                   POD_OUT_RAW, relu, out_f32=out, out_map_stride=H * W * pcv.cout, out_pixel_stride=pcv.cout)
    torch.cuda.synchronize()
    assert ops.conv3x3_tc_status() == 0, "tcgen05 kernel reported an expired barrier wait"
    return out.view(NB, H, W, pcv.cout).permute(0, 3, 1, 2).cpu()


def tc_conv_hidden(x_nchw, w, b, drop):
    """HIDDEN mode (ReLU + optional dropout, fp16 split output) -> reconstructed fp32 (NB,C,H,W) CPU."""
    NB, Cin, H, W = x_nchw.shape
    hi, lo = ops.nchw_to_nhwc_split(x_nchw.cuda().contiguous(), ACT)
    pcv = engine.pack_conv(w, b, "cuda")
    assert pcv.cout == pcv.cout_pad
    ohi = torch.zeros((NB, H, W, pcv.cout), dtype=torch.float16, device="cuda")
    olo = torch.zeros_like(ohi)
    ops.conv3x3_tc(hi, lo, ACT, NB, H, W, Cin, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, pcv.cout, pcv.cout_pad,
                   POD_OUT_HIDDEN, True, out_hi=ohi, out_lo=olo, out_scale=ACT, drop=drop)
    torch.cuda.synchronize()
    assert ops.conv3x3_tc_status() == 0
    y = (ohi.float() + olo.float()) / ACT
    return y.permute(0, 3, 1, 2).cpu()


def rel_err(a, ref):
    a, ref = a.double(), ref.double()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def cand_to_dict(c, device="cuda"):
    """oracle Candidates -> the dict layout of ops.decode_cov's output (B=1)."""
    M = c.boxes.shape[0]
    has_cov = isinstance(c.cov, torch.Tensor)
    return {
        "boxes": c.boxes.reshape(1, M, 4).contiguous().to(device),
        "cov": (c.cov if has_cov else torch.zeros((M, 4, 4))).reshape(1, M, 4, 4).contiguous().to(device),
        "scores": c.scores.reshape(1, M).contiguous().to(device),
        "classes": c.classes.to(torch.int32).reshape(1, M).contiguous().to(device),
        "probs": c.probs.reshape(1, M, -1).contiguous().to(device),
        "count": torch.tensor([M], dtype=torch.int32, device=device),
        "has_cov": has_cov,
    }
