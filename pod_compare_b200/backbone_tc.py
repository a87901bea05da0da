"""ResNet-50 + FPN (P3-P7) on B200 with this repository's own kernels.

The reference calls detectron2's `build_retinanet_resnet_fpn_backbone` (un-vendored dependency; call sites reference
src/probabilistic_modeling/probabilistic_retinanet.py:96-101).  `backbone.py` restates that architecture with torch's
library convolutions and is the ORACLE of this module (tests compare the two at 1e-4 on the same weights); here every
layer runs hand-written sm_100a code:

    preprocess + stem 7x7/2 + FrozenBN + ReLU + max-pool 3x3/2     csrc/backbone.cu   (SIMT fp32: K = 147, 2 % of the FLOPs)
    16 bottleneck blocks: 1x1 / 3x3 / 1x1 (+ 1x1 shortcut),         csrc/conv_tc.cu    pod_conv_tc_general: TMA + tcgen05
      stride 2 on the first 1x1 of res3-5 (STRIDE_IN_1X1),                             implicit GEMM, fp16x3 split operands,
      residual add + ReLU in the epilogue                                              fp32 TMEM accumulation
    FPN lateral 1x1, top-down nearest x2 + add, output 3x3          conv_tc.cu + backbone.cu:k_upsample2_add
    P6 = 3x3/2 on res5, P7 = 3x3/2 on relu(P6)                      conv_tc.cu (TMA element strides do the striding)

Layout: channels-last everywhere; activations are fp16 split pairs (x * ACT ~= hi + lo), FrozenBatchNorm is folded
into the packed weights (per-channel scale) and the epilogue bias.  The five FPN maps come back as fp32 channels-last
memory viewed as (B, 256, H, W), so `infer_from_features` takes them without a layout pass (engine.HeadEngine reads
channels-last inputs directly) and without a host round trip.
"""
import torch

from . import ops
from .backbone import STAGE_BLOCKS, STAGE_CH, expected_keys

ACT = 4.0                 # split scale of backbone activations: |x| < 16376, absolute resolution 2^-26


def _block_cols(cout):
    return 64 if cout <= 64 else (128 if cout <= 128 else 256)


class _Conv:
    """One convolution with folded FrozenBN, packed for pod_conv_tc_general."""

    def __init__(self, w, scale, bias, device):
        w = w.float()
        if scale is not None:
            w = w * scale.float().view(-1, 1, 1, 1)
        w = w.to(device).contiguous()
        self.cout, self.cin, self.k = int(w.shape[0]), int(w.shape[1]), int(w.shape[2])
        self.block = _block_cols(self.cout)
        self.rows = (self.cout + self.block - 1) // self.block * self.block
        self.w_scale = ops.pow2_scale(float(w.abs().max()), 1024.0)
        self.w_hi, self.w_lo = ops.pack_conv_weight_k(w, self.rows, self.w_scale)
        self.bias = torch.zeros((self.rows,), dtype=torch.float32, device=device)
        self.bias[: self.cout] = bias.float().to(device)


class TcResNetFPNBackbone:
    """Callable: list of (3,H,W) images (uint8 or float, one size) -> [P3..P7], each a (B,256,Hl,Wl) fp32 view of
    channels-last memory.  Same constructor arguments and key names as backbone.ResNetFPNBackbone."""

    def __init__(self, state_dict, pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), device="cuda",
                 stride_in_1x1=True, size_divisibility=32, eps=1e-5):
        self.device = torch.device(device)
        self.stride_in_1x1 = stride_in_1x1
        self.div = size_divisibility
        self.mean, self.std = [float(v) for v in pixel_mean], [float(v) for v in pixel_std]
        sd = state_dict
        missing = [k for k in expected_keys() if k not in sd]
        if missing:
            raise KeyError("backbone state dict lacks %d keys, e.g. %s" % (len(missing), missing[:3]))

        def bn(p):
            s = sd[p + ".norm.weight"].float() * (sd[p + ".norm.running_var"].float() + eps).rsqrt()
            return s, sd[p + ".norm.bias"].float() - sd[p + ".norm.running_mean"].float() * s

        s, b = bn("backbone.bottom_up.stem.conv1")
        w = sd["backbone.bottom_up.stem.conv1.weight"].float() * s.view(-1, 1, 1, 1)          # (64, 3, 7, 7)
        self.stem_w = w.permute(2, 3, 1, 0).reshape(147, 64).contiguous().to(self.device)       # [(ky*7+kx)*3+c][64]
        self.stem_b = b.contiguous().to(self.device)
        self.convs = {}
        for st, n in STAGE_BLOCKS.items():
            for i in range(n):
                p = "backbone.bottom_up.%s.%d" % (st, i)
                for c in ("conv1", "conv2", "conv3") + (("shortcut",) if i == 0 else ()):
                    s, b = bn(p + "." + c)
                    self.convs[p + "." + c] = _Conv(sd[p + "." + c + ".weight"], s, b, self.device)
        for name in ["backbone.fpn_lateral%d" % l for l in (3, 4, 5)] + ["backbone.fpn_output%d" % l for l in (3, 4, 5)] + \
                    ["backbone.top_block.p6", "backbone.top_block.p7"]:
            self.convs[name] = _Conv(sd[name + ".weight"], None, sd[name + ".bias"], self.device)
        self._buf = {}

    # ------------------------------------------------------------------------------------------
    def _get(self, name, shape, dtype):
        n = 1
        for v in shape:
            n *= int(v)
        t = self._buf.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty((n,), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t[:n].view(*shape)

    def _pair(self, name, shape):
        return self._get(name + "_hi", shape, torch.float16), self._get(name + "_lo", shape, torch.float16)

    def _conv(self, name, x, NB, H, W, stride, relu, out_name=None, res=None, out_f32=False, in_scale=ACT):
        cv = self.convs[name]
        pad = cv.k // 2
        Ho, Wo = (H + 2 * pad - cv.k) // stride + 1, (W + 2 * pad - cv.k) // stride + 1
        if out_f32:
            out = torch.empty((NB, Ho, Wo, cv.cout), dtype=torch.float32, device=self.device)
            ops.conv_tc_general(x[0], x[1], in_scale, NB, H, W, cv.cin, cv.k, stride, cv.w_hi, cv.w_lo, cv.w_scale, cv.rows, cv.cout,
                                cv.bias, relu, cv.block, out_ch_stride=cv.cout, out_f32=out)
            return out, Ho, Wo
        out = self._pair(out_name, (NB, Ho, Wo, cv.rows))
        ops.conv_tc_general(x[0], x[1], in_scale, NB, H, W, cv.cin, cv.k, stride, cv.w_hi, cv.w_lo, cv.w_scale, cv.rows, cv.cout,
                            cv.bias, relu, cv.block, out_hi=out[0], out_lo=out[1], out_scale=ACT, out_ch_stride=cv.rows,
                            res=res, res_scale=ACT)
        return out, Ho, Wo

    def padded_hw(self, H, W):
        return (H + self.div - 1) // self.div * self.div, (W + self.div - 1) // self.div * self.div

    @torch.no_grad()
    def __call__(self, images):
        torch.cuda.nvtx.range_push("pod.backbone")
        try:
            return self._forward(images)
        finally:
            torch.cuda.nvtx.range_pop()

    def _forward(self, images):
        if isinstance(images, (list, tuple)):
            x = torch.stack([im.to(self.device) for im in images])
        else:
            x = images.to(self.device)
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        x = x.contiguous()
        NB, _, Himg, Wimg = x.shape
        H, W = self.padded_hw(Himg, Wimg)
        Hc, Wc = (H + 1) // 2, (W + 1) // 2
        Hp, Wp = (Hc + 1) // 2, (Wc + 1) // 2
        scratch = self._get("stem_scratch", (NB, Hc, Wc, 64), torch.float32)
        cur = self._pair("stem", (NB, Hp, Wp, 64))
        ops.stem_conv7_pool(x, H, W, self.mean, self.std, self.stem_w, self.stem_b, scratch, cur[0], cur[1], ACT)
        h, w = Hp, Wp
        feats = {}
        flip = 0
        for st, n in STAGE_BLOCKS.items():
            for i in range(n):
                p = "backbone.bottom_up.%s.%d" % (st, i)
                stride = 2 if (i == 0 and st != "res2") else 1
                s1, s3 = (stride, 1) if self.stride_in_1x1 else (1, stride)
                a, h1, w1 = self._conv(p + ".conv1", cur, NB, h, w, s1, True, "mid_a")
                b, h2, w2 = self._conv(p + ".conv2", a, NB, h1, w1, s3, True, "mid_b")
                if i == 0:
                    sc, _, _ = self._conv(p + ".shortcut", cur, NB, h, w, stride, False, "shortcut")
                else:
                    sc = cur
                flip ^= 1
                cur, h, w = self._conv(p + ".conv3", b, NB, h2, w2, 1, True, "blk%d" % flip, res=sc)
            # the stage output feeds the FPN after later stages overwrote the ping-pong buffers: keep a copy of the pair
            keep = self._pair("keep_" + st, tuple(cur[0].shape))
            keep[0].copy_(cur[0]); keep[1].copy_(cur[1])
            feats[st] = (keep, h, w)
        # FPN (detectron2 FPN.forward): lateral 1x1 -> top-down nearest x2 + add -> output 3x3
        lat = {}
        for l in (5, 4, 3):
            src, hh, ww = feats["res%d" % l]
            lat[l], _, _ = self._conv("backbone.fpn_lateral%d" % l, src, NB, hh, ww, 1, False, out_f32=True)
            if l < 5:
                ops.upsample2_add(lat[l], lat[l + 1])
        outs = []
        for l in (3, 4, 5):
            src, hh, ww = feats["res%d" % l]
            sp = ops.split_f32(lat[l], scale=ACT, out_hi=self._get("lat_hi", tuple(lat[l].shape), torch.float16),
                               out_lo=self._get("lat_lo", tuple(lat[l].shape), torch.float16))
            o, _, _ = self._conv("backbone.fpn_output%d" % l, sp, NB, hh, ww, 1, False, out_f32=True)
            outs.append(o)
        # LastLevelP6P7 on res5 (detectron2 v0.2-v0.3): p6 = conv3x3/2(res5), p7 = conv3x3/2(relu(p6))
        src, hh, ww = feats["res5"]
        p6, h6, w6 = self._conv("backbone.top_block.p6", src, NB, hh, ww, 2, False, out_f32=True)
        sp6 = ops.split_f32(p6, scale=ACT, relu=True)
        p7, _, _ = self._conv("backbone.top_block.p7", sp6, NB, h6, w6, 2, False, out_f32=True)
        outs += [p6, p7]
        # (B, H, W, 256) channels-last memory presented with the reference's (B, 256, H, W) shape
        return [o.permute(0, 3, 1, 2) for o in outs]
