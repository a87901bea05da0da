"""ResNet-50 + FPN (P3-P7) feature extractor feeding the probabilistic head path.

The backbone is UPSTREAM of the path rebuilt in this repository: in the
reference it is detectron2's `build_retinanet_resnet_fpn_backbone` (un-vendored dependency; call sites
reference src/probabilistic_modeling/probabilistic_retinanet.py:96-101).  This module restates that
architecture with stock torch ops in fp32 (library convolutions, no hand-written kernels) so that
`predictor(input_im)` works from raw images and detectron2 checkpoints load by key name:

    backbone.bottom_up.stem.conv1.{weight, norm.*}          7x7/2 conv + FrozenBN + ReLU + 3x3/2 max-pool
    backbone.bottom_up.res{2,3,4,5}.<i>.{conv1,conv2,conv3,shortcut}.{weight, norm.*}   bottleneck blocks (3,4,6,3)
    backbone.fpn_lateral{3,4,5}.{weight,bias}, backbone.fpn_output{3,4,5}.{weight,bias}  FPN, nearest top-down
    backbone.top_block.p6/p7.{weight,bias}                  LastLevelP6P7 on res5 (detectron2 v0.2-v0.3)

Preprocessing follows detectron2's `preprocess_image`: (image - PIXEL_MEAN) / PIXEL_STD in the input
channel order (BGR for the MSRA weights), zero-padded (bottom / right) to a multiple of the backbone's
`size_divisibility`.  For detectron2's FPN that is the stride of the LAST BOTTOM-UP level fed to the FPN
(res5: 32), not of the top block's P7: a 1280x720 frame becomes 1280x736 (P3 92x160, P4 46x80, P5 23x40, and the
stride-2 3x3 convolutions give P6 12x20, P7 6x10).  SURVEY 8(d) / BASELINE.md quote the benchmark geometry on a
128-padded frame (768x1280, P3 96x160); `synthetic.level_shapes` keeps that for the features-in benchmark,
the head path itself is shape-generic.
"""
import torch
import torch.nn.functional as F

STAGE_BLOCKS = {"res2": 3, "res3": 4, "res4": 6, "res5": 3}
STAGE_CH = {"res2": (64, 256), "res3": (128, 512), "res4": (256, 1024), "res5": (512, 2048)}


def _conv_bn_names(prefix):
    return [prefix + ".weight", prefix + ".norm.weight", prefix + ".norm.bias", prefix + ".norm.running_mean",
            prefix + ".norm.running_var"]


def expected_keys():
    keys = _conv_bn_names("backbone.bottom_up.stem.conv1")
    for st, n in STAGE_BLOCKS.items():
        for i in range(n):
            p = "backbone.bottom_up.%s.%d" % (st, i)
            for c in ("conv1", "conv2", "conv3"):
                keys += _conv_bn_names(p + "." + c)
            if i == 0:
                keys += _conv_bn_names(p + ".shortcut")
    for l in (3, 4, 5):
        keys += ["backbone.fpn_lateral%d.weight" % l, "backbone.fpn_lateral%d.bias" % l,
                 "backbone.fpn_output%d.weight" % l, "backbone.fpn_output%d.bias" % l]
    keys += ["backbone.top_block.p6.weight", "backbone.top_block.p6.bias", "backbone.top_block.p7.weight",
             "backbone.top_block.p7.bias"]
    return keys


def random_state_dict(seed=0, out_channels=256):
    """Random-init weights with the detectron2 key names (synthetic benchmarks / shape tests)."""
    g = torch.Generator().manual_seed(9001 + seed)
    sd = {}

    def conv(name, cout, cin, k, bn=True):
        fan = cin * k * k
        sd[name + ".weight"] = torch.randn((cout, cin, k, k), generator=g) * (2.0 / fan) ** 0.5
        if name.endswith("stem.conv1"):
            sd[name + ".weight"] /= 64.0            # pixel values are O(100): bring the stem output to O(1)
        if bn:
            # residual branches start small (as zero-gamma initialisation does) so that 16 stacked blocks keep O(1) maps
            sd[name + ".norm.weight"] = torch.full((cout,), 0.25) if name.endswith("conv3") else torch.ones(cout)
            sd[name + ".norm.bias"] = torch.zeros(cout)
            sd[name + ".norm.running_mean"] = torch.zeros(cout)
            sd[name + ".norm.running_var"] = torch.ones(cout)
        else:
            sd[name + ".bias"] = torch.zeros(cout)

    conv("backbone.bottom_up.stem.conv1", 64, 3, 7)
    cin = 64
    for st, n in STAGE_BLOCKS.items():
        mid, cout = STAGE_CH[st]
        for i in range(n):
            p = "backbone.bottom_up.%s.%d" % (st, i)
            conv(p + ".conv1", mid, cin, 1)
            conv(p + ".conv2", mid, mid, 3)
            conv(p + ".conv3", cout, mid, 1)
            if i == 0:
                conv(p + ".shortcut", cout, cin, 1)
            cin = cout
    for l, c in ((3, 512), (4, 1024), (5, 2048)):
        conv("backbone.fpn_lateral%d" % l, out_channels, c, 1, bn=False)
        conv("backbone.fpn_output%d" % l, out_channels, out_channels, 3, bn=False)
    conv("backbone.top_block.p6", out_channels, 2048, 3, bn=False)
    conv("backbone.top_block.p7", out_channels, out_channels, 3, bn=False)
    return sd


class ResNetFPNBackbone:
    """Callable: list of (3,H,W) images (uint8 or float, one size) -> [P3..P7], each (B,256,Hl,Wl) fp32."""

    def __init__(self, state_dict, pixel_mean=(103.530, 116.280, 123.675), pixel_std=(1.0, 1.0, 1.0), device="cuda",
                 stride_in_1x1=True, size_divisibility=32, eps=1e-5):
        self.device = torch.device(device)
        self.stride_in_1x1 = stride_in_1x1
        self.div = size_divisibility
        self.mean = torch.tensor(pixel_mean, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        self.std = torch.tensor(pixel_std, dtype=torch.float32, device=self.device).view(1, 3, 1, 1)
        missing = [k for k in expected_keys() if k not in state_dict]
        if missing:
            raise KeyError("backbone state dict lacks %d keys, e.g. %s" % (len(missing), missing[:3]))
        self.w = {}
        sd = state_dict
        # FrozenBatchNorm folded into a per-channel scale / bias applied after the convolution
        for k in expected_keys():
            if k.endswith(".norm.weight"):
                p = k[: -len(".norm.weight")]
                scale = sd[p + ".norm.weight"].float() * (sd[p + ".norm.running_var"].float() + eps).rsqrt()
                bias = sd[p + ".norm.bias"].float() - sd[p + ".norm.running_mean"].float() * scale
                self.w[p] = (sd[p + ".weight"].float().to(self.device), scale.view(1, -1, 1, 1).to(self.device),
                             bias.view(1, -1, 1, 1).to(self.device))
        for l in (3, 4, 5):
            for n in ("fpn_lateral%d" % l, "fpn_output%d" % l):
                self.w["backbone." + n] = (sd["backbone.%s.weight" % n].float().to(self.device),
                                          sd["backbone.%s.bias" % n].float().to(self.device))
        for n in ("p6", "p7"):
            self.w["backbone.top_block." + n] = (sd["backbone.top_block.%s.weight" % n].float().to(self.device),
                                                sd["backbone.top_block.%s.bias" % n].float().to(self.device))

    def _cbn(self, x, name, stride=1, padding=0, relu=True):
        w, s, b = self.w[name]
        y = F.conv2d(x, w, None, stride=stride, padding=padding) * s + b
        return F.relu_(y) if relu else y

    def _bottleneck(self, x, p, stride, has_shortcut):
        s1, s3 = (stride, 1) if self.stride_in_1x1 else (1, stride)
        out = self._cbn(x, p + ".conv1", stride=s1)
        out = self._cbn(out, p + ".conv2", stride=s3, padding=1)
        out = self._cbn(out, p + ".conv3", relu=False)
        sc = self._cbn(x, p + ".shortcut", stride=stride, relu=False) if has_shortcut else x
        return F.relu_(out + sc)

    def preprocess(self, images):
        x = torch.stack([im.to(self.device, dtype=torch.float32) for im in images])
        x = (x - self.mean) / self.std
        H, W = x.shape[-2:]
        ph, pw = (H + self.div - 1) // self.div * self.div, (W + self.div - 1) // self.div * self.div
        return F.pad(x, (0, pw - W, 0, ph - H))

    @torch.no_grad()
    def __call__(self, images):
        with torch.backends.cudnn.flags(allow_tf32=False), torch.no_grad():
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False
            try:
                x = self.preprocess(images)
                x = self._cbn(x, "backbone.bottom_up.stem.conv1", stride=2, padding=3)
                x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
                feats = {}
                for st, n in STAGE_BLOCKS.items():
                    for i in range(n):
                        stride = 2 if (i == 0 and st != "res2") else 1
                        x = self._bottleneck(x, "backbone.bottom_up.%s.%d" % (st, i), stride, i == 0)
                    feats[st] = x
                lat = {l: F.conv2d(feats["res%d" % l], *self.w["backbone.fpn_lateral%d" % l]) for l in (3, 4, 5)}
                p5 = lat[5]
                p4 = lat[4] + F.interpolate(p5, scale_factor=2.0, mode="nearest")
                p3 = lat[3] + F.interpolate(p4, scale_factor=2.0, mode="nearest")
                outs = [F.conv2d(p, *self.w["backbone.fpn_output%d" % l], padding=1) for p, l in ((p3, 3), (p4, 4), (p5, 5))]
                p6 = F.conv2d(feats["res5"], *self.w["backbone.top_block.p6"], stride=2, padding=1)
                p7 = F.conv2d(F.relu(p6), *self.w["backbone.top_block.p7"], stride=2, padding=1)
                return outs + [p6, p7]
            finally:
                torch.backends.cuda.matmul.allow_tf32 = prev
