"""Host-side orchestration of the B200 probabilistic-inference path (features -> detections).

Mirrors, for a BATCH of images, what the reference does for one image in
`RetinaNetProbabilisticPredictor` (reference src/probabilistic_inference/probabilistic_inference.py):
    head forward over N MC-dropout samples / E ensemble members   (:199-209, probabilistic_retinanet.py:104-108,517-523)
    Q1 sample means                                               (:214-270)
    scores, per-level top-k, threshold                            (:283-308)
    decode + aleatoric / epistemic covariance                     (:310-388)
    standard NMS or BayesOD fusion + rescale                      (:390-407, :536-636, inference_utils.py:374-425)
All arithmetic runs in libpodb200 (hand-written sm_100a CUDA); this file only sequences launches
and owns the (torch-allocated) device buffers.

Scheduling facts used: the first tower layer (conv+ReLU) does not depend on the dropout
mask, so it is evaluated once per image and tower and replicated into the N x passes masked copies;
in eval mode the second tower evaluation of the reference is bit-identical to the first and is
shared.  Samples and passes are folded into the GEMM M dimension (maps), so one launch per
(level, tower, layer) covers every image, sample and pass of the chunk.
"""
import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from ._cabi import POD_OUT_HIDDEN, POD_OUT_RAW, PodError

class nvtx_range:
    """NVTX range around a stage of the path (visible in nsys / ncu --nvtx timelines; a few hundred nanoseconds when no
    profiler is attached)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        torch.cuda.nvtx.range_pop()
        return False


ACT_SCALE = 16.0          # largest fp16 split scale of activations (|x| < 4094, abs. resolution 2^-28); the scale of a call is
                          # min(ACT_SCALE, largest power of two s with max|feature| * s <= FEATURE_TARGET), a device word
FEATURE_TARGET = 128.0    # = 8 * ACT_SCALE: stored features stay below 128, leaving 512x headroom (65504 / 128) for hidden
                          # activations larger than the features; beyond that pod_status reports saturation
TOWER_CLS, TOWER_BOX = 0, 1
Q1_GROUP = 8              # samples per partial sum of the fused Q1 accumulation: fixed, so the fp32 summation order (and
                          # with it every bit of the result) does not depend on the batch size or the GPU count


def _pad_cout(c):
    # <= 64 channels always pad to 64: those convolutions run weights-as-A, whose cost does not depend on Cout
    for p in (64, 80, 96, 128, 256):
        if c <= p:
            return p
    raise PodError("unsupported output channel count %d" % c)


@dataclass
class PackedConv:
    w_hi: torch.Tensor
    w_lo: torch.Tensor
    bias: torch.Tensor
    w_scale: float
    cout: int
    cout_pad: int
    col0: int = 0          # first output channel of this block (convs wider than 256 channels are split)
    total_cout: int = 0    # output channels of the whole convolution


def pack_conv(weight, bias, device, cout_pad=None):
    w = weight.detach().to(device=device, dtype=torch.float32).contiguous()
    cout = w.shape[0]
    cout_pad = cout_pad or _pad_cout(cout)
    scale = ops.pow2_scale(float(w.abs().max()), 1024.0)
    hi, lo = ops.pack_conv_weight(w, cout_pad, scale)
    b = torch.zeros((cout_pad,), dtype=torch.float32, device=device)
    b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
    return PackedConv(hi, lo, b, scale, cout, cout_pad, 0, cout)


def pack_conv_blocks(weight, bias, device, block=256):
    """Output convolution of any width (e.g. A*K = 720 channels for 80 classes) as column blocks of at most
    256 output channels, one tcgen05 launch each, all writing into the same permuted output rows."""
    cout = weight.shape[0]
    blocks = []
    for c0 in range(0, cout, block):
        pcv = pack_conv(weight[c0:c0 + block], bias[c0:c0 + block], device)
        pcv.col0, pcv.total_cout = c0, cout
        blocks.append(pcv)
    return blocks


class HeadWeights:
    """One weight set of the reference head (state-dict keys as in
    reference src/probabilistic_modeling/probabilistic_retinanet.py:401-484) packed for the
    tensor-core kernel."""

    def __init__(self, state_dict, use_dropout, cls_var, bbox_cov, num_convs=4, device="cuda"):
        step = 3 if use_dropout else 2
        sd = state_dict
        self.towers = [[], []]
        for i in range(num_convs):
            self.towers[TOWER_CLS].append(pack_conv(sd["head.cls_subnet.%d.weight" % (i * step)],
                                                    sd["head.cls_subnet.%d.bias" % (i * step)], device))
            self.towers[TOWER_BOX].append(pack_conv(sd["head.bbox_subnet.%d.weight" % (i * step)],
                                                    sd["head.bbox_subnet.%d.bias" % (i * step)], device))
        self.cls_score = pack_conv_blocks(sd["head.cls_score.weight"], sd["head.cls_score.bias"], device)
        self.bbox_pred = pack_conv_blocks(sd["head.bbox_pred.weight"], sd["head.bbox_pred.bias"], device)
        self.cls_var = pack_conv_blocks(sd["head.cls_var.weight"], sd["head.cls_var.bias"], device) if cls_var else None
        self.bbox_cov = pack_conv_blocks(sd["head.bbox_cov.weight"], sd["head.bbox_cov.bias"], device) if bbox_cov else None
        # eval mode: both heads of a tower read the same activations, so mean|variance weights are also packed
        # as ONE convolution (rows [mean; var]) whose epilogue routes the two column ranges to their buffers
        self.cls_fused = self.box_fused = None
        if cls_var and not use_dropout:
            self.cls_fused = pack_conv_blocks(torch.cat([sd["head.cls_score.weight"], sd["head.cls_var.weight"]], 0),
                                              torch.cat([sd["head.cls_score.bias"], sd["head.cls_var.bias"]], 0), device)
        if bbox_cov and not use_dropout:
            self.box_fused = pack_conv_blocks(torch.cat([sd["head.bbox_pred.weight"], sd["head.bbox_cov.weight"]], 0),
                                              torch.cat([sd["head.bbox_pred.bias"], sd["head.bbox_cov.bias"]], 0), device)


@dataclass
class PathConfig:
    """The values the predictor reads from cfg / the model object."""
    num_classes: int = 7
    num_anchors: int = 9
    dropout_rate: float = 0.0
    cls_var: bool = False
    bbox_cov: bool = False
    cov_dims: int = 4
    cls_var_num_samples: int = 10
    box_num_samples: int = 1000
    topk: int = 1000
    score_thresh: float = 0.05
    nms_thresh: float = 0.5
    max_dets: int = 100
    reg_weights: tuple = (1.0, 1.0, 1.0, 1.0)           # MODEL.RETINANET.BBOX_REG_WEIGHTS (apply_deltas)
    sample_reg_weights: tuple = None                    # MODEL.RPN.BBOX_REG_WEIGHTS (sampled decode); None = same
    affinity: float = 0.9
    box_merge: str = "bayesian_inference"
    cls_merge: str = "max_score"


class HeadEngine:
    """Runs the head + statistics + post-processing for a chunk of images on one GPU."""

    def __init__(self, pc: PathConfig, weight_sets: List[HeadWeights], device="cuda"):
        self.pc = pc
        self.ws = weight_sets
        self.device = torch.device(device)
        self._buf = {}
        self.profile_layers = False       # bench --profile-layers: tag tower launches by layer in ops.PROFILE
        # MC-dropout: apply the first dropout inside the layer-1 convolution (pod_conv_args.mask_in) instead of writing the
        # N x passes masked copies of the hoisted first layer with pod_mask_expand_split.  Bit-identical results, but
        # measured SLOWER (two mask warps cannot draw 8192 Philox decisions per K-block in the 1536 cycles its MMAs take:
        # layer 1 goes 164 -> 391 ms for the 16.9 ms of replication saved), so off by default.
        self.mask_in_kernel = False

    # ------------------------------------------------------------------ buffers (reused across calls)
    def _get(self, name, numel, dtype):
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty((numel,), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    # ------------------------------------------------------------------ head
    @staticmethod
    def feature_scale(feats):
        """fp16 split scale of EVERY activation of this call (features and tower layers): ACT_SCALE when
        max|feature| <= 8 (e.g. unit-variance maps), otherwise the largest smaller power of two that keeps the stored
        features below FEATURE_TARGET.  Values are computed in fp32 registers in true units and only the stored
        (hi, lo) pair is scaled, so results do not depend on the scale beyond fp16 round-off; a power-of-two scale
        makes the rescaling exact.  Computed by two small kernels into a device word that every kernel of the call
        reads: no torch ops, no host sync (non-finite inputs raise through ops.check_status at the end of the call)."""
        return ops.feature_scale_dev([HeadEngine._flat(f) for f in feats], ACT_SCALE, FEATURE_TARGET)

    @staticmethod
    def _is_channels_last(f):
        """(B, C, H, W)-shaped view of channels-last memory (what backbone_tc returns): no layout pass needed."""
        return f.dim() == 4 and not f.is_contiguous() and f.permute(0, 2, 3, 1).is_contiguous()

    @classmethod
    def _flat(cls, f):
        return f.permute(0, 2, 3, 1) if cls._is_channels_last(f) else f.contiguous()

    @classmethod
    def _split_input(cls, f, scale_dev):
        """FPN map -> channels-last fp16 split pair scaled by the call's device-resident scale."""
        if cls._is_channels_last(f):
            return ops.split_f32(f.permute(0, 2, 3, 1), scale_dev=scale_dev)
        return ops.nchw_to_nhwc_split_dev(f.contiguous(), scale_dev)

    def _conv_hidden(self, src, NB, H, W, pcv, dst, drop, scale_dev, map_group=0, map_live=0, tag=None):
        """256 -> 256 tower layer; `scale_dev` is the call's activation scale (input and output pairs)."""
        ops.conv3x3_tc(src[0], src[1], 1.0, NB, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, pcv.cout,
                       pcv.cout_pad, POD_OUT_HIDDEN, True, out_hi=dst[0], out_lo=dst[1], out_scale=1.0, drop=drop,
                       map_group=map_group, map_live=map_live, in_scale_dev=scale_dev, out_scale_dev=scale_dev, tag=tag)

    def _conv_out(self, src, NB, H, W, blocks, out, out_offset, out_map_stride, scale_dev, in_map_stride=None, in_offset=0,
                  map_group=0, map_live=0):
        for pcv in blocks:
            ops.conv3x3_tc(src[0], src[1], 1.0, NB, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias, pcv.cout,
                           pcv.cout_pad, POD_OUT_RAW, False, out_f32=out, out_offset=out_offset + pcv.col0,
                           out_map_stride=out_map_stride, out_pixel_stride=pcv.total_cout,
                           in_map_stride=in_map_stride, in_offset=in_offset, map_group=map_group, map_live=map_live,
                           in_scale_dev=scale_dev)

    def head_mc(self, feats, n_mc, seed, image0, skip_unread=False, fuse_q1=False):
        """MC-dropout head loop: feats = list over levels of (B,256,H,W) fp32.
        Returns raw per-sample outputs, each (B, N, R, D).

        fuse_q1 ("stream" | True, or "epilogue"; pre-NMS aggregation only, implies skip_unread): the reference's sample "mean" of box_cls, box_cls_var
        and box_reg_var (probabilistic_inference.py:214-270) is a fixed linear combination of the samples, and cls_score /
        cls_var / bbox_cov are linear (a 3x3 convolution plus bias; the weights sum to one), so
            mean_s head(x_s) = head(mean_s x_s).
        The mean activation (2 x_0 + x_1 + ... + x_{N-2}) / N of the last tower layer is formed either by one streaming
        pass over the per-sample maps ("stream", default) or inside the tcgen05 epilogue ("epilogue": nothing per
        sample is written, but the read-back of the running sums costs more than the streaming pass),
        and each of those three output convolutions runs ONCE per image: 3 + N output convolutions per location instead
        of 4N - 3, and no per-sample logits / variances are ever written (returned with a sample dimension of 1: they
        ARE the Q1 means).  Only box_delta stays per sample: every sample's decoded box enters the epistemic
        covariance (:326-331).  Differs from the unfused evaluation by fp32 round-off (~1e-7 relative).

        skip_unread: the reference's sample "mean" of box_cls / box_cls_var / box_reg_var runs over
        range(len-1) (probabilistic_inference.py:216-267), so those three outputs of the LAST sample are
        computed and never read; only its box_delta enters the epistemic covariance (:326-331).  With
        skip_unread the tower passes that feed only those outputs are not evaluated (both passes of the class
        tower and the variance pass of the box tower of sample N-1: 9 of the 12N+2 tower convolutions and 3 of
        the 4N output convolutions); rows [:, N-1] of logits / logvar / regvar are then left unwritten.  Valid
        only when the caller aggregates with the Q1 mean (pre-NMS modes), never for per-run inference."""
        pc = self.pc
        B = feats[0].shape[0]
        A, K = pc.num_anchors, pc.num_classes
        level_hw = [tuple(f.shape[-2:]) for f in feats]
        level_off = [0]
        for (h, wd) in level_hw:
            level_off.append(level_off[-1] + h * wd * A)
        R = level_off[-1]
        passes = 2 if (pc.cls_var or pc.bbox_cov) else 1
        dev = self.device
        fuse = bool(fuse_q1) and n_mc > 1
        in_epilogue = fuse and fuse_q1 == "epilogue"       # accumulate inside the tcgen05 epilogue (slower: DESIGN.md 3.7)
        if fuse:
            skip_unread = True
        n_stat = 1 if fuse else n_mc                     # sample dimension of the outputs that are only ever averaged
        raw = {"logits": torch.empty((B, n_stat, R, K), dtype=torch.float32, device=dev),
               "deltas": torch.empty((B, n_mc, R, 4), dtype=torch.float32, device=dev),
               "logvar": torch.empty((B, n_stat, R, K), dtype=torch.float32, device=dev) if pc.cls_var else None,
               "regvar": torch.empty((B, n_stat, R, pc.cov_dims), dtype=torch.float32, device=dev) if pc.bbox_cov else None}
        max_hw = max(h * wd for h, wd in level_hw)
        nmaps = B * n_mc * passes
        st = {"n_mc": n_mc, "seed": seed, "image0": image0, "skip_unread": skip_unread, "fuse": fuse, "in_epilogue": in_epilogue,
              "raw": raw, "level_off": level_off, "R": R, "B": B,
              "act": [(self._get("a%d_hi" % i, nmaps * max_hw * 256, torch.float16),
                       self._get("a%d_lo" % i, nmaps * max_hw * 256, torch.float16)) for i in range(2)],
              "c1": self._get("c1", B * max_hw * 256, torch.float32),
              "c1_pair": (self._get("c1_hi", B * max_hw * 256, torch.float16), self._get("c1_lo", B * max_hw * 256, torch.float16)),
              "groups": (n_mc + Q1_GROUP - 1) // Q1_GROUP, "q1_acc": None, "q1_mean": None}
        if fuse:
            if in_epilogue:
                st["q1_acc"] = self._get("q1_acc", B * 2 * st["groups"] * max_hw * 256, torch.float32)
            st["q1_mean"] = (self._get("q1m_hi", B * 2 * max_hw * 256, torch.float16),
                             self._get("q1m_lo", B * 2 * max_hw * 256, torch.float16))
        fscale = self.feature_scale(feats)
        for lvl, f in enumerate(feats):
            H, W = level_hw[lvl]
            fhi, flo = self._split_input(f, fscale)
            for tower in (TOWER_CLS, TOWER_BOX):
                with nvtx_range("pod.head_mc.P%d.%s" % (lvl + 3, "cls" if tower == TOWER_CLS else "box")):
                    self._head_mc_tower(st, lvl, tower, H, W, fhi, flo, fscale)
        return raw, level_off

    def _head_mc_tower(self, st, lvl, tower, H, W, fhi, flo, fscale):
        """One (level, tower) unit of head_mc: hoisted first layer, mask replication, masked tower layers, output heads."""
        pc, w = self.pc, self.ws[0]
        n_mc, seed, image0, B, R = st["n_mc"], st["seed"], st["image0"], st["B"], st["R"]
        skip_unread, fuse, in_epilogue = st["skip_unread"], st["fuse"], st["in_epilogue"]
        act, c1, raw, level_off, groups, q1_acc, q1_mean = (st[k] for k in ("act", "c1", "raw", "level_off", "groups", "q1_acc", "q1_mean"))
        A = pc.num_anchors
        HW = H * W
        tw = w.towers[tower]
        has_var = pc.cls_var if tower == TOWER_CLS else pc.bbox_cov
        t_passes = 2 if has_var else 1
        # layer 0: conv + ReLU once per image (Q2 hoist: it does not depend on the dropout masks)
        p0 = tw[0]
        d0 = ops.make_dropout(pc.dropout_rate, seed, image0, n_mc, t_passes, 0, tower, 0, lvl)
        # maps of one image: sample-major, pass-minor; the unread ones (skip_unread) are its tail
        grp = n_mc * t_passes
        live = grp
        if skip_unread and n_mc > 1:
            live = (n_mc - 1) * t_passes + (0 if tower == TOWER_CLS else 1)
        NB = B * n_mc * t_passes
        # in-kernel input masking needs a plain (non-accumulating) layer 1 on the CTA-pair row-halo kernel
        mask_in = self.mask_in_kernel and len(tw) >= 2 and not (in_epilogue and len(tw) == 2)
        if mask_in:
            # c1 * 1/(1-p) as ONE split-pair map per image; layer 1 applies every (sample, pass) mask to its staged tiles
            c1p = st["c1_pair"]
            ops.conv3x3_tc(fhi, flo, 1.0, B, H, W, 256, p0.w_hi, p0.w_lo, p0.w_scale, p0.bias, 256, 256,
                           POD_OUT_HIDDEN, True, out_hi=c1p[0], out_lo=c1p[1], out_scale=1.0, drop=d0, drop_scale_only=True,
                           in_scale_dev=fscale, out_scale_dev=fscale, tag="conv1")
        else:
            ops.conv3x3_tc(fhi, flo, 1.0, B, H, W, 256, p0.w_hi, p0.w_lo, p0.w_scale, p0.bias, 256, 256,
                           POD_OUT_RAW, True, out_f32=c1, out_map_stride=HW * 256, out_pixel_stride=256,
                           in_scale_dev=fscale)
            # the N x passes masked, rescaled, split copies of the first layer
            ops.mask_expand_split(c1[: B * HW * 256].view(B, HW, 256), d0, 1.0, act[0][0], act[0][1], live_reps=live,
                                  scale_dev=fscale)
        cur = 0
        # passes of this tower whose last layer is only ever averaged over the samples (fused Q1 mean)
        acc_mask = 0
        if fuse:
            acc_mask = ((1 << t_passes) - 1) if tower == TOWER_CLS else (2 if has_var else 0)
        n_acc = bin(acc_mask).count("1")
        q1_live = [n_mc - 1, n_mc - 1] if tower == TOWER_CLS else [n_mc, n_mc - 1]
        for layer in range(1, len(tw)):
            d = ops.make_dropout(pc.dropout_rate, seed, image0, n_mc, t_passes, 0, tower, layer, lvl)
            if acc_mask and layer == len(tw) - 1 and in_epilogue:
                ops.conv3x3_tc(act[cur][0], act[cur][1], 1.0, NB, H, W, 256, tw[layer].w_hi, tw[layer].w_lo,
                               tw[layer].w_scale, tw[layer].bias, 256, 256, POD_OUT_HIDDEN, True,
                               out_hi=act[cur ^ 1][0], out_lo=act[cur ^ 1][1], out_scale=1.0, drop=d,
                               in_scale_dev=fscale, out_scale_dev=fscale,
                               q1={"acc": q1_acc, "samples": n_mc, "passes": t_passes, "live": q1_live[:t_passes],
                                   "mask": acc_mask, "group": Q1_GROUP})
                ops.q1_finish(q1_acc, B * n_acc, groups, HW * 256, n_mc, fscale, q1_mean[0], q1_mean[1])
            elif mask_in and layer == 1:
                ops.conv3x3_tc(c1p[0], c1p[1], 1.0, NB, H, W, 256, tw[layer].w_hi, tw[layer].w_lo, tw[layer].w_scale,
                               tw[layer].bias, 256, 256, POD_OUT_HIDDEN, True, out_hi=act[cur ^ 1][0], out_lo=act[cur ^ 1][1],
                               out_scale=1.0, drop=d, map_group=grp, map_live=live, in_scale_dev=fscale, out_scale_dev=fscale,
                               mask_in=0, tag="tower256" if not self.profile_layers else "tower256_L1")
                if acc_mask and layer == len(tw) - 1:
                    ops.q1_mean_act(act[cur ^ 1][0], act[cur ^ 1][1], B, n_mc, t_passes, acc_mask, q1_live[:t_passes],
                                    HW * 256, fscale, q1_mean[0], q1_mean[1])
            else:
                self._conv_hidden(act[cur], NB, H, W, tw[layer], act[cur ^ 1], d, fscale, map_group=grp, map_live=live,
                                  tag="tower256" if not self.profile_layers else "tower256_L%d" % layer)
                if acc_mask and layer == len(tw) - 1:
                    # streaming form (default): the per-sample maps just written are averaged by one HBM-bound pass
                    ops.q1_mean_act(act[cur ^ 1][0], act[cur ^ 1][1], B, n_mc, t_passes, acc_mask, q1_live[:t_passes],
                                    HW * 256, fscale, q1_mean[0], q1_mean[1])
            cur ^= 1
        n_live = n_mc - 1 if (skip_unread and n_mc > 1) else n_mc        # samples whose mean/var heads are read
        # output convs: pass-0 maps feed the mean head, pass-1 maps the variance head (Q2)
        mean_pc, var_pc = (w.cls_score, w.cls_var) if tower == TOWER_CLS else (w.bbox_pred, w.bbox_cov)
        mean_out = raw["logits"] if tower == TOWER_CLS else raw["deltas"]
        D = mean_pc[0].total_cout // A
        if acc_mask:
            # the averaged outputs: ONE convolution per image on the mean activation (accumulated pass a of image b
            # is map b * n_acc + a of q1_mean)
            a_idx = 0
            if tower == TOWER_CLS:
                self._conv_out(q1_mean, B, H, W, mean_pc, mean_out, level_off[lvl] * D, R * D, fscale,
                               in_map_stride=n_acc * HW * 256, in_offset=0)
                a_idx = 1
            else:
                self._conv_out(act[cur], B * n_mc, H, W, mean_pc, mean_out, level_off[lvl] * D, R * D, fscale,
                               in_map_stride=t_passes * HW * 256, in_offset=0, map_group=n_mc, map_live=n_mc)
            if has_var:
                var_out = raw["logvar"] if tower == TOWER_CLS else raw["regvar"]
                Dv = var_pc[0].total_cout // A
                self._conv_out(q1_mean, B, H, W, var_pc, var_out, level_off[lvl] * Dv, R * Dv, fscale,
                               in_map_stride=n_acc * HW * 256, in_offset=a_idx * HW * 256)
            return
        self._conv_out(act[cur], B * n_mc, H, W, mean_pc, mean_out, level_off[lvl] * D, R * D, fscale,
                       in_map_stride=t_passes * HW * 256, in_offset=0, map_group=n_mc,
                       map_live=n_live if tower == TOWER_CLS else n_mc)
        if has_var:
            var_out = raw["logvar"] if tower == TOWER_CLS else raw["regvar"]
            Dv = var_pc[0].total_cout // A
            self._conv_out(act[cur], B * n_mc, H, W, var_pc, var_out, level_off[lvl] * Dv, R * Dv, fscale,
                           in_map_stride=2 * HW * 256, in_offset=HW * 256, map_group=n_mc, map_live=n_live)

    def head_eval(self, feats, members=None, skip_unread=False, per_member_feats=False):
        """Deterministic head (eval mode): one forward per weight set; the reference's second tower
        evaluation is identical to the first and is shared.  Returns (B, E, R, D) raw outputs.
        per_member_feats: feats[e][l] -- every ensemble member reads its OWN feature maps (each member of the
        reference is a full model with its own backbone, probabilistic_inference.py:58-77,499-501); otherwise one
        list over levels shared by all members.
        skip_unread (see head_mc): the class tower of the LAST member feeds only outputs the Q1 mean never reads
        and is not evaluated; rows [:, E-1] of logits / logvar stay unwritten (regvar is written: it shares the
        fused box convolution with the deltas)."""
        pc = self.pc
        members = members if members is not None else list(range(len(self.ws)))
        E = len(members)
        feat_sets = list(feats) if per_member_feats else [feats]
        if per_member_feats and len(feat_sets) != E:
            raise PodError("head_eval: %d feature sets for %d members" % (len(feat_sets), E))
        feats = feat_sets[0]
        B = feats[0].shape[0]
        A, K = pc.num_anchors, pc.num_classes
        level_hw = [tuple(f.shape[-2:]) for f in feats]
        for fs in feat_sets[1:]:
            if [tuple(f.shape) for f in fs] != [tuple(f.shape) for f in feats]:
                raise PodError("head_eval: every member's feature maps must have the same shapes")
        level_off = [0]
        for (h, wd) in level_hw:
            level_off.append(level_off[-1] + h * wd * A)
        R = level_off[-1]
        dev = self.device
        raw = {"logits": torch.empty((B, E, R, K), dtype=torch.float32, device=dev),
               "deltas": torch.empty((B, E, R, 4), dtype=torch.float32, device=dev),
               "logvar": torch.empty((B, E, R, K), dtype=torch.float32, device=dev) if pc.cls_var else None,
               "regvar": torch.empty((B, E, R, pc.cov_dims), dtype=torch.float32, device=dev) if pc.bbox_cov else None}
        max_hw = max(h * wd for h, wd in level_hw)
        act = [(self._get("a%d_hi" % i, B * max_hw * 256, torch.float16),
                self._get("a%d_lo" % i, B * max_hw * 256, torch.float16)) for i in range(2)]
        fscales = [self.feature_scale(fs) for fs in feat_sets]
        for lvl in range(len(feats)):
            H, W = level_hw[lvl]
            split_maps = [self._split_input(fs[lvl], sc) for fs, sc in zip(feat_sets, fscales)]
            for e, mi in enumerate(members):
                w = self.ws[mi]
                (fhi, flo), fscale = split_maps[e if per_member_feats else 0], fscales[e if per_member_feats else 0]
                for tower in (TOWER_CLS, TOWER_BOX):
                    if skip_unread and E > 1 and e == E - 1 and tower == TOWER_CLS:
                        continue
                    torch.cuda.nvtx.range_push("pod.head_eval.P%d.m%d.%s" % (lvl + 3, e, "cls" if tower == TOWER_CLS else "box"))
                    tw = w.towers[tower]
                    src, cur = (fhi, flo), 0
                    for layer in range(len(tw)):
                        self._conv_hidden(src, B, H, W, tw[layer], act[cur], None, fscale)
                        src = act[cur]
                        cur ^= 1
                    mean_pc, var_pc = (w.cls_score, w.cls_var) if tower == TOWER_CLS else (w.bbox_pred, w.bbox_cov)
                    mean_out = raw["logits"] if tower == TOWER_CLS else raw["deltas"]
                    D = mean_pc[0].total_cout // A
                    fused = w.cls_fused if tower == TOWER_CLS else w.box_fused
                    if fused is not None:
                        var_out = raw["logvar"] if tower == TOWER_CLS else raw["regvar"]
                        Dv = var_pc[0].total_cout // A
                        split = mean_pc[0].total_cout
                        for pcv in fused:
                            ops.conv3x3_tc(src[0], src[1], 1.0, B, H, W, 256, pcv.w_hi, pcv.w_lo, pcv.w_scale, pcv.bias,
                                           pcv.cout, pcv.cout_pad, POD_OUT_RAW, False, in_scale_dev=fscale,
                                           out_f32=mean_out, out_offset=(e * R + level_off[lvl]) * D + pcv.col0,
                                           out_map_stride=E * R * D, out_pixel_stride=A * D,
                                           out2_f32=var_out, out2_offset=(e * R + level_off[lvl]) * Dv,
                                           split_col=split - pcv.col0, out2_map_stride=E * R * Dv, out2_pixel_stride=A * Dv)
                        torch.cuda.nvtx.range_pop()
                        continue
                    self._conv_out(src, B, H, W, mean_pc, mean_out, (e * R + level_off[lvl]) * D, E * R * D, fscale)
                    if var_pc is not None:
                        var_out = raw["logvar"] if tower == TOWER_CLS else raw["regvar"]
                        Dv = var_pc[0].total_cout // A
                        self._conv_out(src, B, H, W, var_pc, var_out, (e * R + level_off[lvl]) * Dv, E * R * Dv, fscale)
                    torch.cuda.nvtx.range_pop()
        return raw, level_off

    # ------------------------------------------------------------------ statistics + post-processing
    def candidates(self, raw, level_off, anchors, seed, image0, runs=1):
        """Q1 means -> scores -> top-k -> decode + covariance.  With runs > 1 every (image, sample) row of
        `raw` is treated as its own single-sample inference (post-NMS merge modes)."""
        pc = self.pc
        if runs > 1:
            raw = {k: (v.reshape((v.shape[0] * v.shape[1], 1) + tuple(v.shape[2:])) if v is not None else None)
                   for k, v in raw.items()}
        S = raw["deltas"].shape[1]
        if S > 1:
            # outputs that arrive with a sample dimension of 1 already ARE the Q1 means (head_mc fuse_q1)
            q1 = lambda t: None if t is None else (ops.sample_mean_q1(t) if t.shape[1] > 1 else
                                                   (t[:, 0] if t.shape[0] == 1 else t[:, 0].contiguous()))
            m_logits, m_deltas = q1(raw["logits"]), ops.sample_mean_q1(raw["deltas"])
            m_logvar, m_regvar = q1(raw["logvar"]), q1(raw["regvar"])
        else:
            m_logits, m_deltas = raw["logits"][:, 0], raw["deltas"][:, 0]
            m_logvar = raw["logvar"][:, 0] if raw["logvar"] is not None else None
            m_regvar = raw["regvar"][:, 0] if raw["regvar"] is not None else None
            if raw["logits"].shape[0] > 1:
                m_logits, m_deltas = m_logits.contiguous(), m_deltas.contiguous()
                m_logvar = m_logvar.contiguous() if m_logvar is not None else None
                m_regvar = m_regvar.contiguous() if m_regvar is not None else None
        with nvtx_range("pod.scores_topk"):
            probs, score, cls = ops.scores(m_logits, m_logvar, level_off, pc.cls_var_num_samples, seed, image0, runs=runs)
            cand_idx, cand_cnt, seg = ops.topk_levels(score, level_off, pc.topk, pc.score_thresh)
        with nvtx_range("pod.decode_cov"):
            cand = ops.decode_cov(m_deltas, m_regvar, raw["deltas"] if S > 1 else None, anchors, probs, score, cls,
                                  cand_idx, cand_cnt, seg, pc.box_num_samples, seed, image0, pc.reg_weights, runs=runs,
                                  sample_reg_weights=pc.sample_reg_weights)
        return cand

    def detections(self, cand, fuse_mode, image_hw, out_hw, nms_variant=ops.NMS_AUTO, skip_post=False):
        """fuse_mode: 0 standard NMS, 1 BayesOD, 2 anchor statistics."""
        pc = self.pc
        with nvtx_range("pod.nms_fuse"):
            return ops.nms_fuse(cand, int(fuse_mode), pc.nms_thresh, pc.affinity, pc.max_dets, image_hw, out_hw,
                                nms_variant=nms_variant, skip_post=skip_post,
                                box_merge=0 if pc.box_merge == "bayesian_inference" else 1,
                                cls_merge=0 if pc.cls_merge == "max_score" else 1)


    def merged_detections(self, raw, level_off, anchors, seed, image0, image_hw, out_hw):
        """Post-NMS merging (reference probabilistic_inference.py:444-481,506-534 and
        inference_utils.py:165-289): per-run inference + NMS, sequential clustering, final NMS + rescale."""
        runs = raw["logits"].shape[1]
        cand = self.candidates(raw, level_off, anchors, seed, image0, runs=runs)
        per_run = self.detections(cand, 0, image_hw, image_hw, skip_post=True)
        clusters = ops.cluster_merge(per_run, runs, self.pc.affinity)
        return cand, per_run, clusters, self.detections(clusters, 0, image_hw, out_hw)


def make_anchors(level_hw, sizes, aspect_ratios, strides, offset=0.0, device="cuda"):
    """detectron2 DefaultAnchorGenerator semantics (un-vendored dependency, call site reference
    probabilistic_retinanet.py:101): cell anchors for size outer / ratio inner, w = sqrt(s^2/r),
    h = r*w, centred; grid shifts row-major, anchor-minor.  Returns (R,4) fp32 on `device`."""
    out = []
    for (gh, gw), stride, sz in zip(level_hw, strides, sizes):
        rows = []
        for s in sz:
            area = float(s) ** 2.0
            for r in aspect_ratios:
                w = math.sqrt(area / r)
                h = r * w
                rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        base = torch.tensor(rows, dtype=torch.float32)
        sx = torch.arange(offset * stride, gw * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(offset * stride, gh * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return torch.cat(out).to(device).contiguous()
