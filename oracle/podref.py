"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's probabilistic-inference path.

This module is the checker, never the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  It restates, in fp32 torch
CPU ops, what the reference computes between the FPN features and the final Instances, so
that it can travel to the GPU box where /root/reference does not exist.

PARITY PIN: the reference ships no tests or golden vectors, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py runs the
unmodified reference modules (through oracle/ref_runner.py) on seeded synthetic inputs and
commits the results under tests/golden/; tests/test_oracle_golden.py requires this file to
reproduce them, and tests/test_oracle_live_reference.py re-runs the live comparison whenever
/root/reference is present.

Reference lines followed (relative to /root/reference/src):
  head forward ............ probabilistic_modeling/probabilistic_retinanet.py:401-441,458-484,512-523
  (N,AK,H,W)->(N,HWA,K) ... probabilistic_retinanet.py:343-349 (detectron2 permute_to_N_HWA_K)
  MC list replication ..... probabilistic_retinanet.py:104-108 ; split :207-209 of probabilistic_inference.py
  quirk mean (Q1) ......... probabilistic_inference/probabilistic_inference.py:214-270
  score path .............. probabilistic_inference.py:283-308
  cholesky ................ probabilistic_modeling/modeling_utils.py:4-22
  epistemic covariance .... probabilistic_inference.py:322-331 ; inference_utils.py:337-371
  aleatoric Monte-Carlo ... probabilistic_inference.py:344-385 ; inference_utils.py:510-547
  standard NMS ............ probabilistic_inference/inference_utils.py:12-54
  post-NMS merging ........ probabilistic_inference.py:444-481,506-534 ; inference_utils.py:165-289
  BayesOD ................. probabilistic_inference.py:536-636 ; inference_utils.py:292-334
  anchor statistics ....... probabilistic_inference.py:409-428 ; inference_utils.py:57-162
  rescale / clip / cov .... inference_utils.py:374-425
  xyxy->xywh, JSON ........ inference_utils.py:428-502
Third-party arithmetic restated from its published semantics (un-vendored, un-pinned
detectron2 of the v0.2-v0.3 era; torchvision 0.26.0 as installed): Box2BoxTransform.apply_deltas,
DefaultAnchorGenerator, pairwise_iou, Boxes.scale/clip/nonempty, torchvision.ops.batched_nms / nms.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle import philox

SCALE_CLAMP = math.log(1000.0 / 16)


@dataclass
class PathParams:
    """Everything the path reads from cfg / the model object."""
    num_classes: int = 7
    num_anchors: int = 9
    num_convs: int = 4
    dropout_rate: float = 0.0
    cls_var: bool = False
    bbox_cov: bool = False
    cov_dims: int = 4
    cls_var_num_samples: int = 10
    box_num_samples: int = 1000                 # hard-coded in the reference (:356,:360)
    topk: int = 1000
    score_thresh: float = 0.05
    nms_thresh: float = 0.5
    max_dets: int = 100
    reg_weights: tuple = (1.0, 1.0, 1.0, 1.0)          # MODEL.RETINANET.BBOX_REG_WEIGHTS -> apply_deltas
    sample_reg_weights: tuple = (1.0, 1.0, 1.0, 1.0)   # MODEL.RPN.BBOX_REG_WEIGHTS -> SampleBox2BoxTransform (:175-176)
    affinity: float = 0.9
    box_merge: str = "bayesian_inference"
    cls_merge: str = "max_score"
    anchor_sizes: list = field(default_factory=lambda: [[x, x * 2 ** (1.0 / 3), x * 2 ** (2.0 / 3)]
                                                        for x in [32, 64, 128, 256, 512]])
    aspect_ratios: tuple = (0.5, 1.0, 2.0)
    strides: tuple = (8, 16, 32, 64, 128)
    anchor_offset: float = 0.0

    @property
    def use_dropout(self):
        return self.dropout_rate != 0.0

    @staticmethod
    def from_cfg(cfg):
        pm = cfg.MODEL.PROBABILISTIC_MODELING
        pi = cfg.PROBABILISTIC_INFERENCE
        ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
        return PathParams(
            num_classes=cfg.MODEL.RETINANET.NUM_CLASSES,
            num_anchors=len(cfg.MODEL.ANCHOR_GENERATOR.SIZES[0]) * len(ratios[0]),
            num_convs=cfg.MODEL.RETINANET.NUM_CONVS,
            dropout_rate=pm.DROPOUT_RATE,
            cls_var=pm.CLS_VAR_LOSS.NAME != "none",
            bbox_cov=pm.BBOX_COV_LOSS.NAME != "none",
            cov_dims=4 if pm.BBOX_COV_LOSS.COVARIANCE_TYPE == "diagonal" else 10,
            cls_var_num_samples=pm.CLS_VAR_LOSS.NUM_SAMPLES,
            topk=cfg.MODEL.RETINANET.TOPK_CANDIDATES_TEST,
            score_thresh=cfg.MODEL.RETINANET.SCORE_THRESH_TEST,
            nms_thresh=cfg.MODEL.RETINANET.NMS_THRESH_TEST,
            max_dets=cfg.TEST.DETECTIONS_PER_IMAGE,
            reg_weights=tuple(cfg.MODEL.RETINANET.BBOX_REG_WEIGHTS),
            sample_reg_weights=tuple(cfg.MODEL.RPN.BBOX_REG_WEIGHTS),
            affinity=pi.AFFINITY_THRESHOLD,
            box_merge=pi.BAYES_OD.BOX_MERGE_MODE,
            cls_merge=pi.BAYES_OD.CLS_MERGE_MODE,
            anchor_sizes=[list(s) for s in cfg.MODEL.ANCHOR_GENERATOR.SIZES],
            aspect_ratios=tuple(ratios[0]),
            anchor_offset=cfg.MODEL.ANCHOR_GENERATOR.OFFSET,
        )


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def unpack_head(sd, pp: PathParams):
    """state dict (reference key names) -> plain dict of (weight, bias) pairs."""
    step = 3 if pp.use_dropout else 2
    out = {"cls": [], "box": []}
    for i in range(pp.num_convs):
        out["cls"].append((sd["head.cls_subnet.%d.weight" % (i * step)], sd["head.cls_subnet.%d.bias" % (i * step)]))
        out["box"].append((sd["head.bbox_subnet.%d.weight" % (i * step)], sd["head.bbox_subnet.%d.bias" % (i * step)]))
    out["cls_score"] = (sd["head.cls_score.weight"], sd["head.cls_score.bias"])
    out["bbox_pred"] = (sd["head.bbox_pred.weight"], sd["head.bbox_pred.bias"])
    if pp.cls_var:
        out["cls_var"] = (sd["head.cls_var.weight"], sd["head.cls_var.bias"])
    if pp.bbox_cov:
        out["bbox_cov"] = (sd["head.bbox_cov.weight"], sd["head.bbox_cov.bias"])
    return out


# --------------------------------------------------------------------------------------
# head (probabilistic_retinanet.py:401-441, 512-523)
# --------------------------------------------------------------------------------------
class DropoutSource:
    """mode 'philox': masks from oracle/philox.py (parity).  mode 'torch': torch's own
    generator (CPU-baseline timing only).  mode 'off': eval-mode (identity)."""

    def __init__(self, mode, p, seed=0, image=0):
        self.mode, self.p, self.seed, self.image = mode, float(p), seed, image

    def __call__(self, x, level, sample, pass_, tower, layer):
        if self.mode == "off" or self.p == 0.0:
            return x
        if self.mode == "torch":
            return F.dropout(x, self.p, True)
        _, C, H, W = x.shape
        keep = philox.dropout_keep_mask(self.seed, self.image, sample, pass_, tower, layer, level, H, W, C, self.p)
        keep = torch.from_numpy(np.ascontiguousarray(keep.transpose(2, 0, 1)))[None]
        scale = torch.tensor(1.0, dtype=x.dtype) / torch.tensor(1.0 - self.p, dtype=x.dtype)
        return x * (keep.to(x.dtype) * scale)


def _tower(x, convs, drop, level, sample, pass_, tower):
    for layer, (w, b) in enumerate(convs):
        x = F.relu(F.conv2d(x, w, b, stride=1, padding=1))
        x = drop(x, level, sample, pass_, tower, layer)
    return x


def _hwa_k(t, K):
    """(1, A*K, H, W) -> (1, H*W*A, K): channel a*K+k, row (h*W+w)*A+a."""
    n, _, H, W = t.shape
    return t.view(n, -1, K, H, W).permute(0, 3, 4, 1, 2).reshape(n, -1, K)


def head_level(feat, hw, pp: PathParams, drop, level, sample):
    """One feature map through the head exactly as the reference evaluates it: the towers are
    run a SECOND time for the variance outputs (independent dropout draws, Q2)."""
    K = pp.num_classes
    out = {}
    out["box_cls"] = _hwa_k(F.conv2d(_tower(feat, hw["cls"], drop, level, sample, 0, 0), *hw["cls_score"], padding=1), K)
    out["box_delta"] = _hwa_k(F.conv2d(_tower(feat, hw["box"], drop, level, sample, 0, 1), *hw["bbox_pred"], padding=1), 4)
    out["box_cls_var"] = None
    out["box_reg_var"] = None
    if pp.cls_var:
        out["box_cls_var"] = _hwa_k(F.conv2d(_tower(feat, hw["cls"], drop, level, sample, 1, 0), *hw["cls_var"], padding=1), K)
    if pp.bbox_cov:
        out["box_reg_var"] = _hwa_k(F.conv2d(_tower(feat, hw["box"], drop, level, sample, 1, 1), *hw["bbox_cov"], padding=1), pp.cov_dims)
    return out


def head_outputs(feats, hw, pp: PathParams, drop, sample=0):
    """All levels of one forward -> dict of per-level lists (the reference's raw_output)."""
    keys = ("box_cls", "box_delta", "box_cls_var", "box_reg_var")
    res = {k: [] for k in keys}
    for level, f in enumerate(feats):
        o = head_level(f, hw, pp, drop, level, sample)
        for k in keys:
            res[k].append(o[k])
    for k in ("box_cls_var", "box_reg_var"):
        if res[k][0] is None:
            res[k] = None
    return res


# --------------------------------------------------------------------------------------
# anchors (detectron2 DefaultAnchorGenerator) and box transform (Box2BoxTransform)
# --------------------------------------------------------------------------------------
def cell_anchors(sizes, ratios):
    rows = []
    for s in sizes:
        area = s ** 2.0
        for r in ratios:
            w = math.sqrt(area / r)
            h = r * w
            rows.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(rows).float()


def make_anchors(level_hw, pp: PathParams):
    out = []
    for (gh, gw), stride, sizes in zip(level_hw, pp.strides, pp.anchor_sizes):
        base = cell_anchors(sizes, pp.aspect_ratios)
        sx = torch.arange(pp.anchor_offset * stride, gw * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(pp.anchor_offset * stride, gh * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        shifts = torch.stack((xx, yy, xx, yy), dim=1)
        out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return out


def apply_deltas(deltas, boxes, weights):
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=SCALE_CLAMP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.zeros_like(deltas)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw
    out[:, 3::4] = pcy + 0.5 * ph
    return out


def apply_samples_deltas(deltas, boxes, weights):
    """(M,4,S) deltas against (M,4,S) anchors (inference_utils.py:510-547)."""
    widths = boxes[:, 2, :] - boxes[:, 0, :]
    heights = boxes[:, 3, :] - boxes[:, 1, :]
    ctr_x = boxes[:, 0, :] + 0.5 * widths
    ctr_y = boxes[:, 1, :] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4, :] / wx, deltas[:, 1::4, :] / wy
    dw = torch.clamp(deltas[:, 2::4, :] / ww, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4, :] / wh, max=SCALE_CLAMP)
    pcx = dx * widths[:, None] + ctr_x[:, None]
    pcy = dy * heights[:, None] + ctr_y[:, None]
    pw = torch.exp(dw) * widths[:, None]
    ph = torch.exp(dh) * heights[:, None]
    out = torch.zeros_like(deltas)
    out[:, 0::4, :] = pcx - 0.5 * pw
    out[:, 1::4, :] = pcy - 0.5 * ph
    out[:, 2::4, :] = pcx + 0.5 * pw
    out[:, 3::4, :] = pcy + 0.5 * ph
    return out


# --------------------------------------------------------------------------------------
# statistics helpers
# --------------------------------------------------------------------------------------
def quirk_mean(per_sample: List[List[torch.Tensor]]):
    """Q1: (2*x_0 + x_1 + ... + x_{n-2}) / n per level (probabilistic_inference.py:216-222)."""
    acc = per_sample[0]
    for i in range(len(per_sample) - 1):
        acc = [acc[j] + per_sample[i][j] for j in range(len(acc))]
    return [a / len(per_sample) for a in acc]


def mean_covariance(samples):
    """inference_utils.py:337-371. samples: list of (N,k) or tensor (N,k,S)."""
    if isinstance(samples, torch.Tensor):
        n = samples.shape[2]
    else:
        n = len(samples)
        samples = torch.stack(samples, 2)
    mean = torch.mean(samples, 2, keepdim=True)
    r = torch.transpose(torch.unsqueeze(samples - mean, 1), 1, 3)
    cov = torch.sum(torch.matmul(r, torch.transpose(r, 3, 2)), 1) / (n - 1)
    return mean.squeeze(2), cov


def cholesky_from_output(v):
    """modeling_utils.py:4-22."""
    L = torch.diag_embed(torch.sqrt(torch.exp(v[:, 0:4])))
    if v.shape[1] > 4:
        ti = torch.tril_indices(row=4, col=4, offset=-1)
        L[:, ti[0], ti[1]] = v[:, 4:]
    return L


# --------------------------------------------------------------------------------------
# anchor-wise inference (probabilistic_inference.py:178-388)
# --------------------------------------------------------------------------------------
@dataclass
class Candidates:
    boxes: torch.Tensor                  # (M,4)
    cov: Optional[torch.Tensor]          # (M,4,4) or None
    scores: torch.Tensor                 # (M,)
    classes: torch.Tensor                # (M,) int64
    probs: torch.Tensor                  # (M,K)
    anchor_ids: np.ndarray               # (M,) global anchor id (level offset + index)
    level_counts: List[int]
    # diagnostics for margin-aware tests
    level_scores: Optional[List[torch.Tensor]] = None     # per level: max-prob of every anchor
    level_probs: Optional[List[torch.Tensor]] = None      # per level: (HWA,K) probabilities


def anchorwise(outputs_list, anchors, pp: PathParams, seed=0, image=0, keep_diag=False, stable_topk=True,
               normal_mode="philox", run=0):
    """outputs_list: list over samples/members of raw-output dicts (len 1 => no epistemic part).
    normal_mode 'torch' draws the logit / box noise from torch's own generator, as the reference does
    (CPU-baseline timing only; parity runs use the Philox streams)."""
    epistemic = len(outputs_list) > 1
    if epistemic:
        outputs = {"box_cls": quirk_mean([o["box_cls"] for o in outputs_list]),
                   "box_delta": quirk_mean([o["box_delta"] for o in outputs_list])}
        outputs["box_cls_var"] = (quirk_mean([o["box_cls_var"] for o in outputs_list])
                                  if outputs_list[0]["box_cls_var"] is not None else None)
        outputs["box_reg_var"] = (quirk_mean([o["box_reg_var"] for o in outputs_list])
                                  if outputs_list[0]["box_reg_var"] is not None else None)
    else:
        outputs = outputs_list[0]
    sizes = [a.shape[0] for a in anchors]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    all_delta, all_chol, all_anc, all_prob, all_cls, all_vec, all_epi, all_ids, counts = [], [], [], [], [], [], [], [], []
    lvl_scores, lvl_probs = [], []
    for i, anc in enumerate(anchors):
        box_cls = outputs["box_cls"][i][0]
        box_delta = outputs["box_delta"][i][0]
        if outputs["box_cls_var"] is not None:
            logvar = outputs["box_cls_var"][i][0]
            if normal_mode == "torch":
                eps = torch.randn((pp.cls_var_num_samples,) + tuple(box_cls.shape))
            else:
                eps = torch.from_numpy(philox.logit_normals(seed, image, i, pp.cls_var_num_samples,
                                                            box_cls.shape[0], box_cls.shape[1], run=run))
            draws = box_cls + eps * torch.sqrt(torch.exp(logvar))       # Normal.rsample: loc + eps*scale
            box_cls = torch.mean(draws.sigmoid_(), 0)
        else:
            box_cls = box_cls.clone().sigmoid_()
        num_topk = min(pp.topk, box_delta.size(0))
        prob, cls = torch.max(box_cls, 1)
        if keep_diag:
            lvl_scores.append(prob.clone())
            lvl_probs.append(box_cls.clone())
        if stable_topk:
            # define ties as "lower anchor index first" (SURVEY H4): stable descending sort
            order = torch.sort(prob, descending=True, stable=True)[1][:num_topk]
            top_prob, top_idx = prob[order], order
        else:
            top_prob, top_idx = prob.topk(num_topk)
        keep = top_prob > pp.score_thresh
        top_prob, top_idx = top_prob[keep], top_idx[keep]
        cls = cls[top_idx]
        d = box_delta[top_idx]
        a = anc[top_idx]
        chol = None
        if outputs["box_reg_var"] is not None:
            chol = cholesky_from_output(outputs["box_reg_var"][i][0][top_idx])
        epi = None
        if epistemic:
            decoded = [apply_deltas(o["box_delta"][i][0][top_idx], a, pp.reg_weights) for o in outputs_list]
            _, epi = mean_covariance(decoded)
        all_delta.append(d); all_chol.append(chol); all_anc.append(a); all_prob.append(top_prob)
        all_vec.append(box_cls[top_idx]); all_cls.append(cls); all_epi.append(epi)
        all_ids.append(top_idx.numpy().astype(np.int64) + int(offs[i])); counts.append(int(top_idx.numel()))
    delta = torch.cat(all_delta)
    anc = torch.cat(all_anc)
    ids = np.concatenate(all_ids)
    if isinstance(all_chol[0], torch.Tensor):
        L = torch.cat(all_chol)
        S = pp.box_num_samples
        if normal_mode == "torch":
            eps = torch.randn((S, delta.shape[0], 4))
        else:
            eps = torch.from_numpy(philox.box_normals(seed, image, ids, S, run=run))  # (S,M,4)
        draws = delta + torch.matmul(L, eps.unsqueeze(-1)).squeeze(-1)                  # loc + L eps
        draws = torch.transpose(torch.transpose(draws, 0, 1), 1, 2)                     # (M,4,S)
        anc_s = torch.repeat_interleave(anc.unsqueeze(2), S, dim=2)
        boxes, cov = mean_covariance(apply_samples_deltas(draws, anc_s, pp.sample_reg_weights))
        if isinstance(all_epi[0], torch.Tensor):
            cov += torch.cat(all_epi)
    else:
        cov = torch.cat(all_epi) if epistemic else None
        boxes = apply_deltas(delta, anc, pp.reg_weights)
    return Candidates(boxes, cov, torch.cat(all_prob), torch.cat(all_cls), torch.cat(all_vec), ids, counts,
                      lvl_scores if keep_diag else None, lvl_probs if keep_diag else None)


# --------------------------------------------------------------------------------------
# NMS (torchvision.ops.batched_nms semantics, CPU branch rule) -- two statements:
# the installed op, and a scalar restatement of its C++ loop used to pin tie-breaking.
# --------------------------------------------------------------------------------------
def nms_loop(boxes, scores, thr):
    """Scalar restatement of torchvision/csrc/ops/cpu/nms_kernel.cpp (fp32, strict >, stable
    descending order, lower index first on equal scores)."""
    b = boxes.numpy().astype(np.float32)
    s = scores.numpy().astype(np.float32)
    n = b.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    order = np.argsort(-s, kind="stable")
    sup = np.zeros(n, dtype=bool)
    keep = []
    zero = np.float32(0)
    for _i in range(n):
        i = order[_i]
        if sup[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(zero, (xx2 - xx1).astype(np.float32))
        h = np.maximum(zero, (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / ((areas[i] + areas[rest]).astype(np.float32) - inter).astype(np.float32)
        sup[rest[ovr.astype(np.float64) > thr]] = True
    return torch.from_numpy(np.asarray(keep, dtype=np.int64))


def batched_nms(boxes, scores, idxs, thr, impl="torchvision"):
    """torchvision/ops/boxes.py:51-120 as installed (0.26.0), CPU branch: numel > 4000 ->
    per-class loop ('vanilla'), else the coordinate-offset trick."""
    nms = nms_loop
    if impl == "torchvision":
        import torchvision
        nms = torchvision.ops.nms
    boxes = boxes.float()
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    if boxes.numel() > 4000:
        keep_mask = torch.zeros_like(scores, dtype=torch.bool)
        for c in torch.unique(idxs):
            cur = torch.where(idxs == c)[0]
            keep_mask[cur[nms(boxes[cur], scores[cur], thr)]] = True
        keep = torch.where(keep_mask)[0]
        return keep[torch.sort(scores[keep], descending=True, stable=True)[1]]
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, thr)


def pairwise_iou(b1, b2):
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype))


# --------------------------------------------------------------------------------------
# post-processing modes
# --------------------------------------------------------------------------------------
@dataclass
class Detections:
    boxes: torch.Tensor
    scores: torch.Tensor
    classes: torch.Tensor
    probs: torch.Tensor
    cov: torch.Tensor
    image_size: tuple
    keep: Optional[torch.Tensor] = None      # NMS survivor indices into the candidate list


def standard_nms_post(c: Candidates, pp: PathParams, image_hw, nms_impl="torchvision"):
    """inference_utils.py:12-54."""
    keep = batched_nms(c.boxes, c.scores, c.classes, pp.nms_thresh, nms_impl)[: pp.max_dets]
    cov = c.cov[keep] if isinstance(c.cov, torch.Tensor) else torch.zeros(c.boxes[keep].shape + (4,))
    return Detections(c.boxes[keep], c.scores[keep], c.classes[keep], c.probs[keep], cov, tuple(image_hw), keep)


def bayesian_box_fusion(means, covs, mode):
    """inference_utils.py:292-334 (numpy / LAPACK, fp32 in -> fp32 out)."""
    precs = np.linalg.inv(covs)
    if mode == "bayesian_inference":
        final_cov = np.linalg.inv(precs.sum(0))
        final_mean = np.matmul(precs, np.expand_dims(means, 2)).sum(0)
        final_mean = np.squeeze(np.matmul(final_cov, final_mean))
    elif mode == "covariance_intersection":
        diff = precs.sum(0) - precs
        d_p = np.linalg.det(precs)
        d_tot = np.linalg.det(precs.sum(0))
        d_diff = np.linalg.det(diff)
        omegas = (d_tot - d_diff + d_p) / (precs.shape[0] * d_tot + (d_p - d_diff).sum(0))
        wp = np.expand_dims(omegas, (1, 2)) * precs
        final_cov = np.linalg.inv(wp.sum(0))
        final_mean = np.matmul(final_cov, np.matmul(wp, np.expand_dims(means, 2)).sum(0))
    else:
        raise ValueError(mode)
    return final_mean, final_cov


def bayes_od_post(c: Candidates, pp: PathParams, image_hw, nms_impl="torchvision", dtype=np.float32):
    """probabilistic_inference.py:536-636. dtype=np.float64 gives the high-precision ground truth
    used for condition-aware tolerances."""
    keep = batched_nms(c.boxes, c.scores, c.classes, pp.nms_thresh, nms_impl)[: pp.max_dets]
    iou = pairwise_iou(c.boxes, c.boxes)
    member = iou[keep, :] > pp.affinity
    vec_list, box_list, cov_list = [], [], []
    centers = c.probs[keep]
    for row, center in zip(member, centers):
        cluster_probs = c.probs[row]
        _, center_cat = torch.max(center, 0)
        _, cat = cluster_probs.max(1)
        same = cat == center_cat
        if pp.cls_merge == "bayesian_inference":
            vec_list.append(cluster_probs.mean(0).unsqueeze(0))
        else:
            vec_list.append(center.unsqueeze(0))
        mu = c.boxes[row, :][same].numpy().astype(dtype)
        sg = c.cov[row, :][same].numpy().astype(dtype)
        m, s = bayesian_box_fusion(mu, sg, pp.box_merge)
        box_list.append(torch.from_numpy(np.squeeze(m)))
        cov_list.append(torch.from_numpy(s))
    if len(box_list) > 0:
        if pp.cls_merge == "bayesian_inference":
            probs = torch.cat(vec_list, 0)
            scores, classes = torch.max(probs, 1)
        else:
            probs, scores, classes = c.probs[keep], c.scores[keep], c.classes[keep]
        return Detections(torch.stack(box_list, 0), scores, classes, probs, torch.stack(cov_list, 0),
                          tuple(image_hw), keep)
    return Detections(c.boxes, torch.zeros(c.boxes.shape[0]), c.classes, c.probs,
                      torch.empty(c.boxes.shape + (4,)), tuple(image_hw), keep)


def anchor_statistics_post(c: Candidates, pp: PathParams, image_hw, nms_impl="torchvision"):
    """inference_utils.py:57-162 (general_anchor_statistics_postprocessing)."""
    iou = pairwise_iou(c.boxes, c.boxes)
    keep = batched_nms(c.boxes, c.scores, c.classes, pp.nms_thresh, nms_impl)[: pp.max_dets]
    member = iou[keep, :] > pp.affinity
    has_cov = isinstance(c.cov, torch.Tensor) and len(c.cov) > 0
    vec_list, box_list, cov_list = [], [], []
    for row, center in zip(member, keep):
        if row.sum(0) >= 2:
            same = c.classes[row] == c.classes[center]
            cluster = c.boxes[row, :][same, :]
            mean = cluster.mean(0)
            res = (cluster - mean).unsqueeze(2)
            cov = torch.sum(torch.matmul(res, torch.transpose(res, 2, 1)), 0) / max((cluster.shape[0] - 1), 1.0)
            if has_cov:
                cov = cov + c.cov[row, :][same, :].mean(0)
            vec = c.probs[row, :][same, :].mean(0)
        else:
            mean = c.boxes[center]
            vec = c.probs[center]
            cov = c.cov[center] if has_cov else 1e-4 * torch.eye(4, 4)
        box_list.append(mean); cov_list.append(cov); vec_list.append(vec)
    if len(box_list) > 0:
        probs = torch.stack(vec_list, 0)
        scores, classes = torch.max(probs, 1)
        return Detections(torch.stack(box_list, 0), scores, classes, probs, torch.stack(cov_list, 0), tuple(image_hw), keep)
    return Detections(c.boxes, torch.zeros(c.boxes.shape[0]), c.classes, c.probs, torch.empty(c.boxes.shape + (4,)),
                      tuple(image_hw), keep)


def black_box_post(dets: List[Detections], pp: PathParams, image_hw, nms_impl="torchvision"):
    """inference_utils.py:165-289 (general_black_box_ensembles_post_processing): merge of per-run,
    post-NMS detections by sequential IoU clustering, then one more NMS."""
    boxes = torch.cat([d.boxes for d in dets], 0)
    covs = torch.cat([d.cov for d in dets], 0)
    probs = torch.cat([d.probs for d in dets], 0)
    cls = torch.cat([d.classes for d in dets], 0)
    iou = pairwise_iou(boxes, boxes)
    clusters = []
    for i in range(iou.shape[0]):
        if i != 0:
            allc = torch.cat(clusters, 0)
            if (allc == i).any():
                continue
        clusters.extend(torch.where((iou[i, :] >= pp.affinity) & (cls == cls[i])))
    box_list, cov_list, vec_list = [], [], []
    for cl in clusters:
        bc, cc = boxes[cl], covs[cl]
        if bc.shape[0] >= 2:
            mean = bc.mean(0)
            res = (bc - mean).unsqueeze(2)
            cov = torch.sum(torch.matmul(res, torch.transpose(res, 2, 1)), 0) / (bc.shape[0] - 1)
            cov = cov + cc.mean(0)
            box_list.append(mean); cov_list.append(cov); vec_list.append(probs[cl].mean(0))
        else:
            box_list.append(boxes[cl].mean(0)); cov_list.append(covs[cl].mean(0)); vec_list.append(probs[cl].mean(0))
    if len(box_list) > 0:
        pv = torch.stack(vec_list, 0)
        score, classes = torch.max(pv, 1)
        mb = torch.stack(box_list, 0)
        keep = batched_nms(mb, score, classes, pp.nms_thresh, nms_impl)[: pp.max_dets]
        return Detections(mb[keep], score[keep], classes[keep], pv[keep], torch.stack(cov_list, 0)[keep], tuple(image_hw), keep)
    return Detections(boxes, torch.zeros(boxes.shape[0]), cls, probs, torch.empty(boxes.shape + (4,)), tuple(image_hw), None)


def detector_postprocess(d: Detections, out_h, out_w):
    """inference_utils.py:374-425 with detectron2 Boxes.scale/clip/nonempty."""
    sx, sy = out_w / d.image_size[1], out_h / d.image_size[0]
    boxes = d.boxes.clone().float()
    boxes[:, 0::2] *= sx
    boxes[:, 1::2] *= sy
    boxes[:, 0].clamp_(min=0, max=out_w); boxes[:, 1].clamp_(min=0, max=out_h)
    boxes[:, 2].clamp_(min=0, max=out_w); boxes[:, 3].clamp_(min=0, max=out_h)
    ne = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    cov = d.cov[ne] + 1e-4 * torch.eye(4)
    Smat = torch.diag_embed(torch.as_tensor((sx, sy, sx, sy))).unsqueeze(0)
    Smat = torch.repeat_interleave(Smat, cov.shape[0], 0).to(cov.dtype)
    cov = torch.matmul(torch.matmul(Smat, cov), torch.transpose(Smat, 2, 1))
    return Detections(boxes[ne], d.scores[ne], d.classes[ne], d.probs[ne], cov, (out_h, out_w),
                      d.keep[ne] if d.keep is not None and d.keep.shape[0] == ne.shape[0] else None)


def covar_xyxy_to_xywh(cov):
    """inference_utils.py:428-451."""
    T = torch.as_tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [-1.0, 0, 1.0, 0], [0, -1.0, 0, 1.0]]).unsqueeze(0)
    T = torch.repeat_interleave(T, cov.shape[0], 0).to(cov.dtype)
    return torch.matmul(torch.matmul(T, cov), torch.transpose(T, 2, 1))


def detections_to_json(d: Detections, img_id, cat_mapping):
    """inference_utils.py:454-502."""
    n = d.boxes.shape[0]
    if n == 0:
        return []
    b = d.boxes.clone()
    b[:, 2] -= b[:, 0]
    b[:, 3] -= b[:, 1]
    b = b.numpy().tolist()
    scores = d.scores.tolist()
    classes = [cat_mapping[c] if c in cat_mapping else -1 for c in d.classes.tolist()]
    probs = d.probs.tolist()
    cov = covar_xyxy_to_xywh(d.cov).tolist()
    return [{"image_id": img_id, "category_id": classes[k], "bbox": b[k], "score": scores[k],
             "cls_prob": probs[k], "bbox_covar": cov[k]} for k in range(n) if classes[k] != -1]


def read_results_json(entries, min_allowed_score=0.0):
    """The READER's side of the wire contract: src/core/evaluation_tools/evaluation_utils.py:28-69
    (`eval_predictions_preprocess`): skips category -1 and low scores, turns XYWH boxes back into XYXY in float64 numpy and
    the covariance back with T' = [[1,0,0,0],[0,1,0,0],[1,0,1,0],[0,1,0,1]], then casts to fp32 tensors per image."""
    boxes, probs, covs = {}, {}, {}
    Tm = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [1.0, 0, 1.0, 0], [0, 1.0, 0.0, 1.0]])
    for e in entries:
        if e["category_id"] == -1 or np.array(e["cls_prob"]).max(0) < min_allowed_score:
            continue
        b = e["bbox"]
        xyxy = np.array([b[0], b[1], b[0] + b[2], b[1] + b[3]])
        cov = np.matmul(np.matmul(Tm, np.array(e["bbox_covar"])), Tm.T)
        k = e["image_id"]
        boxes.setdefault(k, []).append(torch.as_tensor(xyxy, dtype=torch.float32))
        probs.setdefault(k, []).append(torch.as_tensor(e["cls_prob"], dtype=torch.float32))
        covs.setdefault(k, []).append(torch.as_tensor(cov, dtype=torch.float32))
    return ({k: torch.stack(v) for k, v in boxes.items()}, {k: torch.stack(v) for k, v in probs.items()},
            {k: torch.stack(v) for k, v in covs.items()})


# --------------------------------------------------------------------------------------
# whole path, one image:  features -> detections   (predictor.__call__, :86-111)
# --------------------------------------------------------------------------------------
def predict(feats, weight_sets, pp: PathParams, mode, image_hw, out_hw=None, n_mc=1, seed=0, image=0,
            dropout_mode="philox", nms_impl="torchvision", return_candidates=False, keep_diag=False, post_nms=False,
            mc_single=False):
    """mode: 'standard_nms' | 'mc_dropout_ensembles' (pre_nms) | 'ensembles' (pre_nms) | 'bayes_od' |
    'anchor_statistics'.
    weight_sets: list of unpacked heads (len E for 'ensembles', else 1). n_mc>1 enables MC-dropout
    (model.train(), probabilistic_inference.py:52-56)."""
    out_hw = out_hw or image_hw
    per_member = isinstance(feats[0], (list, tuple))      # ensembles: feats[e][l], one feature set per member
    level_hw = [tuple(f.shape[-2:]) for f in (feats[0] if per_member else feats)]
    anchors = make_anchors(level_hw, pp)
    if mode == "ensembles":
        # every member is a full model with its own backbone (probabilistic_inference.py:58-77,499-501)
        drop = DropoutSource("off", 0.0)
        outs = [head_outputs(feats[e] if per_member else feats, hw, pp, drop) for e, hw in enumerate(weight_sets)]
    elif mc_single:
        # MC_DROPOUT.ENABLE with NUM_RUNS == 1: the model stays in train() (:52-56) and the single forward (:272-273)
        # has active dropout; no epistemic term
        drop = DropoutSource(dropout_mode, pp.dropout_rate, seed, image)
        outs = [head_outputs(feats, weight_sets[0], pp, drop, sample=0)]
    elif n_mc > 1:
        drop = DropoutSource(dropout_mode, pp.dropout_rate, seed, image)
        outs = [head_outputs(feats, weight_sets[0], pp, drop, sample=s) for s in range(n_mc)]
    else:
        # a single forward; dropout is active only if the predictor put the model in train()
        # (MC_DROPOUT.ENABLE with NUM_RUNS == 1), which the shipped configs never do.
        drop = DropoutSource("off", 0.0)
        outs = [head_outputs(feats, weight_sets[0], pp, drop)]
    if post_nms:
        # probabilistic_inference.py:444-481 / 506-534: every run is a complete single-sample inference
        # (own noise draws, no epistemic term) followed by standard NMS; the runs are merged afterwards
        runs = []
        for r, o in enumerate(outs):
            c = anchorwise([o], anchors, pp, seed, image, run=r)
            runs.append(standard_nms_post(c, pp, image_hw, nms_impl))
        det = black_box_post(runs, pp, image_hw, nms_impl)
        final = detector_postprocess(det, out_hw[0], out_hw[1])
        if return_candidates:
            return final, runs, det
        return final
    cand = anchorwise(outs, anchors, pp, seed, image, keep_diag=keep_diag)
    if mode == "bayes_od":
        det = bayes_od_post(cand, pp, image_hw, nms_impl)
    elif mode == "anchor_statistics":
        det = anchor_statistics_post(cand, pp, image_hw, nms_impl)
    else:
        det = standard_nms_post(cand, pp, image_hw, nms_impl)
    final = detector_postprocess(det, out_hw[0], out_hw[1])
    if return_candidates:
        return final, cand, det
    return final
