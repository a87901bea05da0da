"""Seeded synthetic inputs for the probabilistic head path (no datasets / checkpoints offline).

Geometry and initialisation follow SURVEY.md section 8(d):
  * FPN maps (B, 256, ceil(H/s), ceil(W/s)) for strides 8..128 of the image padded to a
    multiple of 128 (the benchmark geometry SURVEY 8(d) quotes; detectron2 itself pads to 32, see level_shapes);
  * head weight sets keyed like the reference module's state dict
    (reference src/probabilistic_modeling/probabilistic_retinanet.py:401-484): tower convs
    `head.{cls,bbox}_subnet.<i>` with i stepping by 3 when dropout layers are present and
    by 2 otherwise, outputs `head.cls_score / bbox_pred / cls_var / bbox_cov`.
The reference's own N(0, 0.01) initialisation collapses activations and, with the
-log(99) prior bias, yields no score above SCORE_THRESH_TEST; the synthetic
sets use He-scaled towers and output scales that make top-k, the 0.05 threshold and
NMS all binding.
"""
import math

import torch

STRIDES = (8, 16, 32, 64, 128)


def padded_size(height, width, divisibility=128):
    ph = (height + divisibility - 1) // divisibility * divisibility
    pw = (width + divisibility - 1) // divisibility * divisibility
    return ph, pw


def level_shapes(height, width, strides=STRIDES, divisibility=None):
    """FPN map sizes.  divisibility None: the benchmark geometry of SURVEY 8(d) / BASELINE.md (frame padded to a
    multiple of the largest stride, 720x1280 -> 768x1280, 20 460 locations).  divisibility 32: what detectron2's
    RetinaNet backbone produces (frame padded to the res5 stride; P6 / P7 from stride-2 3x3 convolutions with
    padding 1: 720x1280 -> 92x160, 46x80, 23x40, 12x20, 6x10 = 19 620 locations), cf. backbone.py."""
    if divisibility is None:
        ph, pw = padded_size(height, width, max(strides))
        return [(ph // s, pw // s) for s in strides]
    ph, pw = padded_size(height, width, divisibility)
    shapes = []
    for s in strides:
        if s <= divisibility:
            shapes.append((ph // s, pw // s))
        else:
            h, w = shapes[-1]
            shapes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return shapes


def make_features(seed, image_idx, height, width, channels=256, strides=STRIDES, dtype=torch.float32, divisibility=None):
    """FPN maps of ONE image: list of (1, C, Hl, Wl) fp32 CPU tensors.  `seed` selects an independent feature set
    for the same image (ensemble member e uses seed e: every member of the reference has its own backbone)."""
    g = torch.Generator().manual_seed(4321 + 7919 * int(seed) + int(image_idx))
    return [torch.randn((1, channels, h, w), generator=g, dtype=dtype)
            for (h, w) in level_shapes(height, width, strides, divisibility)]


def make_member_features(members, image_idx, height, width, divisibility=None):
    """Feature maps of ONE image as E ensemble members see it: every member of the reference is a full model with its
    own backbone (reference probabilistic_inference.py:58-77,499-501), so the maps differ per member while
    describing the same image -- modelled as a shared component plus a member-specific one (unit variance overall).
    Returns feats[e][l]."""
    base = make_features(0, image_idx, height, width, divisibility=divisibility)
    out = []
    for e in range(members):
        own = make_features(100 + e, image_idx, height, width, divisibility=divisibility)
        out.append([(0.9 * b + 0.4358898943540674 * o).contiguous() for b, o in zip(base, own)])
    return out


def make_head_state_dict(seed, num_classes=7, num_anchors=9, channels=256, num_convs=4,
                         use_dropout=True, cls_var=True, bbox_cov=True, cov_dims=4,
                         logit_scale=0.8, logit_bias=-4.6):
    """Random-init weight set `seed` (cf. RANDOM_SEED_NUMS, reference src/core/setup.py:132-133)."""
    g = torch.Generator().manual_seed(1000003 * int(seed) + 17)
    fan = channels * 9
    sd = {}

    def conv(name, cout, std, bias):
        sd[name + ".weight"] = torch.randn((cout, channels, 3, 3), generator=g) * std
        sd[name + ".bias"] = torch.full((cout,), float(bias))

    step = 3 if use_dropout else 2
    for tower in ("cls_subnet", "bbox_subnet"):
        for i in range(num_convs):
            conv("head.%s.%d" % (tower, i * step), channels, math.sqrt(2.0 / fan), 0.0)
            # small non-zero biases so the bias path is exercised
            sd["head.%s.%d.bias" % (tower, i * step)] = torch.randn((channels,), generator=g) * 0.05
    conv("head.cls_score", num_anchors * num_classes, logit_scale / math.sqrt(fan), logit_bias)
    conv("head.bbox_pred", num_anchors * 4, 0.1 / math.sqrt(fan), 0.0)
    if cls_var:
        conv("head.cls_var", num_anchors * num_classes, 0.5 / math.sqrt(fan), -4.0)
    if bbox_cov:
        conv("head.bbox_cov", num_anchors * cov_dims, 0.3 / math.sqrt(fan), -6.0)
        if cov_dims > 4:
            # off-diagonal Cholesky entries enter un-exponentiated: keep them small
            w = sd["head.bbox_cov.weight"].view(num_anchors, cov_dims, channels, 3, 3)
            b = sd["head.bbox_cov.bias"].view(num_anchors, cov_dims)
            w[:, 4:] *= 0.02
            b[:, 4:] = 0.0
    return sd


def make_member_state_dicts(members, correlation=0.95, **kw):
    """E weight sets that behave like members of a trained ensemble: independently initialised networks trained on the
    same data agree on most predictions, so the members share a common component (correlation) plus a member-specific
    one.  Fully independent random heads (make_head_state_dict(1000 * e)) average their logits to nothing: with the
    reference's Q1 mean over 5 such members no anchor passes SCORE_THRESH_TEST and the post-processing would run on
    empty candidate lists.  Biases keep their (deterministic) values."""
    base = make_head_state_dict(0, **kw)
    own_w = math.sqrt(1.0 - correlation * correlation)
    out = []
    for e in range(members):
        own = make_head_state_dict(1000 * (e + 1), **kw)
        out.append({k: ((correlation * base[k] + own_w * own[k]) if k.endswith(".weight") else base[k].clone()) for k in base})
    return out


def make_image(seed, image_idx, height=720, width=1280):
    g = torch.Generator().manual_seed(1234 + 7919 * int(seed) + int(image_idx))
    return torch.randint(0, 256, (3, height, width), generator=g, dtype=torch.uint8)


def make_planted_candidates(seed, num_gt=20, per_gt=(1, 40), image_hw=(720, 1280), num_classes=7,
                            jitter=1.5, score_margin=1e-3):
    """Planted-cluster candidate set for stage-isolated NMS / BayesOD tests:
    jittered copies of `num_gt` boxes so that IoU>0.9 clusters of several sizes exist;
    scores are distinct, spaced by min(score_margin, 0.93/M).
    Returns boxes (M,4), cov (M,4,4) SPD, scores (M,), classes (M,) int64, prob vectors (M,K)."""
    g = torch.Generator().manual_seed(2024 + int(seed))
    H, W = image_hw
    boxes, classes = [], []
    for _ in range(num_gt):
        w = float(torch.empty(1).uniform_(60, 400, generator=g))
        h = float(torch.empty(1).uniform_(60, 300, generator=g))
        x = float(torch.empty(1).uniform_(0, W - w, generator=g))
        y = float(torch.empty(1).uniform_(0, H - h, generator=g))
        n = int(torch.randint(per_gt[0], per_gt[1] + 1, (1,), generator=g))
        c = int(torch.randint(0, num_classes, (1,), generator=g))
        base = torch.tensor([x, y, x + w, y + h])
        jit = torch.randn((n, 4), generator=g) * jitter
        boxes.append(base[None] + jit)
        cls = torch.full((n,), c, dtype=torch.int64)
        flip = torch.rand((n,), generator=g) < 0.15  # some members of another class
        cls[flip] = (c + 1) % num_classes
        classes.append(cls)
    boxes = torch.cat(boxes).float()
    classes = torch.cat(classes)
    M = boxes.shape[0]
    perm = torch.randperm(M, generator=g)
    boxes, classes = boxes[perm], classes[perm]
    # distinct scores with a guaranteed margin, in (0.05, 1)
    ranks = torch.randperm(M, generator=g).float()
    scores = (0.06 + ranks * min(score_margin, 0.93 / max(M, 1))).float()
    probs = torch.rand((M, num_classes), generator=g) * 0.04
    probs[torch.arange(M), classes] = scores
    A = torch.randn((M, 4, 4), generator=g) * 1.5
    cov = A @ A.transpose(1, 2) + torch.eye(4)[None] * 2.0
    return boxes, cov.float(), scores, classes, probs.float()
