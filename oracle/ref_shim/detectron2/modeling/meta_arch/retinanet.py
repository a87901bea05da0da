import torch
from torch import nn
from detectron2.layers import ShapeSpec
from detectron2.modeling.anchor_generator import build_anchor_generator
from detectron2.modeling.box_regression import Box2BoxTransform


def permute_to_N_HWA_K(tensor, K):
    """(N, Ai*K, H, W) -> (N, H*W*Ai, K)   [detectron2 meta_arch/retinanet.py]"""
    assert tensor.dim() == 4, tensor.shape
    N, _, H, W = tensor.shape
    tensor = tensor.view(N, -1, K, H, W)
    tensor = tensor.permute(0, 3, 4, 1, 2)
    tensor = tensor.reshape(N, -1, K)
    return tensor


class _ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes


class FeatureInjectionBackbone(nn.Module):
    """Stand-in for the ResNet-FPN backbone (out of the path under test): returns the FPN
    maps the runner supplies, so the reference's forward() is exercised unmodified from
    `features = self.backbone(images.tensor)` on (probabilistic_retinanet.py:99)."""

    def __init__(self, in_features, channels=256, strides=(8, 16, 32, 64, 128)):
        super().__init__()
        self._names = list(in_features)
        self._shapes = {n: ShapeSpec(channels=channels, stride=s) for n, s in zip(self._names, strides)}
        self.current = None

    def output_shape(self):
        return self._shapes

    def forward(self, x):
        assert self.current is not None, "runner must set backbone.current = {name: tensor}"
        return self.current


class RetinaNet(nn.Module):
    """The attributes/methods of detectron2's RetinaNet that the reference subclass and
    predictor read (both API generations: v0.2 `in_features/score_threshold`, v0.3+
    `head_in_features/test_score_thresh`)."""

    def __init__(self, cfg):
        super().__init__()
        self.num_classes = cfg.MODEL.RETINANET.NUM_CLASSES
        self.in_features = cfg.MODEL.RETINANET.IN_FEATURES
        self.head_in_features = self.in_features
        self.focal_loss_alpha = cfg.MODEL.RETINANET.FOCAL_LOSS_ALPHA
        self.focal_loss_gamma = cfg.MODEL.RETINANET.FOCAL_LOSS_GAMMA
        self.smooth_l1_beta = cfg.MODEL.RETINANET.SMOOTH_L1_LOSS_BETA
        self.test_score_thresh = self.score_threshold = cfg.MODEL.RETINANET.SCORE_THRESH_TEST
        self.test_topk_candidates = self.topk_candidates = cfg.MODEL.RETINANET.TOPK_CANDIDATES_TEST
        self.test_nms_thresh = self.nms_threshold = cfg.MODEL.RETINANET.NMS_THRESH_TEST
        self.max_detections_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        self.vis_period = 0
        self.input_format = cfg.INPUT.FORMAT
        self.backbone = FeatureInjectionBackbone(self.in_features, cfg.MODEL.FPN.OUT_CHANNELS)
        backbone_shape = self.backbone.output_shape()
        feature_shapes = [backbone_shape[f] for f in self.in_features]
        self.head = RetinaNetHead(cfg, feature_shapes)
        self.anchor_generator = build_anchor_generator(cfg, feature_shapes)
        self.box2box_transform = Box2BoxTransform(weights=cfg.MODEL.RETINANET.BBOX_REG_WEIGHTS)
        self.register_buffer("pixel_mean", torch.Tensor(cfg.MODEL.PIXEL_MEAN).view(-1, 1, 1))
        self.register_buffer("pixel_std", torch.Tensor(cfg.MODEL.PIXEL_STD).view(-1, 1, 1))
        self.loss_normalizer = 100
        self.loss_normalizer_momentum = 0.9

    @property
    def device(self):
        return self.pixel_mean.device

    def preprocess_image(self, batched_inputs):
        images = [x["image"].to(self.device) for x in batched_inputs]
        return _ImageList(torch.stack([im.float() for im in images]), [im.shape[-2:] for im in images])


class RetinaNetHead(nn.Module):
    def __init__(self, cfg, input_shape):
        super().__init__()
