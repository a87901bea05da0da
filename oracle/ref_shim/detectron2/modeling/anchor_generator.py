import math
import torch
from detectron2.structures import Boxes


class DefaultAnchorGenerator(torch.nn.Module):
    """detectron2.modeling.anchor_generator.DefaultAnchorGenerator (v0.2-v0.3)."""

    def __init__(self, cfg, input_shape):
        super().__init__()
        sizes = cfg.MODEL.ANCHOR_GENERATOR.SIZES
        aspect_ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
        self.strides = [x.stride for x in input_shape]
        self.offset = cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        n = len(self.strides)
        if len(sizes) == 1:
            sizes = list(sizes) * n
        if len(aspect_ratios) == 1:
            aspect_ratios = list(aspect_ratios) * n
        self.cell_anchors = [self.generate_cell_anchors(s, a).float()
                             for s, a in zip(sizes, aspect_ratios)]

    @property
    def num_cell_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @staticmethod
    def generate_cell_anchors(sizes, aspect_ratios):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for aspect_ratio in aspect_ratios:
                w = math.sqrt(area / aspect_ratio)
                h = aspect_ratio * w
                x0, y0, x1, y1 = -w / 2.0, -h / 2.0, w / 2.0, h / 2.0
                anchors.append([x0, y0, x1, y1])
        return torch.tensor(anchors)

    def forward(self, features):
        out = []
        for feat, stride, base in zip(features, self.strides, self.cell_anchors):
            gh, gw = feat.shape[-2:]
            shifts_x = torch.arange(self.offset * stride, gw * stride, step=stride, dtype=torch.float32)
            shifts_y = torch.arange(self.offset * stride, gh * stride, step=stride, dtype=torch.float32)
            shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
            shift_x = shift_x.reshape(-1)
            shift_y = shift_y.reshape(-1)
            shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
            out.append(Boxes((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4)))
        return out


def build_anchor_generator(cfg, input_shape):
    return DefaultAnchorGenerator(cfg, input_shape)
