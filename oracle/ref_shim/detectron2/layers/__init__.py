from collections import namedtuple
import torch
import torchvision


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
    def __new__(cls, channels=None, height=None, width=None, stride=None):
        return super().__new__(cls, channels, height, width, stride)


def cat(tensors, dim=0):
    """detectron2.layers.cat: single-element passthrough, else torch.cat."""
    assert isinstance(tensors, (list, tuple))
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2.layers.batched_nms -> torchvision.ops.batched_nms on fp32 boxes."""
    assert boxes.shape[-1] == 4
    return torchvision.ops.batched_nms(boxes.float(), scores, idxs, iou_threshold)
