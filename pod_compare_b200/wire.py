"""Wire format of the path: detections -> the reference's `coco_instances_results.json` entries.

The reference converts one image at a time on the host with five `.cpu().tolist()` synchronisations
(src/probabilistic_inference/inference_utils.py:454-502, called from src/apply_net.py:91-96).  Here ONE kernel
(`pod_wire_records`, csrc/wire.cu) turns the detections of a whole batch into fixed-size records already in the JSON
layout -- XYWH boxes, `T Sigma T^T` covariances (inference_utils.py:428-451), dataset category ids -- ONE asynchronous
device->host copy into pinned memory brings them over, and the dict list is built from a single numpy view.
The schema is the reader's contract (src/core/evaluation_tools/evaluation_utils.py:28-69):
    {"image_id", "category_id", "bbox": [x, y, w, h], "score", "cls_prob": [K], "bbox_covar": [[4x4]]}
detections whose class has no id in the test dataset (category -1) are dropped (inference_utils.py:489).
"""
import numpy as np
import torch

from . import ops

# class lists of the reference's datasets (src/core/datasets/metadata.py:8-21): dataset id = contiguous id + 1
BDD_THING_CLASSES = ['car', 'bus', 'truck', 'person', 'rider', 'bike', 'motor']
KITTI_THING_CLASSES = ['car', 'person']
BDD_THING_DATASET_ID_TO_CONTIGUOUS_ID = {i + 1: i for i in range(len(BDD_THING_CLASSES))}
KITTI_THING_DATASET_ID_TO_CONTIGUOUS_ID = {i + 1: i for i in range(len(KITTI_THING_CLASSES))}
BDD_TO_KITTI_CONTIGUOUS_ID = {BDD_THING_CLASSES.index(c): KITTI_THING_CLASSES.index(c) for c in KITTI_THING_CLASSES}


def build_category_mapping(train_dataset, test_dataset, train_id_to_contiguous, test_id_to_contiguous,
                           coco_to_voc_contiguous=None, bdd_to_kitti_contiguous=None):
    """contiguous class id of the network -> category id of the TEST dataset, as src/apply_net.py:53-79 builds it from
    the two datasets' `thing_dataset_id_to_contiguous_id` tables:
      * same tables (or out-of-distribution training set 'coco_not_in_voc_2017_train'): the flipped test table;
      * otherwise (BDD -> KITTI, COCO -> VOC): the flipped test table re-keyed through the training -> test
        contiguous-id map, so classes the test dataset lacks have no entry (-> category -1, dropped by the writer)."""
    flipped = {v: k for k, v in test_id_to_contiguous.items()}
    if train_id_to_contiguous == test_id_to_contiguous or train_dataset == 'coco_not_in_voc_2017_train':
        return flipped
    if 'voc' in test_dataset and 'coco' in train_dataset:
        if coco_to_voc_contiguous is None:
            raise ValueError("COCO -> VOC needs the COCO_TO_VOC_CONTIGUOUS_ID table")
        dataset_mapping = {v: k for k, v in coco_to_voc_contiguous.items()}
    elif 'kitti' in test_dataset and 'bdd' in train_dataset:
        table = BDD_TO_KITTI_CONTIGUOUS_ID if bdd_to_kitti_contiguous is None else bdd_to_kitti_contiguous
        dataset_mapping = {v: k for k, v in table.items()}
    else:
        # the reference constructs (and forgets to raise) this error and then fails on the undefined mapping
        raise ValueError('Cannot generate category mapping dictionary. Please check if training and inference datasets '
                         'are compatible.')
    return {dataset_mapping[k]: v for k, v in flipped.items()}


def category_map_tensor(cat_mapping_dict, num_classes, device):
    """dict (contiguous class -> dataset id) -> int32 device array of K entries, -1 where the class has no id."""
    m = torch.full((num_classes,), -1, dtype=torch.int32)
    for c, cid in (cat_mapping_dict or {}).items():
        if 0 <= int(c) < num_classes:
            m[int(c)] = int(cid)
    return m.to(device)


class BatchJsonWriter:
    """detections of a batch (the `det` dict of the engine) -> list of result dicts, one D2H copy per batch."""

    def __init__(self, num_classes, max_dets, cat_mapping_dict, device):
        self.K, self.D = int(num_classes), int(max_dets)
        self.device = torch.device(device)
        # the reference maps through the dict only when one is given... and crashes on None (:475-477); None here means
        # "report the contiguous class id"
        self.cat_map = category_map_tensor(cat_mapping_dict, self.K, self.device) if cat_mapping_dict is not None else None
        self._pinned = None

    def records_async(self, det):
        """Launch the record kernel and the device->host copy on the current stream; returns (pinned host tensor, event)."""
        rec = ops.wire_records(det, xywh=True, cat_map=self.cat_map)
        if self._pinned is None or self._pinned.shape != rec.shape:
            self._pinned = torch.empty(rec.shape, dtype=torch.float32, pin_memory=True)
        self._pinned.copy_(rec, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return self._pinned, ev

    def to_json(self, det, image_ids):
        host, ev = self.records_async(det)
        ev.synchronize()
        return records_to_json(host.numpy(), image_ids, self.K, self.D)


def records_to_json(records, image_ids, num_classes, max_dets):
    """(B, 1 + max_dets*(22+K)) fp32 records in the JSON layout -> list of dicts (reference schema and order: images in
    batch order, detections in descending score; rows with category -1 dropped)."""
    K, D = int(num_classes), int(max_dets)
    w = 22 + K
    records = np.asarray(records, dtype=np.float32)
    out = []
    for b, img_id in enumerate(image_ids):
        n = int(records[b, 0])
        if n == 0:
            continue
        rows = records[b, 1:1 + D * w].reshape(D, w)[:n]
        cats = rows[:, 5].astype(np.int64)
        keep = np.nonzero(cats != -1)[0]
        if keep.size == 0:
            continue
        rows = rows[keep]
        boxes = rows[:, 0:4].tolist()
        scores = rows[:, 4].tolist()
        probs = rows[:, 6:6 + K].tolist()
        covs = rows[:, 6 + K:].reshape(-1, 4, 4).tolist()
        cat_list = cats[keep].tolist()
        for k in range(len(cat_list)):
            out.append({"image_id": img_id, "category_id": cat_list[k], "bbox": boxes[k], "score": scores[k],
                        "cls_prob": probs[k], "bbox_covar": covs[k]})
    return out


def det_from_instances(instances_list, num_classes, max_dets, device):
    """list of Instances (one per image) -> the engine's `det` dict (padded to max_dets), so that results which already
    left the engine can go through the same writer."""
    B = len(instances_list)
    det = {"boxes": torch.zeros((B, max_dets, 4), dtype=torch.float32, device=device),
           "cov": torch.zeros((B, max_dets, 4, 4), dtype=torch.float32, device=device),
           "scores": torch.zeros((B, max_dets), dtype=torch.float32, device=device),
           "classes": torch.zeros((B, max_dets), dtype=torch.int32, device=device),
           "probs": torch.zeros((B, max_dets, num_classes), dtype=torch.float32, device=device),
           "count": torch.zeros((B,), dtype=torch.int32, device=device)}
    for b, inst in enumerate(instances_list):
        n = len(inst)
        if n == 0:
            continue
        det["boxes"][b, :n] = inst.pred_boxes.tensor.to(device)
        det["scores"][b, :n] = inst.scores.to(device)
        det["classes"][b, :n] = inst.pred_classes.to(device=device, dtype=torch.int32)
        det["probs"][b, :n] = inst.pred_cls_probs.to(device)
        if inst.has("pred_boxes_covariance"):
            det["cov"][b, :n] = inst.pred_boxes_covariance.to(device)
        det["count"][b] = n
    return det
