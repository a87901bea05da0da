"""Wire format of the path: Instances -> COCO-style result dicts (the file the reference's offline
evaluation reads, reference src/probabilistic_inference/inference_utils.py:428-502 and
src/apply_net.py:91-102).  Host-side: at most 100 detections per image."""
import torch

_T_XYXY_TO_XYWH = [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [0.0, -1.0, 0.0, 1.0]]


def covar_xyxy_to_xywh(output_boxes_covariance):
    """Sigma -> T Sigma T^T with T mapping (x1,y1,x2,y2) to (x,y,w,h)   (inference_utils.py:428-451)."""
    cov = output_boxes_covariance
    T = torch.as_tensor(_T_XYXY_TO_XYWH, dtype=cov.dtype, device=cov.device).unsqueeze(0)
    T = torch.repeat_interleave(T, cov.shape[0], 0)
    return torch.matmul(torch.matmul(T, cov), torch.transpose(T, 2, 1))


def instances_to_json(instances, img_id, cat_mapping_dict=None):
    """Same schema and filtering as the reference (inference_utils.py:454-502): image_id, category_id,
    bbox (XYWH), score, cls_prob, bbox_covar; detections whose class has no dataset id are dropped."""
    num_instance = len(instances)
    if num_instance == 0:
        return []
    boxes = instances.pred_boxes.tensor.detach().cpu().clone()
    boxes[:, 2] -= boxes[:, 0]
    boxes[:, 3] -= boxes[:, 1]
    boxes = boxes.tolist()
    scores = instances.scores.cpu().tolist()
    classes = instances.pred_classes.cpu().tolist()
    if cat_mapping_dict is not None:
        classes = [cat_mapping_dict[c] if c in cat_mapping_dict.keys() else -1 for c in classes]
    pred_cls_probs = instances.pred_cls_probs.cpu().tolist()
    if instances.has("pred_boxes_covariance"):
        pred_boxes_covariance = covar_xyxy_to_xywh(instances.pred_boxes_covariance).cpu().tolist()
    else:
        pred_boxes_covariance = []
    results = []
    for k in range(num_instance):
        if classes[k] != -1:
            results.append({"image_id": img_id, "category_id": classes[k], "bbox": boxes[k], "score": scores[k],
                            "cls_prob": pred_cls_probs[k],
                            "bbox_covar": pred_boxes_covariance[k] if pred_boxes_covariance else []})
    return results
