"""Reference-named entry points of the wire format (src/probabilistic_inference/inference_utils.py:428-502), kept so
that `from probabilistic_inference.inference_utils import instances_to_json` in the reference's harness can be pointed
at this package.  Both go through the batched GPU writer of wire.py (one kernel + one device->host copy); the harness
loop of src/apply_net.py:88-98 should prefer `predictor.predict_batch_json`, which never leaves the batch layout."""
import torch

from . import ops, wire


def covar_xyxy_to_xywh(output_boxes_covariance):
    """(n,4,4) xyxy covariances -> T Sigma T^T in the XYWH parametrisation (inference_utils.py:428-451)."""
    cov = torch.as_tensor(output_boxes_covariance, dtype=torch.float32)
    n = int(cov.shape[0])
    if n == 0:
        return cov.reshape(0, 4, 4)
    dev = cov.device if cov.is_cuda else torch.device("cuda", torch.cuda.current_device())
    det = {"boxes": torch.zeros((1, n, 4), device=dev), "cov": cov.to(dev).reshape(1, n, 4, 4).contiguous(),
           "scores": torch.zeros((1, n), device=dev), "classes": torch.zeros((1, n), dtype=torch.int32, device=dev),
           "probs": torch.zeros((1, n, 1), device=dev), "count": torch.full((1,), n, dtype=torch.int32, device=dev)}
    rec = ops.wire_records(det, xywh=True)
    return rec[0, 1:].reshape(n, 23)[:, 7:].reshape(n, 4, 4).to(cov.device)


def instances_to_json(instances, img_id, cat_mapping_dict=None):
    """One image: Instances -> list of result dicts with the reference's schema and filtering (:454-502)."""
    n = len(instances)
    if n == 0:
        return []
    K = int(instances.pred_cls_probs.shape[1])
    dev = instances.scores.device if instances.scores.is_cuda else torch.device("cuda", torch.cuda.current_device())
    writer = wire.BatchJsonWriter(K, n, cat_mapping_dict, dev)
    res = writer.to_json(wire.det_from_instances([instances], K, n, dev), [img_id])
    if not instances.has("pred_boxes_covariance"):
        for r in res:
            r["bbox_covar"] = []
    return res
