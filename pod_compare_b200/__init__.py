"""pod_compare_b200 -- B200-native probabilistic-inference path of asharakeh/pod_compare.

Public surface (mirrors the reference plugin API for this path):
    build_predictor(cfg)            reference src/probabilistic_inference/probabilistic_inference.py:20-33
    setup_config(...), get_cfg()    reference src/core/setup.py:79-212 (keys on this path)
    predictor.predict_batch_json    reference src/apply_net.py:88-98 (predict + instances_to_json for a batch)
    inference_utils.instances_to_json / covar_xyxy_to_xywh   reference src/probabilistic_inference/inference_utils.py:428-502
    wire.build_category_mapping     reference src/apply_net.py:53-79
"""
from .config import CfgNode, get_cfg, setup_config  # noqa: F401


def build_predictor(cfg):
    from .predictor import build_predictor as _bp
    return _bp(cfg)
