"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden [case names ... | planted]

For every case of oracle/cases.py the reference's own `build_predictor(cfg)` /
`predictor(input_im)` (src/probabilistic_inference/probabilistic_inference.py:20-111) is run on
seeded synthetic features and weights, with randomness injected from oracle/philox.py, and the
results are stored:  final Instances fields and the anchor-wise 5-tuple of
`retinanet_probabilistic_inference` (:178-388).  Stage-isolated fixtures feed planted candidate
sets through the reference's `general_standard_nms_postprocessing` (inference_utils.py:12-54),
`post_processing_bayes_od` (:536-636) and `bounding_box_bayesian_inference`
(inference_utils.py:292-334).  The fixtures pin oracle/podref.py (tests/test_oracle_golden.py)
and, through it, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

from oracle import cases as C
from oracle import podref as O
from oracle import ref_runner as R
from pod_compare_b200 import synthetic as S

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def feats_checksum(feats):
    if isinstance(feats[0], (list, tuple)):
        feats = [f for fs in feats for f in fs]
    return np.array([float(f.double().abs().sum()) for f in feats])


def state_dicts_for(name):
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg = C.build_cfg(name)
    pp = O.PathParams.from_cfg(cfg)
    sds = [S.make_head_state_dict(s, num_classes=pp.num_classes, use_dropout=pp.use_dropout,
                                  cls_var=pp.cls_var, bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims) for s in seeds]
    return cfg, pp, sds


def model_case(name):
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds = state_dicts_for(name)
    pred = R.build_reference_predictor(cfg, sds if len(sds) > 1 else sds[0])
    feats = C.case_features(name)
    final, _ = R.run_reference(pred, feats, hw, out_hw=out_hw, seed=seed, image_idx=img, stage="final")
    if C.is_post_nms(name):
        d = {"final_" + k: v for k, v in R.instances_to_arrays(final).items()}
        d["feats_checksum"] = feats_checksum(feats)
        np.savez_compressed(os.path.join(OUT, "case_%s.npz" % name), **d)
        print("case %-20s detections=%3d (post-NMS merge)" % (name, len(final)))
        return
    aw, rng = R.run_reference(pred, feats, hw, out_hw=out_hw, seed=seed, image_idx=img, stage="anchorwise")
    boxes, cov, prob, cls, vec = aw
    d = {"final_" + k: v for k, v in R.instances_to_arrays(final).items()}
    d["cand_boxes"] = boxes.numpy()
    d["cand_has_cov"] = np.array(isinstance(cov, torch.Tensor))
    d["cand_cov"] = cov.numpy() if isinstance(cov, torch.Tensor) else np.zeros((0, 4, 4), np.float32)
    d["cand_scores"] = prob.numpy()
    d["cand_classes"] = cls.numpy()
    d["cand_probs"] = vec.numpy()
    d["cand_anchor_ids"] = np.asarray(rng.cand_ids, dtype=np.int64)
    d["feats_checksum"] = feats_checksum(feats)
    np.savez_compressed(os.path.join(OUT, "case_%s.npz" % name), **d)
    print("case %-20s detections=%3d candidates=%4d" % (name, len(final), boxes.shape[0]))


JSON_CASES = ("regclsvar_std", "bayesod_plain", "baseline_std", "mcdrop_pre_n4")


def json_case(name):
    """The reference's own wire format: `instances_to_json` (inference_utils.py:454-502) applied to the final Instances
    of a case, with the two category mappings src/apply_net.py:53-79 can produce for a BDD-trained model (BDD test
    set: contiguous id + 1; KITTI test set: only car / person survive).  Stored as JSON text, like the file the
    reference writes (src/apply_net.py:100-102)."""
    import json
    from pod_compare_b200 import wire
    IU = R.load_reference()["inference_utils"]
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds = state_dicts_for(name)
    pred = R.build_reference_predictor(cfg, sds if len(sds) > 1 else sds[0])
    final, _ = R.run_reference(pred, C.case_features(name), hw, out_hw=out_hw, seed=seed, image_idx=img, stage="final")
    bdd, kitti = wire.BDD_THING_DATASET_ID_TO_CONTIGUOUS_ID, wire.KITTI_THING_DATASET_ID_TO_CONTIGUOUS_ID
    maps = {"bdd": wire.build_category_mapping("bdd_train", "bdd_val", bdd, bdd),
            "kitti": wire.build_category_mapping("bdd_train", "kitti_val", bdd, kitti)}
    out = {k: IU.instances_to_json(final, 1000 + img, m) for k, m in maps.items()}
    with open(os.path.join(OUT, "json_%s.json" % name), "w") as f:
        json.dump(out, f, indent=1, separators=(",", ": "))
    print("json %-20s bdd=%d kitti=%d entries" % (name, len(out["bdd"]), len(out["kitti"])))


def planted_cases():
    mods = R.load_reference()
    IU, PI = mods["inference_utils"], mods["inference"]
    from detectron2.structures import Instances  # shim
    cfg = C.build_cfg("bayesod_plain")
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(0, use_dropout=False, cls_var=True, bbox_cov=True)
    input_im = [{"image": torch.zeros((3, 720, 1280), dtype=torch.uint8), "height": 720, "width": 1280}]
    for tag, num_gt, per_gt in (("small", 12, (1, 30)), ("large", 40, (10, 50))):
        boxes, cov, scores, classes, probs = S.make_planted_candidates(7 if tag == "small" else 8, num_gt, per_gt)
        tup = (boxes, cov, scores, classes, probs)
        res = IU.general_standard_nms_postprocessing(input_im, tup, 0.5, 100)
        d = {"in_boxes": boxes.numpy(), "in_cov": cov.numpy(), "in_scores": scores.numpy(),
             "in_classes": classes.numpy(), "in_probs": probs.numpy()}
        d.update({"std_" + k: v for k, v in R.instances_to_arrays(res).items()})
        for cm in ("max_score", "bayesian_inference"):
            for bm in ("bayesian_inference", "covariance_intersection"):
                c2 = cfg.clone()
                c2.defrost()
                c2.PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE = cm
                c2.PROBABILISTIC_INFERENCE.BAYES_OD.BOX_MERGE_MODE = bm
                c2.freeze()
                pred = R.build_reference_predictor(c2, sd)
                pred.retinanet_probabilistic_inference = lambda im, t=tup: t
                out = pred.post_processing_bayes_od(input_im)
                post = IU.probabilistic_detector_postprocess(out, 720, 1280)
                key = "bod_%s_%s_" % ("ms" if cm == "max_score" else "avg", "bi" if bm == "bayesian_inference" else "ci")
                d.update({key + k: v for k, v in R.instances_to_arrays(post).items()})
        for use_cov, key in ((True, "ast_cov_"), (False, "ast_nocov_")):
            t2 = (boxes, cov if use_cov else [], scores, classes, probs)
            out = IU.general_anchor_statistics_postprocessing(input_im, t2, 0.5, 100, 0.9)
            post = IU.probabilistic_detector_postprocess(out, 720, 1280)
            d.update({key + k: v for k, v in R.instances_to_arrays(post).items()})
        np.savez_compressed(os.path.join(OUT, "planted_%s.npz" % tag), **d)
        print("planted %-6s M=%d std=%d" % (tag, boxes.shape[0], len(res)))


def main():
    if not R.reference_available():
        sys.exit("reference tree not available; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    for name in (only or C.CASES):
        if name not in ("planted", "json"):
            model_case(name)
    if not only or "planted" in only:
        planted_cases()
    if not only or "json" in only:
        for name in JSON_CASES:
            json_case(name)


if __name__ == "__main__":
    main()
