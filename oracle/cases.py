"""TEST INFRASTRUCTURE -- the parity cases shared by oracle/make_golden.py and tests/.

Each case names a model variant + inference mode the reference ships
(src/configs/BDD-Detection/retinanet/*.yaml x src/configs/Inference/*.yaml) at a geometry small
enough for the CPU suite; cfg objects are assembled from pod_compare_b200.config defaults with the
same key/value pairs those YAML files set, so nothing under /root/reference is read at test time.
"""
from pod_compare_b200.config import get_cfg

_VAR = ["MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NAME", "loss_attenuation",
        "MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NUM_SAMPLES", 10,
        "MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.NAME", "negative_log_likelihood",
        "MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.COVARIANCE_TYPE", "diagonal"]
_FULL = ["MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.COVARIANCE_TYPE", "full"]
_DROP = ["MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE", 0.2]


def _mc(n):
    return ["PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE", True, "PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS", n]


def _mode(m):
    return ["PROBABILISTIC_INFERENCE.INFERENCE_MODE", m, "PROBABILISTIC_INFERENCE.AFFINITY_THRESHOLD", 0.9]


_BOD = ["PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE", "max_score",
        "PROBABILISTIC_INFERENCE.BAYES_OD.BOX_MERGE_MODE", "bayesian_inference"]

# name -> (opts, mode, n_mc, member seeds, image_hw, out_hw, rng seed, image idx)
CASES = {
    "baseline_std": ([] + _mode("standard_nms"), "standard_nms", 1, [0], (96, 160), (96, 160), 11, 0),
    "regclsvar_std": (_VAR + _mode("standard_nms"), "standard_nms", 1, [0], (96, 160), (120, 200), 12, 1),
    "mcdrop_pre_n4": (_VAR + _DROP + _mode("mc_dropout_ensembles") + _mc(4), "mc_dropout_ensembles", 4, [0],
                      (96, 160), (96, 160), 13, 2),
    "droponly_pre_n3": (_DROP + _mode("mc_dropout_ensembles") + _mc(3), "mc_dropout_ensembles", 3, [1000],
                        (128, 128), (128, 128), 14, 3),
    "bayesod_mc_n3": (_VAR + _DROP + _mode("bayes_od") + _BOD + _mc(3), "bayes_od", 3, [0], (96, 160), (96, 160), 15, 4),
    "bayesod_plain": (_VAR + _mode("bayes_od") + _BOD, "bayes_od", 1, [2000], (96, 160), (48, 80), 16, 5),
    "bayesod_clsavg_ci": (_VAR + _mode("bayes_od") + ["PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE",
                          "bayesian_inference", "PROBABILISTIC_INFERENCE.BAYES_OD.BOX_MERGE_MODE",
                          "covariance_intersection"], "bayes_od", 1, [0], (96, 160), (96, 160), 17, 6),
    "ensembles_e3": (_VAR + _mode("ensembles") + ["PROBABILISTIC_INFERENCE.ENSEMBLES.RANDOM_SEED_NUMS", [0, 1000, 2000]],
                     "ensembles", 1, [0, 1000, 2000], (96, 160), (96, 160), 18, 7),
    "anchorstats_var": (_VAR + _mode("anchor_statistics"), "anchor_statistics", 1, [0], (96, 160), (96, 160), 20, 9),
    "anchorstats_base": ([] + _mode("anchor_statistics"), "anchor_statistics", 1, [1000], (96, 160), (96, 160), 21, 10),
    "mcdrop_post_n3": (_VAR + _DROP + _mode("mc_dropout_ensembles") + _mc(3) +
                       ["PROBABILISTIC_INFERENCE.ENSEMBLES_DROPOUT.BOX_MERGE_MODE", "post_nms"], "mc_dropout_ensembles", 3, [0],
                       (96, 160), (96, 160), 22, 11),
    "ensembles_post_e3": (_VAR + _mode("ensembles") + ["PROBABILISTIC_INFERENCE.ENSEMBLES.RANDOM_SEED_NUMS", [0, 1000, 2000],
                          "PROBABILISTIC_INFERENCE.ENSEMBLES.BOX_MERGE_MODE", "post_nms"], "ensembles", 1, [0, 1000, 2000],
                          (96, 160), (96, 160), 23, 12),
    "coco80_std": (_VAR + _mode("standard_nms") + ["MODEL.RETINANET.NUM_CLASSES", 80], "standard_nms", 1, [0],
                   (64, 64), (64, 64), 24, 13),
    # MC_DROPOUT.ENABLE with NUM_RUNS == 1: the reference keeps the model in train() (:52-56), so the single forward
    # has active dropout (mean / variance heads from independently masked tower passes) and no epistemic term
    "mcdrop_single": (_VAR + _DROP + _mode("standard_nms") + _mc(1), "standard_nms", 1, [2000], (96, 160), (96, 160), 25, 14),
    # the sampled (aleatoric) decode uses MODEL.RPN.BBOX_REG_WEIGHTS, the deterministic one MODEL.RETINANET.BBOX_REG_WEIGHTS
    "regclsvar_rpnw": (_VAR + _mode("standard_nms") + ["MODEL.RPN.BBOX_REG_WEIGHTS", (2.0, 2.0, 1.5, 1.5)], "standard_nms", 1,
                       [1000], (96, 160), (96, 160), 26, 15),
    "fullcov_mc_n3": (_VAR + _FULL + _DROP + _mode("mc_dropout_ensembles") + _mc(3), "mc_dropout_ensembles", 3, [3000],
                      (96, 160), (96, 160), 29, 8),   # seed chosen so that no two candidates tie exactly (torch.topk leaves
                                                      # the order of equal scores unspecified; the oracle defines lower index first)
}


def case_features(name):
    """Synthetic FPN maps of a case.  Ensemble cases get ONE FEATURE SET PER MEMBER (correlated but different, synthetic.make_member_features):
    each member of the reference is a full model with its own backbone (probabilistic_inference.py:58-77,499-501),
    so the members never see the same feature maps."""
    from pod_compare_b200 import synthetic as S
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = CASES[name]
    if mode == "ensembles":
        return S.make_member_features(len(seeds), img, hw[0], hw[1])
    return S.make_features(0, img, hw[0], hw[1])


def build_cfg(name):
    opts = CASES[name][0]
    cfg = get_cfg()
    cfg.MODEL.RETINANET.NUM_CLASSES = 7        # BDD (reference Base-BDD-RetinaNet.yaml:12) unless the case overrides it
    cfg.merge_from_list(list(opts))
    cfg.MODEL.DEVICE = "cpu"
    cfg.freeze()
    return cfg


def is_mc_single(name):
    """MC-dropout enabled with a single run: dropout active in the one forward, no sample aggregation."""
    opts, mode, n_mc = CASES[name][0], CASES[name][1], CASES[name][2]
    enabled = any(isinstance(o, str) and o == "PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE" for o in opts)
    has_drop = any(isinstance(o, str) and o == "MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE" for o in opts)
    return enabled and has_drop and n_mc == 1 and mode != "ensembles"


def is_post_nms(name):
    opts = CASES[name][0]
    return any(isinstance(o, str) and o == "post_nms" for o in opts)
