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
    "fullcov_mc_n3": (_VAR + _FULL + _DROP + _mode("mc_dropout_ensembles") + _mc(3), "mc_dropout_ensembles", 3, [3000],
                      (96, 160), (96, 160), 19, 8),
}


def build_cfg(name):
    opts = CASES[name][0]
    cfg = get_cfg()
    cfg.MODEL.RETINANET.NUM_CLASSES = 7        # BDD (reference Base-BDD-RetinaNet.yaml:12) unless the case overrides it
    cfg.merge_from_list(list(opts))
    cfg.MODEL.DEVICE = "cpu"
    cfg.freeze()
    return cfg


def is_post_nms(name):
    opts = CASES[name][0]
    return any(isinstance(o, str) and o == "post_nms" for o in opts)
