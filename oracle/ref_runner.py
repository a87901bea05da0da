"""TEST INFRASTRUCTURE (oracle) -- runs the UNMODIFIED reference predictor on CPU.

Usable where /root/reference exists (the build container) or where oracle/make_ref.py staged the reference's
four hot-path files under oracle/_ref (git-ignored; travels to the GPU box).  It puts the shim
packages of oracle/ref_shim (detectron2 / fvcore stand-ins) and /root/reference/src on
sys.path, imports the reference's own modules

    probabilistic_inference.probabilistic_inference   (build_predictor, RetinaNetProbabilisticPredictor)
    probabilistic_inference.inference_utils
    probabilistic_modeling.probabilistic_retinanet    (ProbabilisticRetinaNet + Head)

and drives `build_predictor(cfg)` / `predictor(input_im)` exactly as src/apply_net.py:82-91
does.  Randomness is injected from the counter-based streams of oracle/philox.py by
patching the three places the reference draws from torch's generator:
    torch.nn.functional.dropout                         (nn.Dropout in the towers)
    torch.distributions.normal._standard_normal         (logit sampling)
    torch.distributions.multivariate_normal._standard_normal   (box-delta sampling)
Used by oracle/make_golden.py (golden fixtures) and by tests that cross-check the
restatement oracle/podref.py against the live reference.
"""
import contextlib
import os
import sys

import numpy as np
import torch

from oracle import philox

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "ref_shim")


def _reference_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte staged copy of the reference's four
    hot-path files under oracle/_ref (oracle/make_ref.py)."""
    for root in (os.environ.get("POD_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if root and os.path.isfile(os.path.join(root, "src", "probabilistic_inference", "probabilistic_inference.py")):
            return root
    return None


REFERENCE_ROOT = _reference_root()


def reference_available():
    return REFERENCE_ROOT is not None


_mods = {}


def load_reference():
    """Import the reference modules (once). Returns a dict of modules."""
    if _mods:
        return _mods
    if not reference_available():
        raise RuntimeError("reference sources not present (neither /root/reference nor oracle/_ref)")
    src = os.path.join(REFERENCE_ROOT, "src")
    for p in (src, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
    # shim first: its `core.visualization_tools` stub must shadow the matplotlib one
    sys.path.insert(0, src)
    sys.path.insert(0, _SHIM)
    import importlib
    _mods["retinanet"] = importlib.import_module("probabilistic_modeling.probabilistic_retinanet")
    _mods["modeling_utils"] = importlib.import_module("probabilistic_modeling.modeling_utils")
    _mods["inference_utils"] = importlib.import_module("probabilistic_inference.inference_utils")
    _mods["inference"] = importlib.import_module("probabilistic_inference.probabilistic_inference")
    # the reference picks its post-processing device at import time (inference_utils.py:9: cuda if available); the oracle
    # is the reference's CPU path, also on a machine that has a GPU
    _mods["inference_utils"].device = torch.device("cpu")
    return _mods


class RngInjector:
    """Counter-based randomness for one `predictor(input_im)` call of the reference."""

    def __init__(self, seed, image_idx, n_levels, per_entry_dropouts, num_classes, p):
        self.seed, self.image = seed, image_idx
        self.n_levels = n_levels
        self.per_entry = per_entry_dropouts          # 8 or 16 dropout calls per feature entry
        self.K = num_classes
        self.p = p
        self.n_dropout = 0
        self.n_normal = 0
        self.cand_ids = []                            # global anchor ids in concatenation order
        self._ids_start = 0
        self.level_sizes = None                       # anchors per level (set by the runner)
        self.log = {"dropout_calls": 0, "normal_calls": 0, "mvn_calls": 0}

    # --- nn.Dropout -> F.dropout(input, p, training, inplace) -----------------------------
    def dropout(self, input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        i = self.n_dropout
        self.n_dropout += 1
        self.log["dropout_calls"] += 1
        entry, r = divmod(i, self.per_entry)
        sample, level = divmod(entry, self.n_levels)
        pass_, r = divmod(r, 8)
        tower, layer = divmod(r, 4)
        n, C, H, W = input.shape
        assert n == 1
        keep = philox.dropout_keep_mask(self.seed, self.image, sample, pass_, tower, layer,
                                        level, H, W, C, p)
        keep = torch.from_numpy(np.ascontiguousarray(keep.transpose(2, 0, 1)))[None]
        scale = torch.tensor(1.0, dtype=input.dtype) / torch.tensor(1.0 - p, dtype=input.dtype)
        return input * (keep.to(input.dtype) * scale)

    # --- Normal.rsample((S,)) -> _standard_normal((S, HWA, K)) ----------------------------
    def normal_eps(self, shape, dtype, device):
        S, n_anchor, K = shape
        run, level = divmod(self.n_normal, self.n_levels)      # one call per level per inference run
        self.n_normal += 1
        self.log["normal_calls"] += 1
        return torch.from_numpy(philox.logit_normals(self.seed, self.image, level, S, n_anchor, K, run=run)).to(dtype)

    # --- MultivariateNormal.rsample((1000,)) -> _standard_normal((1000, M, 4)) ------------
    def mvn_eps(self, shape, dtype, device):
        S, M, D = shape
        ids = self.cand_ids[self._ids_start:]                  # candidates gathered since the previous run
        assert D == 4 and M == len(ids), (shape, len(ids))
        self._ids_start = len(self.cand_ids)
        run = self.log["mvn_calls"]
        self.log["mvn_calls"] += 1
        return torch.from_numpy(philox.box_normals(self.seed, self.image, ids, S, run=run)).to(dtype)


@contextlib.contextmanager
def injected(rng):
    """Patch the reference's three randomness sources + capture candidate anchor ids."""
    mods = load_reference()
    PI = mods["inference"]
    import torch.distributions.normal as tdn
    import torch.distributions.multivariate_normal as tdm
    import torch.nn.functional as F
    saved = (F.dropout, tdn._standard_normal, tdm._standard_normal, PI.covariance_output_to_cholesky)
    real_chol = PI.covariance_output_to_cholesky

    def chol_spy(pred_bbox_cov):
        # called at probabilistic_inference.py:318 inside the per-level loop; the caller's
        # locals hold the level index `i` and the kept `anchor_idxs`.
        fr = sys._getframe(1).f_locals
        lvl = int(fr["i"])
        ids = fr["anchor_idxs"].cpu().numpy().astype(np.int64)
        off = int(np.sum(rng.level_sizes[:lvl]))
        rng.cand_ids.extend((ids + off).tolist())
        return real_chol(pred_bbox_cov)

    F.dropout = rng.dropout
    tdn._standard_normal = rng.normal_eps
    tdm._standard_normal = rng.mvn_eps
    PI.covariance_output_to_cholesky = chol_spy
    try:
        yield rng
    finally:
        F.dropout, tdn._standard_normal, tdm._standard_normal, PI.covariance_output_to_cholesky = saved


def build_reference_predictor(cfg, state_dicts):
    """`build_predictor(cfg)` of the reference with weight set(s) loaded.
    state_dicts: one dict, or a list of E dicts for INFERENCE_MODE == 'ensembles'."""
    mods = load_reference()
    cfg = cfg.clone()
    cfg.defrost()
    cfg.MODEL.DEVICE = "cpu"
    cfg.freeze()
    predictor = mods["inference"].build_predictor(cfg)
    if isinstance(state_dicts, dict):
        state_dicts = [state_dicts]
    missing, unexpected = predictor.model.load_state_dict(state_dicts[0], strict=False)
    assert not [k for k in missing if k.startswith("head.")], missing
    assert not unexpected, unexpected
    if predictor.model_list:
        assert len(predictor.model_list) == len(state_dicts)
        for m, sd in zip(predictor.model_list, state_dicts):
            m.load_state_dict(sd, strict=False)
    return predictor


def _set_features(predictor, feats):
    """feats: list over levels (shared), or -- ensembles -- a list over members of lists over levels: every member
    of the reference is a full model with its own backbone and therefore its own feature maps
    (probabilistic_inference.py:58-77,499-501)."""
    names = predictor.model.in_features
    per_member = isinstance(feats[0], (list, tuple))
    sets = feats if per_member else [feats]
    predictor.model.backbone.current = {n: f for n, f in zip(names, sets[0])}
    if per_member:
        assert len(sets) == len(predictor.model_list), (len(sets), len(predictor.model_list))
    for e, m in enumerate(predictor.model_list):
        m.backbone.current = {n: f for n, f in zip(names, sets[e if per_member else 0])}


def run_reference(predictor, feats, image_hw, out_hw=None, seed=1, image_idx=0, stage="final"):
    """One image through the reference.
    feats: list of 5 (1,C,H,W) tensors. image_hw: (H,W) of the (unpadded) network input.
    stage: 'final' -> Instances after probabilistic_detector_postprocess (== predictor(input_im));
           'anchorwise' -> the 5-tuple of retinanet_probabilistic_inference (standard/MC modes)."""
    cfg = predictor.cfg
    H, W = image_hw
    out_hw = out_hw or image_hw
    input_im = [{"image": torch.zeros((3, H, W), dtype=torch.uint8), "height": out_hw[0],
                 "width": out_hw[1], "image_id": image_idx}]
    _set_features(predictor, feats)
    if isinstance(feats[0], (list, tuple)):
        feats = feats[0]
    head = predictor.model.head
    per_entry = 8 + (4 if head.compute_cls_var else 0) + (4 if head.compute_bbox_cov else 0)
    A = 9
    rng = RngInjector(seed, image_idx, len(feats), per_entry, cfg.MODEL.RETINANET.NUM_CLASSES,
                      cfg.MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE)
    rng.level_sizes = [int(f.shape[-2] * f.shape[-1] * A) for f in feats]
    with injected(rng), torch.no_grad():
        if stage == "final":
            out = predictor(input_im)
        elif stage == "anchorwise":
            if predictor.inference_mode not in ("standard_nms", "mc_dropout_ensembles", "ensembles", "bayes_od",
                                                "anchor_statistics"):
                raise ValueError(predictor.inference_mode)
            if predictor.inference_mode == "ensembles":
                outs = [m(input_im, return_anchorwise_output=True) for m in predictor.model_list]
                out = predictor.retinanet_probabilistic_inference(
                    input_im, ensemble_inference=True, outputs_list=outs)
            else:
                out = predictor.retinanet_probabilistic_inference(input_im)
        else:
            raise ValueError(stage)
    return out, rng


def instances_to_arrays(inst):
    """Instances -> dict of numpy arrays (fixture format)."""
    n = len(inst)
    return {
        "boxes": inst.pred_boxes.tensor.detach().cpu().numpy().astype(np.float32).reshape(n, 4),
        "scores": inst.scores.detach().cpu().numpy().astype(np.float32),
        "classes": inst.pred_classes.detach().cpu().numpy().astype(np.int64),
        "probs": inst.pred_cls_probs.detach().cpu().numpy().astype(np.float32),
        "cov": inst.pred_boxes_covariance.detach().cpu().numpy().astype(np.float32).reshape(n, 4, 4),
    }
