"""Configuration surface of the probabilistic-inference path.

Mirrors the keys the reference reads on this path so the reference's own YAML
files load unchanged:
  * project keys   : reference src/core/setup.py:79-133 (add_probabilistic_config)
  * two-stage merge: reference src/core/setup.py:156,166 (model YAML, then inference YAML)
  * detectron2 defaults the predictor/model read (un-vendored dependency, values from
    detectron2/config/defaults.py of the v0.2-v0.3 era): TOPK_CANDIDATES_TEST=1000,
    SCORE_THRESH_TEST=0.05, NMS_THRESH_TEST=0.5, TEST.DETECTIONS_PER_IMAGE=100,
    NUM_CONVS=4, PRIOR_PROB=0.01, BBOX_REG_WEIGHTS=(1,1,1,1), anchor aspect ratios
    (0.5,1,2), FPN strides 8..128, anchor offset 0.
yacs/detectron2 are not available here, so `CfgNode` is a small attribute dict with
the subset of the yacs API the reference touches (clone, defrost, freeze,
merge_from_file, merge_from_list, item access).
"""
import ast
import copy
import os

import yaml


def _decode(value):
    """yacs `_decode_cfg_value`: strings that are Python literals (the reference YAMLs write tuples as
    `("bdd_train",)` or `(60000, 80000)`) become the literal; anything else stays a string."""
    if not isinstance(value, str):
        return value
    try:
        return ast.literal_eval(value)
    except (ValueError, SyntaxError):
        return value


class CfgNode(dict):
    """Attribute-style nested dict (subset of yacs.config.CfgNode)."""

    def __init__(self, init=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen", False):
            raise AttributeError("Attempted to set {} on a frozen CfgNode".format(name))
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            out[k] = copy.deepcopy(v, memo)
        out.__dict__["_frozen"] = self.__dict__.get("_frozen", False)
        return out

    def _set_frozen(self, flag):
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return self.__dict__.get("_frozen", False)

    def merge_from_other_cfg(self, other):
        _merge(self, other, [])

    def merge_from_file(self, path):
        self.merge_from_other_cfg(load_yaml_with_base(path))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if parts[-1] not in node:
                raise KeyError("Non-existent config key: {}".format(key))
            if isinstance(val, str):
                try:
                    val = yaml.safe_load(val)
                except yaml.YAMLError:
                    pass
            node[parts[-1]] = val


def _merge(dst, src, path):
    for k, v in src.items():
        if isinstance(v, dict):
            if k not in dst:
                dst[k] = CfgNode()
            if not isinstance(dst[k], dict):
                raise KeyError("Config key {} is not a node".format(".".join(path + [k])))
            _merge(dst[k], v, path + [k])
        else:
            v = _decode(v)
            if isinstance(v, (list, tuple)):
                v = type(dst.get(k, v))(v) if isinstance(dst.get(k, None), (list, tuple)) else v
            dst[k] = v


class _Loader(yaml.SafeLoader):
    pass


def _construct_eval(loader, node):
    # reference src/configs/Base-RetinaNet.yaml:8 encodes the anchor sizes as
    # `!!python/object/apply:eval ["[[x, x * 2**(1.0/3), ...] for x in [...]]"]`
    args = loader.construct_sequence(node, deep=True)
    if len(args) != 1 or not isinstance(args[0], str):
        raise yaml.YAMLError("unsupported eval payload in config")
    return _safe_arith_eval(args[0])


_ARITH_BIN = {ast.Add: lambda a, b: a + b, ast.Sub: lambda a, b: a - b, ast.Mult: lambda a, b: a * b,
              ast.Div: lambda a, b: a / b, ast.Pow: lambda a, b: a ** b, ast.FloorDiv: lambda a, b: a // b}


def _safe_arith_eval(src):
    """Evaluates the ONE construct the reference's YAML uses -- (nested) list displays / list comprehensions over
    literal lists whose elements are +,-,*,/,//,** arithmetic on numbers and the comprehension variables -- by walking
    the AST with a whitelist.  No names other than comprehension targets, no calls, attributes or subscripts: a
    config file shipped next to a model cannot execute code (python's eval() with emptied builtins can be escaped)."""
    tree = ast.parse(src, mode="eval")

    def ev(n, env):
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)) and not isinstance(n.value, bool):
            return n.value
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id in env:
            return env[n.id]
        if isinstance(n, ast.BinOp) and type(n.op) in _ARITH_BIN:
            a, b = ev(n.left, env), ev(n.right, env)
            if isinstance(n.op, ast.Pow) and abs(b) > 64:
                raise yaml.YAMLError("exponent too large in config expression")
            return _ARITH_BIN[type(n.op)](a, b)
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            v = ev(n.operand, env)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, (ast.List, ast.Tuple)):
            if len(n.elts) > 4096:
                raise yaml.YAMLError("list too long in config expression")
            return [ev(e, env) for e in n.elts]
        if isinstance(n, ast.ListComp) and len(n.generators) == 1:
            g = n.generators[0]
            if g.ifs or g.is_async or not isinstance(g.target, ast.Name):
                raise yaml.YAMLError("unsupported comprehension in config expression")
            seq = ev(g.iter, env)
            if not isinstance(seq, list):
                raise yaml.YAMLError("comprehension must iterate over a list literal")
            return [ev(n.elt, dict(env, **{g.target.id: x})) for x in seq]
        raise yaml.YAMLError("unsupported syntax in config expression: %s" % type(n).__name__)

    return ev(tree.body, {})


_Loader.add_constructor("tag:yaml.org,2002:python/object/apply:eval", _construct_eval)


def load_yaml_with_base(path):
    """Load a YAML file honouring detectron2's `_BASE_:` inheritance."""
    with open(path, "r") as f:
        cfg = yaml.load(f, Loader=_Loader) or {}
    base = cfg.pop("_BASE_", None)
    if base is not None:
        if not os.path.isabs(base):
            base = os.path.join(os.path.dirname(path), base)
        merged = load_yaml_with_base(base)
        _merge_plain(merged, cfg)
        return merged
    return cfg


def _merge_plain(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_plain(dst[k], v)
        else:
            dst[k] = v


def get_cfg():
    """Defaults: the detectron2 keys this path reads + the project keys
    (reference src/core/setup.py:79-133)."""
    C = CfgNode()
    C.VERSION = 2
    C.SEED = -1
    C.OUTPUT_DIR = "./output"
    C.MODEL = CfgNode()
    C.MODEL.META_ARCHITECTURE = "ProbabilisticRetinaNet"
    C.MODEL.DEVICE = "cuda"
    C.MODEL.WEIGHTS = ""
    C.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]
    C.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]
    C.MODEL.BACKBONE = CfgNode({"NAME": "build_retinanet_resnet_fpn_backbone", "FREEZE_AT": 2})
    C.MODEL.RESNETS = CfgNode({"DEPTH": 50, "OUT_FEATURES": ["res3", "res4", "res5"]})
    C.MODEL.FPN = CfgNode({"IN_FEATURES": ["res3", "res4", "res5"], "OUT_CHANNELS": 256})
    C.MODEL.ANCHOR_GENERATOR = CfgNode()
    C.MODEL.ANCHOR_GENERATOR.NAME = "DefaultAnchorGenerator"
    C.MODEL.ANCHOR_GENERATOR.SIZES = [[x, x * 2 ** (1.0 / 3), x * 2 ** (2.0 / 3)]
                                      for x in [32, 64, 128, 256, 512]]
    C.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS = [[0.5, 1.0, 2.0]]
    C.MODEL.ANCHOR_GENERATOR.OFFSET = 0.0
    C.MODEL.RPN = CfgNode({"BBOX_REG_WEIGHTS": (1.0, 1.0, 1.0, 1.0)})
    C.MODEL.ROI_BOX_HEAD = CfgNode({"DROPOUT_RATE": 0.0})
    C.MODEL.RETINANET = CfgNode()
    C.MODEL.RETINANET.NUM_CLASSES = 80
    C.MODEL.RETINANET.IN_FEATURES = ["p3", "p4", "p5", "p6", "p7"]
    C.MODEL.RETINANET.NUM_CONVS = 4
    C.MODEL.RETINANET.IOU_THRESHOLDS = [0.4, 0.5]
    C.MODEL.RETINANET.IOU_LABELS = [0, -1, 1]
    C.MODEL.RETINANET.PRIOR_PROB = 0.01
    C.MODEL.RETINANET.SCORE_THRESH_TEST = 0.05
    C.MODEL.RETINANET.TOPK_CANDIDATES_TEST = 1000
    C.MODEL.RETINANET.NMS_THRESH_TEST = 0.5
    C.MODEL.RETINANET.BBOX_REG_WEIGHTS = (1.0, 1.0, 1.0, 1.0)
    C.MODEL.RETINANET.FOCAL_LOSS_GAMMA = 2.0
    C.MODEL.RETINANET.FOCAL_LOSS_ALPHA = 0.25
    C.MODEL.RETINANET.SMOOTH_L1_LOSS_BETA = 0.1
    C.INPUT = CfgNode({"MIN_SIZE_TRAIN": (800,), "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333,
                       "FORMAT": "BGR"})
    C.DATASETS = CfgNode({"TRAIN": (), "TEST": ()})
    C.DATALOADER = CfgNode({"NUM_WORKERS": 4})
    C.SOLVER = CfgNode({"IMS_PER_BATCH": 16, "BASE_LR": 0.001, "STEPS": (30000, 40000),
                        "MAX_ITER": 40000, "CHECKPOINT_PERIOD": 5000})
    C.TEST = CfgNode({"DETECTIONS_PER_IMAGE": 100})
    add_probabilistic_config(C)
    return C


def add_probabilistic_config(cfg):
    """Project keys, names and defaults as in reference src/core/setup.py:79-133."""
    _C = cfg
    _C.MODEL.PROBABILISTIC_MODELING = CfgNode()
    _C.MODEL.PROBABILISTIC_MODELING.MC_DROPOUT = CfgNode()
    _C.MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS = CfgNode()
    _C.MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS = CfgNode()
    _C.MODEL.PROBABILISTIC_MODELING.ANNEALING_STEP = 0
    _C.MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE = 0.0
    _C.MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NAME = "none"
    _C.MODEL.PROBABILISTIC_MODELING.CLS_VAR_LOSS.NUM_SAMPLES = 3
    _C.MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.NAME = "none"
    _C.MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.COVARIANCE_TYPE = "diagonal"
    _C.MODEL.PROBABILISTIC_MODELING.BBOX_COV_LOSS.NUM_SAMPLES = 1000
    _C.PROBABILISTIC_INFERENCE = CfgNode()
    _C.PROBABILISTIC_INFERENCE.MC_DROPOUT = CfgNode()
    _C.PROBABILISTIC_INFERENCE.BAYES_OD = CfgNode()
    _C.PROBABILISTIC_INFERENCE.ENSEMBLES_DROPOUT = CfgNode()
    _C.PROBABILISTIC_INFERENCE.ENSEMBLES = CfgNode()
    _C.PROBABILISTIC_INFERENCE.INFERENCE_MODE = "standard_nms"
    _C.PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE = False
    _C.PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS = 1
    _C.PROBABILISTIC_INFERENCE.AFFINITY_THRESHOLD = 0.7
    _C.PROBABILISTIC_INFERENCE.BAYES_OD.BOX_MERGE_MODE = "bayesian_inference"
    _C.PROBABILISTIC_INFERENCE.BAYES_OD.CLS_MERGE_MODE = "bayesian_inference"
    _C.PROBABILISTIC_INFERENCE.BAYES_OD.DIRCH_PRIOR = "uniform"
    _C.PROBABILISTIC_INFERENCE.ENSEMBLES_DROPOUT.BOX_MERGE_MODE = "pre_nms"
    _C.PROBABILISTIC_INFERENCE.ENSEMBLES.BOX_MERGE_MODE = "pre_nms"
    _C.PROBABILISTIC_INFERENCE.ENSEMBLES.RANDOM_SEED_NUMS = [0, 1000, 2000, 3000, 4000]


def setup_config(config_file, inference_config="", opts=None, random_seed=0, output_dir=None):
    """Two-stage merge as reference src/core/setup.py:136-212 (without the dataset /
    logger / output-directory side effects, which belong to the harness)."""
    cfg = get_cfg()
    cfg.merge_from_file(config_file)
    cfg.MODEL.ROI_BOX_HEAD.DROPOUT_RATE = cfg.MODEL.PROBABILISTIC_MODELING.DROPOUT_RATE
    if inference_config:
        cfg.merge_from_file(inference_config)
    if opts:
        cfg.merge_from_list(list(opts))
    if output_dir is not None:
        cfg.OUTPUT_DIR = output_dir
    cfg.SEED = random_seed
    cfg.freeze()
    return cfg
