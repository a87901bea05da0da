"""Reference-facing plugin surface: `build_predictor(cfg)` -> predictor with `__call__(input_im)`.

Same names, argument meaning and error behaviour as the reference
(src/probabilistic_inference/probabilistic_inference.py:20-33, 36-111, 169-176, 390-407, 430-443,
483-505, 536-636), with the arithmetic moved to libpodb200 and two batched entry points added for the
B200 path: `predict_batch(list_of_inputs)` and `infer_from_features(feats, ...)`.

Input convention: each `input_im` is the list-of-one-dict of the detectron2 data loader
({"image": (3,H,W) tensor, "height", "width", "image_id"}, src/apply_net.py:88-96).  The ResNet-FPN
backbone is upstream of the path rebuilt here: a dict may carry precomputed
"features" (list of 5 (1|B,256,Hl,Wl) tensors); otherwise `predictor.backbone` (any callable mapping
the preprocessed image batch to the 5 FPN maps) must be set.
"""
import os

import torch

from . import _cabi, engine, ops
from .engine import HeadEngine, HeadWeights, PathConfig
from .structures import Boxes, Instances

SUPPORTED_PRE_NMS_MODES = ("standard_nms", "mc_dropout_ensembles", "ensembles", "bayes_od", "anchor_statistics")


def build_predictor(cfg):
    """reference probabilistic_inference.py:20-33."""
    if cfg.MODEL.META_ARCHITECTURE == 'ProbabilisticRetinaNet':
        return RetinaNetProbabilisticPredictor(cfg)
    raise ValueError('Invalid meta-architecture {}.'.format(cfg.MODEL.META_ARCHITECTURE))


class _ModelFacade:
    """The attributes of `predictor.model` the reference reads (probabilistic_inference.py:207,294,
    300,304,326,407,449,621): both detectron2 API generations are exposed."""

    def __init__(self, cfg, device):
        r = cfg.MODEL.RETINANET
        pm = cfg.MODEL.PROBABILISTIC_MODELING
        self.num_classes = r.NUM_CLASSES
        self.in_features = self.head_in_features = list(r.IN_FEATURES)
        self.test_score_thresh = self.score_threshold = r.SCORE_THRESH_TEST
        self.test_topk_candidates = self.topk_candidates = r.TOPK_CANDIDATES_TEST
        self.test_nms_thresh = self.nms_threshold = r.NMS_THRESH_TEST
        self.max_detections_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        self.cls_var_num_samples = pm.CLS_VAR_LOSS.NUM_SAMPLES
        self.bbox_cov_num_samples = pm.BBOX_COV_LOSS.NUM_SAMPLES
        self.compute_cls_var = pm.CLS_VAR_LOSS.NAME != 'none'
        self.compute_bbox_cov = pm.BBOX_COV_LOSS.NAME != 'none'
        self.bbox_cov_dims = 4 if pm.BBOX_COV_LOSS.COVARIANCE_TYPE == 'diagonal' else 10
        self.dropout_rate = pm.DROPOUT_RATE
        self.use_dropout = self.dropout_rate != 0.0
        self.device = device
        self.training = False

    def train(self):
        self.training = True

    def eval(self):
        self.training = False


class ProbabilisticPredictor:
    """reference probabilistic_inference.py:36-111 (construction, mode dispatch, final rescale)."""

    def __init__(self, cfg):
        _cabi.require_device()          # fail loudly: no CPU / eager fallback exists for this path
        self.cfg = cfg.clone()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = _ModelFacade(self.cfg, self.device)
        self.model_list = []
        self.inference_mode = self.cfg.PROBABILISTIC_INFERENCE.INFERENCE_MODE
        self.mc_dropout_enabled = self.cfg.PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE
        self.num_mc_dropout_runs = self.cfg.PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS
        if self.mc_dropout_enabled:
            self.model.train()
        else:
            self.model.eval()
        self.backbone = None
        self.member_backbones = None      # ensembles: one backbone per member (load_backbone with E state dicts)
        # Pre-NMS aggregation never reads box_cls / box_cls_var / box_reg_var of the last sample or member
        # (reference probabilistic_inference.py:216-267 loops over range(len-1)): the tower passes feeding
        # only those outputs are left out.  Results are unchanged; set False to evaluate them anyway.
        self.skip_unread_outputs = True
        # Pre-NMS MC-dropout aggregation: the sample "mean" of box_cls / box_cls_var / box_reg_var commutes with the linear
        # output convolutions, so the reference's weighted sample mean is taken of the LAST TOWER LAYER and cls_score /
        # cls_var / bbox_cov run once per image (engine.HeadEngine.head_mc fuse_q1).  True / "stream": one HBM-bound pass
        # over the per-sample maps (default); "epilogue": accumulated inside the tcgen05 epilogue (slower).
        # Needs skip_unread_outputs; False evaluates and averages every sample's outputs as the reference does.
        self.fuse_sample_mean = True
        self.max_activation_bytes = 64e9  # auto-chunking budget of infer_from_features / predict_batch
        self._copy_stream = None          # host->device prefetch of the next chunk (infer_from_features chunk_images)
        self.rng_seed = int(self.cfg.SEED) if int(self.cfg.SEED) >= 0 else 0
        self.weight_sets = []
        self._engine = None
        self._anchor_cache = {}
        # checkpoints (probabilistic_inference.py:58-84): <OUTPUT_DIR>/model_final.pth, or the sibling
        # random_seed_<s> directories for ensembles; absent files leave the weights to load_weight_sets().
        paths = []
        if self.inference_mode == 'ensembles':
            for s in self.cfg.PROBABILISTIC_INFERENCE.ENSEMBLES.RANDOM_SEED_NUMS:
                paths.append(os.path.join(os.path.split(self.cfg.OUTPUT_DIR)[0], 'random_seed_' + str(s), 'model_final.pth'))
        else:
            paths.append(os.path.join(self.cfg.OUTPUT_DIR, 'model_final.pth'))
        if all(os.path.isfile(p) for p in paths):
            sds = []
            for p in paths:
                sd = torch.load(p, map_location="cpu")
                sds.append(sd.get("model", sd))
            self.load_weight_sets(sds)
            # a full-model checkpoint also carries the ResNet-FPN weights (detectron2 key names): with them
            # build_predictor(cfg)(input_im) works from raw images exactly like the reference's
            if all(any(k.startswith("backbone.") for k in sd) for sd in sds):
                self.load_backbone(sds if len(sds) > 1 else sds[0])

    # ---------------------------------------------------------------------------------------------
    def path_config(self):
        cfg, m = self.cfg, self.model
        pi = cfg.PROBABILISTIC_INFERENCE
        ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
        return PathConfig(
            num_classes=m.num_classes,
            num_anchors=len(cfg.MODEL.ANCHOR_GENERATOR.SIZES[0]) * len(ratios[0]),
            dropout_rate=m.dropout_rate, cls_var=m.compute_cls_var, bbox_cov=m.compute_bbox_cov,
            cov_dims=m.bbox_cov_dims, cls_var_num_samples=m.cls_var_num_samples, box_num_samples=1000,
            topk=m.test_topk_candidates, score_thresh=m.test_score_thresh, nms_thresh=m.test_nms_thresh,
            max_dets=m.max_detections_per_image, reg_weights=tuple(cfg.MODEL.RETINANET.BBOX_REG_WEIGHTS),
            # the sampled decode uses the RPN weights (reference probabilistic_inference.py:175-176)
            sample_reg_weights=tuple(cfg.MODEL.RPN.BBOX_REG_WEIGHTS),
            affinity=pi.AFFINITY_THRESHOLD, box_merge=pi.BAYES_OD.BOX_MERGE_MODE, cls_merge=pi.BAYES_OD.CLS_MERGE_MODE)

    def load_weight_sets(self, state_dicts):
        """One state dict (reference key names), or E of them for INFERENCE_MODE 'ensembles'."""
        if isinstance(state_dicts, dict):
            state_dicts = [state_dicts]
        m = self.model
        self.weight_sets = [HeadWeights(sd, m.use_dropout, m.compute_cls_var, m.compute_bbox_cov,
                                        self.cfg.MODEL.RETINANET.NUM_CONVS, self.device) for sd in state_dicts]
        self._engine = HeadEngine(self.path_config(), self.weight_sets, self.device)
        return self

    def load_backbone(self, state_dicts, impl="tc"):
        """Install the ResNet-50-FPN feature extractor (detectron2 key names, backbone.py) so that
        `predictor(input_im)` accepts raw images.  One state dict, or E of them for INFERENCE_MODE 'ensembles':
        every ensemble member of the reference is a full model with its own backbone, hence its own feature maps
        (reference probabilistic_inference.py:58-77,499-501; probabilistic_retinanet.py:99)."""
        if isinstance(state_dicts, dict):
            state_dicts = [state_dicts]
        if impl == "tc":
            from .backbone_tc import TcResNetFPNBackbone as Net      # hand-written sm_100a kernels (default)
        elif impl == "torch":
            from .backbone import ResNetFPNBackbone as Net           # torch library convolutions: the oracle of backbone_tc
        else:
            raise ValueError("backbone impl must be 'tc' or 'torch'")
        nets = [Net(sd, self.cfg.MODEL.PIXEL_MEAN, self.cfg.MODEL.PIXEL_STD, self.device) for sd in state_dicts]
        self.backbone = nets[0]
        self.member_backbones = nets if len(nets) > 1 else None
        return self

    def _anchors(self, level_hw):
        key = tuple(level_hw)
        if key not in self._anchor_cache:
            ag = self.cfg.MODEL.ANCHOR_GENERATOR
            sizes = ag.SIZES if len(ag.SIZES) == len(level_hw) else list(ag.SIZES) * len(level_hw)
            self._anchor_cache[key] = engine.make_anchors(level_hw, sizes, ag.ASPECT_RATIOS[0], (8, 16, 32, 64, 128),
                                                          ag.OFFSET, self.device)
        return self._anchor_cache[key]

    # ---------------------------------------------------------------------------------------------
    def __call__(self, input_im):
        """reference probabilistic_inference.py:86-111: one image in, Instances out."""
        if self.inference_mode not in ('standard_nms', 'mc_dropout_ensembles', 'anchor_statistics', 'ensembles', 'bayes_od'):
            raise ValueError('Invalid inference mode {}.'.format(self.inference_mode))
        return self.predict_batch([input_im])[0]

    def predict_batch_json(self, inputs, cat_mapping_dict):
        """The body of the reference's harness loop (src/apply_net.py:88-98: `predictor(input_im)` followed by
        `instances_to_json(outputs, image_id, cat_mapping_dict)`) for a whole batch: the detections stay on the device
        until ONE record kernel and ONE device->host copy (wire.py); returns the list of result dicts that
        `json.dump` writes to coco_instances_results.json."""
        from . import wire
        dicts = [x[0] if isinstance(x, (list, tuple)) else x for x in inputs]
        _, det = self.predict_batch(inputs, return_det=True)
        key = (id(cat_mapping_dict), int(det["probs"].shape[2]), int(det["probs"].shape[1]))
        if getattr(self, "_json_writer_key", None) != key:
            self._json_writer = wire.BatchJsonWriter(det["probs"].shape[2], det["probs"].shape[1], cat_mapping_dict, self.device)
            self._json_writer_key = key
        return self._json_writer.to_json(det, [d.get("image_id", i) for i, d in enumerate(dicts)])

    def predict_batch(self, inputs, return_det=False):
        """B independent single-image problems in one pass. `inputs` is a list of
        reference-style `input_im` lists.  return_det: also return the batch's detection buffers (device)."""
        dicts = [x[0] if isinstance(x, (list, tuple)) else x for x in inputs]
        hw = [tuple(d["image"].shape[-2:]) if "image" in d else tuple(d["image_hw"]) for d in dicts]
        if len(set(hw)) != 1:
            raise ValueError("predict_batch needs images of one size; got {}".format(sorted(set(hw))))
        out_hw = [(d.get("height", hw[0][0]), d.get("width", hw[0][1])) for d in dicts]
        if len(set(out_hw)) != 1:
            raise ValueError("predict_batch needs one output resolution per batch")
        if all("features" in d for d in dicts):
            def batch_levels(get):
                n_lvl = len(get(dicts[0]))
                return [torch.cat([get(d)[l].reshape((-1,) + tuple(get(d)[l].shape[-3:])) for d in dicts], 0)
                        for l in range(n_lvl)]
            if isinstance(dicts[0]["features"][0], (list, tuple)):      # ensembles: features[e][l], one set per member
                feats = [batch_levels(lambda d, e=e: d["features"][e]) for e in range(len(dicts[0]["features"]))]
            else:
                feats = batch_levels(lambda d: d["features"])
        else:
            if self.backbone is None:
                raise _cabi.PodError("no 'features' in the inputs and predictor.backbone is not set "
                                     "(the ResNet-FPN backbone is upstream of this path)")
            images = [d["image"] for d in dicts]
            if self.inference_mode == 'ensembles' and self.member_backbones is not None:
                feats = [net(images) for net in self.member_backbones]   # every member sees its own feature maps
            else:
                feats = self.backbone(images)
        ids = [d.get("image_id", i) for i, d in enumerate(dicts)]
        image0 = ids[0] if all(isinstance(i, int) for i in ids) and ids == list(range(ids[0], ids[0] + len(ids))) else 0
        if return_det:
            res, _, _, det = self.infer_from_features(feats, hw[0], out_hw[0], image0=image0, return_candidates=True)
            return res, det
        return self.infer_from_features(feats, hw[0], out_hw[0], image0=image0)

    def infer_from_images(self, images, out_hw=None, image0=0, seed=None, chunk_images=None, return_det=False):
        """images: (B, 3, H, W) uint8 / float tensor (host or device) of raw frames in the model's channel order.
        Backbone(s) + head path per chunk of `chunk_images` images (host frames of chunk i+1 are uploaded on the copy
        stream while chunk i computes).  The per-image entry `predictor(input_im)` and `predict_batch` end up here too."""
        if self.backbone is None:
            raise _cabi.PodError("predictor.backbone is not set: load_backbone(state_dict) first")
        B, _, H, W = images.shape
        chunk = int(chunk_images) if chunk_images else B
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)

        def stage(c0, c1):
            if images.is_cuda:
                return images[c0:c1], None
            with torch.cuda.stream(self._copy_stream):
                dev = images[c0:c1].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            dev.record_stream(main)
            return dev, ev

        bounds = [(c0, min(B, c0 + chunk)) for c0 in range(0, B, chunk)]
        res, dets = [], []
        nxt = stage(*bounds[0])
        for i, (c0, c1) in enumerate(bounds):
            cur, ev = nxt
            if i + 1 < len(bounds):
                nxt = stage(*bounds[i + 1])
            if ev is not None:
                main.wait_event(ev)
            if self.inference_mode == 'ensembles' and self.member_backbones is not None:
                feats = [net(cur) for net in self.member_backbones]      # every member sees its own feature maps
            else:
                feats = self.backbone(cur)
            out = self.infer_from_features(feats, (H, W), out_hw or (H, W), image0=image0 + c0, seed=seed, return_candidates=True)
            res.extend(out[0])
            dets.append(out[3])
        if not return_det:
            return res
        det = {k: (torch.cat([d[k] for d in dets], 0) if isinstance(v, torch.Tensor) else v) for k, v in dets[0].items()}
        return res, det

    def capture(self, example, out_hw=None, image0=0, seed=None):
        """CUDA-graph the whole call for a fixed input shape (launch-bound small batches: a single-forward batch of 8 is
        ~250 kernel launches of a few tens of microseconds each, and the Python / ctypes / tensor-map-encode cost per
        launch is of the same order).  `example` is a (B,3,H,W) frame tensor (backbone + head) or a list of FPN maps.
        Returns run(new_input, image_ids=None) -> (list of Instances, det dict); new inputs are copied into the captured
        buffers, the graph is replayed, and only the final count read-back + pod_status check happen on the host.
        The noise streams are keyed by image0 + position (fixed at capture), exactly as in the eager call."""
        from_images = isinstance(example, torch.Tensor)
        if from_images and self.backbone is None:
            raise _cabi.PodError("predictor.backbone is not set: load_backbone(state_dict) first")
        if from_images:
            static_in = example.to(self.device).contiguous().clone()
            H, W = int(static_in.shape[-2]), int(static_in.shape[-1])
        else:
            static_in = [f.to(self.device, dtype=torch.float32).contiguous().clone() for f in example]
            if out_hw is None:
                raise ValueError("capture(features) needs out_hw / image size")
            H, W = out_hw
        out_hw = tuple(out_hw) if out_hw is not None else (H, W)

        def forward():
            if from_images:
                if self.inference_mode == 'ensembles' and self.member_backbones is not None:
                    feats = [net(static_in) for net in self.member_backbones]
                else:
                    feats = self.backbone(static_in)
            else:
                feats = static_in
            return self._forward_det(feats, (H, W), out_hw, image0, seed)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                       # warm-up: one-off attribute calls, buffer growth, lazy init
            forward()
            forward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        ops.check_status()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(graph):
            det = forward()
        launches = ops.launch_count() - n0                  # kernels of this library inside one replay

        def run(new_input):
            if from_images:
                static_in.copy_(new_input, non_blocking=True)
            else:
                for dst, src in zip(static_in, new_input):
                    dst.copy_(src, non_blocking=True)
            graph.replay()
            ops._count(launches)
            return self._to_instances(det, out_hw), det
        run.graph, run.launches = graph, launches
        return run

    def _forward_det(self, feats, image_hw, out_hw, image0, seed):
        """Device-only part of infer_from_features (no host synchronisation): features -> detection buffers."""
        out = self.infer_from_features(feats, image_hw, out_hw, image0=image0, seed=seed, return_candidates=True, _no_host=True)
        return out

    def infer_from_features(self, feats, image_hw, out_hw=None, image0=0, seed=None, return_raw=False,
                            return_candidates=False, chunk_images=None, _no_host=False):
        """feats: list over FPN levels of (B,256,Hl,Wl) fp32 tensors (host or device).
        image_hw: size of the network input (Instances.image_size before rescale,
        inference_utils.py:39-41); out_hw: requested output resolution (:374-397).
        chunk_images: evaluate the batch in chunks of at most this many images (bounds the activation memory:
        2 x chunk x N x passes maps per level); host-resident features of chunk i+1 are uploaded on a copy stream
        while chunk i computes.  Results are identical to one call per chunk (noise streams are keyed by image id)."""
        if self._engine is None:
            raise _cabi.PodError("no weights loaded: call load_weight_sets(state_dicts) first")
        # ensembles may carry one feature set per member: feats[e][l] (each member of the reference is a full model
        # with its own backbone, probabilistic_inference.py:499-501); a flat list is shared by all members
        per_member = isinstance(feats[0], (list, tuple))
        if per_member and self.inference_mode != 'ensembles':
            raise ValueError("per-member feature sets are only meaningful for INFERENCE_MODE 'ensembles'")
        if per_member and len(feats) != len(self.weight_sets):
            raise _cabi.PodError("got %d feature sets for %d ensemble members" % (len(feats), len(self.weight_sets)))
        lv0 = feats[0] if per_member else feats
        B = int(lv0[0].shape[0])
        if chunk_images is None and B > 1 and not _no_host:
            # keep the activation working set inside the budget (default 64 GB of the 180 GB): a batch of 64 images at
            # N=30 would otherwise ask for 2 x 64 x 60 maps of 96x160x256 split pairs = 240 GB
            auto = max(1, int(self.max_activation_bytes // self._activation_bytes_per_image(lv0)))
            if auto < B and not return_raw:
                chunk_images = auto
        if chunk_images is not None and 0 < int(chunk_images) < B:
            if return_raw:
                raise ValueError("return_raw is per chunk: call infer_from_features once per chunk")
            return self._infer_chunked(feats, image_hw, out_hw, image0, seed, return_candidates, int(chunk_images))
        mode = self.inference_mode
        pi = self.cfg.PROBABILISTIC_INFERENCE
        post_nms = ((mode == 'mc_dropout_ensembles' and pi.ENSEMBLES_DROPOUT.BOX_MERGE_MODE != 'pre_nms') or
                    (mode == 'ensembles' and pi.ENSEMBLES.BOX_MERGE_MODE != 'pre_nms'))
        if mode not in SUPPORTED_PRE_NMS_MODES:
            raise ValueError('Invalid inference mode {}.'.format(mode))
        out_hw = tuple(out_hw) if out_hw is not None else tuple(image_hw)
        seed = self.rng_seed if seed is None else seed
        eng = self._engine
        def to_dev(fs):
            # channels-last views (backbone_tc output) keep their layout; everything else becomes NCHW-contiguous
            return [f if (f.is_cuda and f.dtype == torch.float32 and HeadEngine._is_channels_last(f))
                    else f.to(self.device, dtype=torch.float32, non_blocking=True).contiguous() for f in fs]
        feats = [to_dev(fs) for fs in feats] if per_member else to_dev(feats)
        level_hw = [tuple(f.shape[-2:]) for f in (feats[0] if per_member else feats)]
        anchors = self._anchors(level_hw)
        if mode == 'ensembles':
            if self.mc_dropout_enabled:
                # The reference does not compose the two: with NUM_RUNS > 1 its pre-NMS branch discards the members'
                # outputs and runs MC-dropout on the un-loaded base model (probabilistic_inference.py:196-205), and its
                # post-NMS branch replicates every eval()-mode member NUM_RUNS times.  None of its configs enables both.
                raise _cabi.PodError("INFERENCE_MODE 'ensembles' with MC_DROPOUT.ENABLE is not supported (the reference "
                                     "does not combine them either: see predictor.py)")
            if len(self.weight_sets) != len(pi.ENSEMBLES.RANDOM_SEED_NUMS):
                raise _cabi.PodError("ensembles mode needs one weight set per RANDOM_SEED_NUMS entry")
            raw, level_off = eng.head_eval(feats, skip_unread=self.skip_unread_outputs and not post_nms,
                                           per_member_feats=per_member)
        elif self.mc_dropout_enabled and self.model.use_dropout:
            # NUM_RUNS == 1 keeps the reference's behaviour too: the model stays in train mode (:52-56), so the single
            # forward has active dropout and the mean / variance heads see independently masked tower passes (Q2)
            n_runs = max(1, int(self.num_mc_dropout_runs))
            skip = self.skip_unread_outputs and not post_nms and n_runs > 1
            raw, level_off = eng.head_mc(feats, n_runs, seed, image0, skip_unread=skip,
                                         fuse_q1=(self.fuse_sample_mean if isinstance(self.fuse_sample_mean, str) else "stream")
                                         if (skip and self.fuse_sample_mean and not return_raw) else False)
        elif self.mc_dropout_enabled and self.num_mc_dropout_runs > 1:
            raise _cabi.PodError("MC_DROPOUT.ENABLE with DROPOUT_RATE == 0 is not supported")
        else:
            raw, level_off = eng.head_eval(feats, members=[0])
        if post_nms:
            cand, _, _, det = eng.merged_detections(raw, level_off, anchors, seed, image0, image_hw, out_hw)
        else:
            cand = eng.candidates(raw, level_off, anchors, seed, image0)
            det = eng.detections(cand, {'bayes_od': 1, 'anchor_statistics': 2}.get(mode, 0), image_hw, out_hw)
        if _no_host:
            return det                      # CUDA-graph capture: the host-side read-back happens after the replay
        res = self._to_instances(det, out_hw)
        if return_raw or return_candidates:
            return res, (raw if return_raw else None), cand, det
        return res

    def _activation_bytes_per_image(self, feats):
        """Device bytes one image adds to a head pass: the two ping-pong activation buffers (maps x largest level x
        256 channels x 4 B) plus the raw per-sample outputs (22-32 floats per anchor and sample)."""
        m = self.model
        mc = self.mc_dropout_enabled and self.num_mc_dropout_runs > 1
        samples = self.num_mc_dropout_runs if mc else max(1, len(self.weight_sets))
        maps = samples * (2 if (m.compute_cls_var or m.compute_bbox_cov) else 1) if mc else 1
        hw = [int(f.shape[-2]) * int(f.shape[-1]) for f in feats]
        anchors = sum(hw) * len(self.cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS[0]) * len(self.cfg.MODEL.ANCHOR_GENERATOR.SIZES[0])
        per_anchor = 2 * m.num_classes + 4 + max(4, m.bbox_cov_dims)
        return 2 * maps * max(hw) * 256 * 4 + max(hw) * 256 * 4 + samples * anchors * per_anchor * 4

    def _infer_chunked(self, feats, image_hw, out_hw, image0, seed, return_candidates, chunk):
        per_member = isinstance(feats[0], (list, tuple))
        flat = [f for fs in feats for f in fs] if per_member else list(feats)
        n_lvl = len(feats[0]) if per_member else len(feats)
        B = int(flat[0].shape[0])
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        bounds = [(c0, min(B, c0 + chunk)) for c0 in range(0, B, chunk)]

        def nest(ts):
            return [ts[i:i + n_lvl] for i in range(0, len(ts), n_lvl)] if per_member else ts

        def stage(c0, c1):
            """Slice of the batch on the device + the event after which it may be read."""
            if all(f.is_cuda for f in flat):
                return nest([f[c0:c1] for f in flat]), None
            with torch.cuda.stream(self._copy_stream):
                dev = [f[c0:c1].to(self.device, dtype=torch.float32, non_blocking=True) for f in flat]
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            for t in dev:
                t.record_stream(main)              # allocated on the copy stream, consumed on the compute stream
            return nest(dev), ev

        res, cands, dets = [], [], []
        nxt = stage(*bounds[0])
        for i, (c0, c1) in enumerate(bounds):
            cur, ev = nxt
            if i + 1 < len(bounds):
                nxt = stage(*bounds[i + 1])        # issued before this chunk's kernels: overlaps with them
            if ev is not None:
                main.wait_event(ev)
            out = self.infer_from_features(cur, image_hw, out_hw, image0=image0 + c0, seed=seed, return_candidates=True)
            res.extend(out[0])
            cands.append(out[2])
            dets.append(out[3])
        if not return_candidates:
            return res

        def merge(ds):
            m = {}
            for k, v in ds[0].items():
                m[k] = torch.cat([d[k] for d in ds], 0) if isinstance(v, torch.Tensor) else v
            return m
        return res, None, merge(cands), merge(dets)

    def _to_instances(self, det, out_hw):
        counts = det["count"].cpu().tolist()
        # the count read above synchronised the stream: one look at the library's device-side error word per call
        # (expired bounded barrier wait, activation outside the fp16 split range, non-finite input) -> PodError
        ops.check_status()
        out = []
        for b, n in enumerate(counts):
            inst = Instances((int(out_hw[0]), int(out_hw[1])))
            inst.pred_boxes = Boxes(det["boxes"][b, :n])
            inst.scores = det["scores"][b, :n]
            inst.pred_classes = det["classes"][b, :n].to(torch.int64)
            inst.pred_cls_probs = det["probs"][b, :n]
            inst.pred_boxes_covariance = det["cov"][b, :n]
            out.append(inst)
        return out


class RetinaNetProbabilisticPredictor(ProbabilisticPredictor):
    """reference probabilistic_inference.py:169-636; the five `post_processing_*` entry points keep
    their names and accept the same `input_im`."""

    def post_processing_standard_nms(self, input_im):
        return self._single(input_im, 'standard_nms')

    def post_processing_mc_dropout_ensembles(self, input_im):
        return self._single(input_im, 'mc_dropout_ensembles')

    def post_processing_ensembles(self, input_im, model_list=None):
        return self._single(input_im, 'ensembles')

    def post_processing_bayes_od(self, input_im):
        return self._single(input_im, 'bayes_od')

    def post_processing_anchor_statistics(self, input_im):
        return self._single(input_im, 'anchor_statistics')

    def _single(self, input_im, mode):
        saved = self.inference_mode
        self.inference_mode = mode
        try:
            return self.predict_batch([input_im])[0]
        finally:
            self.inference_mode = saved
