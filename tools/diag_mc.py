"""Diagnostic (not a test): per-sample / per-level comparison of the GPU head's raw outputs with the
oracle's for one parity case.  python tools/diag_mc.py <case>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import cases as C
from oracle import podref as O
from pod_compare_b200 import synthetic as S
from pod_compare_b200.predictor import build_predictor

name = sys.argv[1] if len(sys.argv) > 1 else "bayesod_mc_n3"
opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
cfg = C.build_cfg(name)
pp = O.PathParams.from_cfg(cfg)
sds = [S.make_head_state_dict(s, num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                              bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims) for s in seeds]
feats = S.make_features(0, img, hw[0], hw[1])
pred = build_predictor(cfg)
pred.skip_unread_outputs = False      # the raw outputs of every sample are inspected below
pred.load_weight_sets(sds if len(sds) > 1 else sds[0])
res, raw, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_raw=True)
torch.set_num_threads(8)
hws = [O.unpack_head(sd, pp) for sd in sds]
drop = O.DropoutSource("philox" if n_mc > 1 else "off", pp.dropout_rate, seed, img)
S_ = raw["logits"].shape[1]
if mode == "ensembles":
    outs = [O.head_outputs(feats, h, pp, O.DropoutSource("off", 0.0)) for h in hws]
else:
    outs = [O.head_outputs(feats, hws[0], pp, drop, sample=s) for s in range(S_)]
keys = (("logits", "box_cls"), ("deltas", "box_delta"), ("logvar", "box_cls_var"), ("regvar", "box_reg_var"))
print("case", name, "S", S_)
for s in range(S_):
    for gk, ok in keys:
        if raw[gk] is None:
            continue
        off = 0
        for l in range(len(feats)):
            ref = outs[s][ok][l][0]
            n = ref.shape[0]
            got = raw[gk][0, s, off:off + n].cpu()
            off += n
            d = (got - ref).double()
            print("sample %d %-7s level %d: max|d| %.3e  mean d %+.3e  max|ref| %.3f  frac(|d|>1e-4) %.4f" % (
                s, gk, l, float(d.abs().max()), float(d.mean()), float(ref.abs().max()), float((d.abs() > 1e-4).double().mean())))
