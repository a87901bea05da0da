#!/usr/bin/env python
"""Measures, per BayesOD fixture, the two deviations that bound the a12 tolerance:

  A = | reference fp32 (LAPACK sgetrf/sgetri, tests/golden)  -  oracle evaluated in fp64 |     the reference's own error
  B = | GPU (fp64 closed-form 4x4 inverses in k_nms_fuse)     -  reference fp32            |     what the tests bound

for the stage-isolated planted candidate sets (identical inputs on both sides) and for the end-to-end BayesOD cases
(the GPU's candidates carry the head's ~1e-6 numerical difference, which the fusion amplifies by the conditioning of
the summed precisions).  Boxes in pixels (absolute), covariances relative to the largest entry of each matrix.

    python tools/bayesod_envelope.py > gpurun_out/bayesod_envelope.txt      (needs a B200)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases as C  # noqa: E402
from oracle import podref as O  # noqa: E402
from pod_compare_b200 import ops, synthetic as S  # noqa: E402
from pod_compare_b200.predictor import build_predictor  # noqa: E402
from tests import gpu_util as G  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def covrel(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    if ref.size == 0:
        return 0.0
    scale = np.abs(ref).reshape(ref.shape[0], -1).max(1).reshape(-1, 1, 1)
    return float((np.abs(a - ref) / scale).max())


def boxabs(a, ref):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max()) if ref.size else 0.0


def main():
    print("# BayesOD fusion: measured deviations (boxes: max abs px; cov: max rel to the matrix scale)")
    print("# fixture | n | A box  A cov (ref fp32 vs oracle fp64) | B box  B cov (GPU vs ref fp32) | GPU vs oracle fp64 box  cov | max cond(sum P)")
    for tag in ("small", "large"):
        g = np.load(os.path.join(GOLDEN, "planted_%s.npz" % tag))
        pp = O.PathParams(cls_var=True, bbox_cov=True)
        cand = O.Candidates(torch.from_numpy(g["in_boxes"]), torch.from_numpy(g["in_cov"]), torch.from_numpy(g["in_scores"]),
                            torch.from_numpy(g["in_classes"]), torch.from_numpy(g["in_probs"]),
                            np.arange(g["in_boxes"].shape[0]), [g["in_boxes"].shape[0]])
        cd = G.cand_to_dict(cand)
        for cm, ck in (("max_score", "ms"), ("bayesian_inference", "avg")):
            for bm, bk in (("bayesian_inference", "bi"), ("covariance_intersection", "ci")):
                pp.cls_merge, pp.box_merge = cm, bm
                key = "bod_%s_%s_" % (ck, bk)
                det = ops.nms_fuse(cd, 1, 0.5, 0.9, 100, (720, 1280), (720, 1280), box_merge=0 if bk == "bi" else 1,
                                   cls_merge=0 if ck == "ms" else 1)
                n = int(det["count"][0])
                r64 = O.detector_postprocess(O.bayes_od_post(cand, pp, (720, 1280), dtype=np.float64), 720, 1280)
                gb, gc = det["boxes"][0, :n].cpu().numpy(), det["cov"][0, :n].cpu().numpy()
                conds = np.linalg.cond(np.linalg.inv(g[key + "cov"].astype(np.float64)))
                print("planted_%s/%s | %d | %.2e %.2e | %.2e %.2e | %.2e %.2e | %.1e" % (
                    tag, key[:-1], n, boxabs(r64.boxes.numpy(), g[key + "boxes"]), covrel(r64.cov.numpy(), g[key + "cov"]),
                    boxabs(gb, g[key + "boxes"]), covrel(gc, g[key + "cov"]),
                    boxabs(gb, r64.boxes.numpy()), covrel(gc, r64.cov.numpy()), conds.max()))
    # end-to-end BayesOD cases: GPU head + GPU fusion vs the reference fixture, matched by nearest box
    for name in ("bayesod_plain", "bayesod_mc_n3", "bayesod_clsavg_ci"):
        opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
        cfg = C.build_cfg(name)
        pp = O.PathParams.from_cfg(cfg)
        sd = S.make_head_state_dict(seeds[0], num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                    bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims)
        feats = C.case_features(name)
        pred = build_predictor(cfg)
        pred.load_weight_sets(sd)
        res = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
        g = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
        gb, rb = res.pred_boxes.tensor.cpu().numpy(), g["final_boxes"]
        torch.set_num_threads(8)
        f32, c32, _ = O.predict(feats, [O.unpack_head(sd, pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed, image=img,
                                return_candidates=True)
        r64 = O.detector_postprocess(O.bayes_od_post(c32, pp, hw, dtype=np.float64), out_hw[0], out_hw[1])
        if gb.shape == rb.shape:
            print("e2e/%s | %d | %.2e %.2e | %.2e %.2e | %.2e %.2e | %.1e" % (
                name, len(rb), boxabs(r64.boxes.numpy(), rb), covrel(r64.cov.numpy(), g["final_cov"]),
                boxabs(gb, rb), covrel(res.pred_boxes_covariance.cpu().numpy(), g["final_cov"]),
                boxabs(gb, r64.boxes.numpy()), covrel(res.pred_boxes_covariance.cpu().numpy(), r64.cov.numpy()),
                np.linalg.cond(np.linalg.inv(g["final_cov"].astype(np.float64))).max()))
        else:
            print("e2e/%s | detection count differs: GPU %d vs reference %d" % (name, len(gb), len(rb)))


if __name__ == "__main__":
    main()
