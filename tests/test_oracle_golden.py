"""The restated oracle (oracle/podref.py) must reproduce the fixtures that oracle/make_golden.py
produced by running the UNMODIFIED reference (tests/golden/*.npz) -- bit for bit: both are fp32
torch CPU programs evaluating the same operations in the same order."""
import os

import numpy as np
import pytest
import torch

from oracle import cases as C
from oracle import philox
from oracle import podref as O
from pod_compare_b200 import synthetic as S


def test_philox_known_answers():
    for ctr, key, expect in philox.KAT:
        got = tuple(int(x) for x in philox.philox4x32(*ctr, key[0], key[1]))
        assert got == expect


def _run_case(name):
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg = C.build_cfg(name)
    pp = O.PathParams.from_cfg(cfg)
    hws = [O.unpack_head(S.make_head_state_dict(s, num_classes=pp.num_classes, use_dropout=pp.use_dropout,
                                                cls_var=pp.cls_var, bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims), pp)
           for s in seeds]
    feats = C.case_features(name)
    return feats, O.predict(feats, hws, pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed, image=img,
                            return_candidates=True, post_nms=C.is_post_nms(name), mc_single=C.is_mc_single(name))


@pytest.mark.parametrize("name", list(C.CASES))
def test_case_matches_reference_fixture(name, golden_dir):
    torch.set_num_threads(8)   # fixtures were generated with 8 threads (summation order of conv)
    g = np.load(os.path.join(golden_dir, "case_%s.npz" % name))
    feats, (final, cand, det) = _run_case(name)
    flat = [f for fs in feats for f in fs] if isinstance(feats[0], (list, tuple)) else feats
    chk = np.array([float(f.double().abs().sum()) for f in flat])
    assert np.allclose(chk, g["feats_checksum"], rtol=0, atol=0), "synthetic feature generator drifted"
    exact = torch.get_num_threads() == 8
    cmp = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, rtol=1e-5, atol=1e-6))
    if C.is_post_nms(name):
        assert cmp(final.boxes.numpy(), g["final_boxes"]) and cmp(final.scores.numpy(), g["final_scores"])
        assert np.array_equal(final.classes.numpy(), g["final_classes"])
        assert cmp(final.probs.numpy(), g["final_probs"]) and cmp(final.cov.numpy(), g["final_cov"])
        return
    # anchor-wise stage (probabilistic_inference.py:178-388)
    if g["cand_anchor_ids"].size:   # captured only when the aleatoric branch (:344-374) ran
        assert np.array_equal(cand.anchor_ids, g["cand_anchor_ids"])
    exact = torch.get_num_threads() == 8
    cmp = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, rtol=1e-5, atol=1e-6))
    assert cmp(cand.boxes.numpy(), g["cand_boxes"])
    assert cmp(cand.scores.numpy(), g["cand_scores"])
    assert np.array_equal(cand.classes.numpy(), g["cand_classes"])
    assert cmp(cand.probs.numpy(), g["cand_probs"])
    if bool(g["cand_has_cov"]):
        assert cmp(cand.cov.numpy(), g["cand_cov"])
    # final Instances (predictor.__call__, :86-111)
    assert cmp(final.boxes.numpy(), g["final_boxes"])
    assert cmp(final.scores.numpy(), g["final_scores"])
    assert np.array_equal(final.classes.numpy(), g["final_classes"])
    assert cmp(final.probs.numpy(), g["final_probs"])
    assert cmp(final.cov.numpy(), g["final_cov"])


@pytest.mark.parametrize("tag", ["small", "large"])
def test_planted_postprocessing_matches_reference_fixture(tag, golden_dir):
    g = np.load(os.path.join(golden_dir, "planted_%s.npz" % tag))
    pp = O.PathParams(cls_var=True, bbox_cov=True)
    cand = O.Candidates(torch.from_numpy(g["in_boxes"]), torch.from_numpy(g["in_cov"]), torch.from_numpy(g["in_scores"]),
                        torch.from_numpy(g["in_classes"]), torch.from_numpy(g["in_probs"]),
                        np.arange(g["in_boxes"].shape[0]), [g["in_boxes"].shape[0]])
    for impl in ("torchvision", "loop"):
        d = O.standard_nms_post(cand, pp, (720, 1280), nms_impl=impl)
        assert np.array_equal(d.boxes.numpy(), g["std_boxes"])
        assert np.array_equal(d.scores.numpy(), g["std_scores"])
        assert np.array_equal(d.classes.numpy(), g["std_classes"])
        assert np.array_equal(d.cov.numpy(), g["std_cov"])
    for use_cov, key in ((True, "ast_cov_"), (False, "ast_nocov_")):
        c2 = O.Candidates(cand.boxes, cand.cov if use_cov else None, cand.scores, cand.classes, cand.probs,
                          cand.anchor_ids, cand.level_counts)
        d = O.detector_postprocess(O.anchor_statistics_post(c2, pp, (720, 1280)), 720, 1280)
        assert np.array_equal(d.boxes.numpy(), g[key + "boxes"]), key
        assert np.array_equal(d.scores.numpy(), g[key + "scores"]), key
        assert np.array_equal(d.classes.numpy(), g[key + "classes"]), key
        assert np.array_equal(d.probs.numpy(), g[key + "probs"]), key
        assert np.array_equal(d.cov.numpy(), g[key + "cov"]), key
    for cm, ck in (("max_score", "ms"), ("bayesian_inference", "avg")):
        for bm, bk in (("bayesian_inference", "bi"), ("covariance_intersection", "ci")):
            pp.cls_merge, pp.box_merge = cm, bm
            d = O.detector_postprocess(O.bayes_od_post(cand, pp, (720, 1280)), 720, 1280)
            key = "bod_%s_%s_" % (ck, bk)
            assert np.array_equal(d.boxes.numpy(), g[key + "boxes"]), key
            assert np.array_equal(d.scores.numpy(), g[key + "scores"]), key
            assert np.array_equal(d.classes.numpy(), g[key + "classes"]), key
            assert np.array_equal(d.probs.numpy(), g[key + "probs"]), key
            assert np.array_equal(d.cov.numpy(), g[key + "cov"]), key
