"""Live cross-check of the oracle restatement against the UNMODIFIED reference at geometries, sample counts and
seeds that are NOT among the committed fixtures.  Runs where /root/reference exists (the build container) or where
oracle/make_ref.py staged the reference's files under oracle/_ref (they travel to the GPU box); otherwise the
fixtures of tests/golden pin the oracle (tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_runner as _R

pytestmark = pytest.mark.skipif(not _R.reference_available(),
                                reason="the reference sources are neither at /root/reference nor staged under oracle/_ref")

from oracle import cases as C
from oracle import podref as O
from pod_compare_b200 import synthetic as S

LIVE = [  # (fixture case whose cfg is reused, image (H, W), output (H, W), MC samples or None, seed, image id, weight seed)
    ("mcdrop_pre_n4", (128, 192), (96, 144), 5, 71, 13, 4000),
    ("bayesod_mc_n3", (64, 128), (64, 128), 2, 72, 14, 3000),
    ("regclsvar_std", (160, 96), (200, 120), None, 73, 15, 2000),
    ("anchorstats_var", (96, 96), (96, 96), None, 74, 16, 1000),
]


@pytest.mark.parametrize("name,hw,out_hw,n_mc,seed,img,wseed", LIVE)
def test_oracle_equals_live_reference(name, hw, out_hw, n_mc, seed, img, wseed):
    from oracle import ref_runner as R
    torch.set_num_threads(8)
    opts, mode, n_mc0, seeds, _, _, _, _ = C.CASES[name]
    cfg = C.build_cfg(name)
    if n_mc is not None:
        cfg.defrost()
        cfg.PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS = n_mc
        cfg.freeze()
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(wseed, num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims)
    feats = S.make_features(0, img, hw[0], hw[1])
    pred = R.build_reference_predictor(cfg, sd)
    final, _ = R.run_reference(pred, feats, hw, out_hw=out_hw, seed=seed, image_idx=img, stage="final")
    ref = R.instances_to_arrays(final)
    got = O.predict(feats, [O.unpack_head(sd, pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc or n_mc0, seed=seed, image=img)
    assert got.boxes.shape[0] == ref["boxes"].shape[0] and got.boxes.shape[0] > 0
    assert np.array_equal(got.boxes.numpy(), ref["boxes"])
    assert np.array_equal(got.scores.numpy(), ref["scores"])
    assert np.array_equal(got.classes.numpy(), ref["classes"])
    assert np.array_equal(got.probs.numpy(), ref["probs"])
    assert np.array_equal(got.cov.numpy(), ref["cov"])


def test_oracle_equals_live_reference_ensembles_with_per_member_features():
    """a15: every ensemble member of the reference is a full model with its own backbone, i.e. its own feature maps
    (probabilistic_inference.py:58-77,499-501).  Fresh geometry / seeds, pre- and post-NMS merging."""
    from oracle import ref_runner as R
    torch.set_num_threads(8)
    for name, hw, seed, img in (("ensembles_e3", (64, 128), 81, 21), ("ensembles_post_e3", (96, 96), 82, 22)):
        opts, mode, _, seeds, _, _, _, _ = C.CASES[name]
        cfg = C.build_cfg(name)
        pp = O.PathParams.from_cfg(cfg)
        sds = [S.make_head_state_dict(s + 7, num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                      bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims) for s in seeds]
        feats = S.make_member_features(len(seeds), img, hw[0], hw[1])
        assert not torch.equal(feats[0][0], feats[1][0])
        pred = R.build_reference_predictor(cfg, sds)
        final, _ = R.run_reference(pred, feats, hw, seed=seed, image_idx=img, stage="final")
        ref = R.instances_to_arrays(final)
        got = O.predict(feats, [O.unpack_head(sd, pp) for sd in sds], pp, mode, hw, seed=seed, image=img,
                        post_nms=C.is_post_nms(name))
        assert got.boxes.shape[0] == ref["boxes"].shape[0] and got.boxes.shape[0] > 0
        assert np.array_equal(got.boxes.numpy(), ref["boxes"]) and np.array_equal(got.scores.numpy(), ref["scores"])
        assert np.array_equal(got.cov.numpy(), ref["cov"]) and np.array_equal(got.probs.numpy(), ref["probs"])
        # and the members' maps matter: the shared-feature evaluation gives a different result
        shared = O.predict(feats[0], [O.unpack_head(sd, pp) for sd in sds], pp, mode, hw, seed=seed, image=img,
                           post_nms=C.is_post_nms(name))
        assert shared.boxes.shape != got.boxes.shape or not np.array_equal(shared.boxes.numpy(), got.boxes.numpy())


@pytest.mark.parametrize("name,hw,seed,img", [("mcdrop_single", (64, 96), 83, 23), ("regclsvar_rpnw", (96, 64), 84, 24)])
def test_oracle_equals_live_reference_single_run_dropout_and_rpn_weights(name, hw, seed, img):
    from oracle import ref_runner as R
    torch.set_num_threads(8)
    opts, mode, n_mc, seeds, _, _, _, _ = C.CASES[name]
    cfg = C.build_cfg(name)
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(seeds[0] + 3, num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims)
    feats = S.make_features(0, img, hw[0], hw[1])
    pred = R.build_reference_predictor(cfg, sd)
    final, _ = R.run_reference(pred, feats, hw, seed=seed, image_idx=img, stage="final")
    ref = R.instances_to_arrays(final)
    got = O.predict(feats, [O.unpack_head(sd, pp)], pp, mode, hw, n_mc=n_mc, seed=seed, image=img,
                    mc_single=C.is_mc_single(name))
    assert got.boxes.shape[0] == ref["boxes"].shape[0] and got.boxes.shape[0] > 0
    assert np.array_equal(got.boxes.numpy(), ref["boxes"]) and np.array_equal(got.scores.numpy(), ref["scores"])
    assert np.array_equal(got.cov.numpy(), ref["cov"])
