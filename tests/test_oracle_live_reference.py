"""Live cross-check of the oracle restatement against the UNMODIFIED reference at geometries, sample counts and
seeds that are NOT among the committed fixtures.  Runs only where /root/reference exists (the build container);
on the GPU box the fixtures of tests/golden pin the oracle instead (tests/test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference sources are not on this machine")

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
