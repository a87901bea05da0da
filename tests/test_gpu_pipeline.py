"""GPU parity tests of the post-processing stages and of the whole path (features -> Instances)
against oracle/podref.py and the reference-generated fixtures in tests/golden.

Selection steps (top-k, 0.05 threshold, NMS at 0.5, 0.9 affinity) are discontinuous: they are
checked bit-exactly stage-isolated (identical inputs), and end-to-end on the fixture cases, whose
margins at those discontinuities are far above the 1e-6 numerical differences of the head.
BayesOD fuses with fp64 4x4 inverses in-kernel while the reference uses fp32 LAPACK: identical
inputs are compared to the oracle evaluated in fp64 at 1e-5, and to the fp32 reference fixtures at
the measured envelope of the reference's own fp32 rounding (2e-3 px / 2e-5; profiles/r2a_bayesod_envelope.txt).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cases as C
from oracle import podref as O
from pod_compare_b200 import engine, ops, synthetic as S
from pod_compare_b200._cabi import PodError
from pod_compare_b200.predictor import build_predictor
from tests import gpu_util as G

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    from pod_compare_b200 import _cabi
    _cabi.require_device()


def _cov_close(got, ref, tol):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    if ref.size == 0:
        return got.size == 0
    scale = np.abs(ref).reshape(ref.shape[0], -1).max(1).reshape(-1, 1, 1)
    return bool((np.abs(got - ref) <= tol * scale + 1e-7).all())


def _close(got, ref, rtol, atol, what):
    """np.allclose with a diagnosis: where and by how much the worst element deviates."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    excess = np.abs(got - ref) - (atol + rtol * np.abs(ref))
    if excess.size and excess.max() > 0:
        i = np.unravel_index(int(excess.argmax()), excess.shape)
        raise AssertionError("%s: worst element %s got %.9g ref %.9g (abs %.3g, rel %.3g; rtol %g atol %g); %d of %d outside"
                             % (what, i, got[i], ref[i], abs(got[i] - ref[i]), abs(got[i] - ref[i]) / max(abs(ref[i]), 1e-300),
                                rtol, atol, int((excess > 0).sum()), excess.size))
    return True


def _oracle_case(name, keep_diag=False):
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg = C.build_cfg(name)
    pp = O.PathParams.from_cfg(cfg)
    sds = [S.make_head_state_dict(s, num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                  bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims) for s in seeds]
    feats = C.case_features(name)          # ensemble cases: one feature set per member
    return cfg, pp, sds, feats


# ------------------------------------------------------------------------------------------ decode + covariance
@pytest.mark.parametrize("name", ["mcdrop_pre_n4", "regclsvar_std", "droponly_pre_n3", "fullcov_mc_n3", "baseline_std"])
def test_candidates_from_oracle_raw_outputs(name):
    """Stage-isolated: the oracle's raw head outputs are uploaded, then Q1 means, scores, top-k,
    decode and covariance run on the GPU and must reproduce probabilistic_inference.py:214-388."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    hws = [O.unpack_head(sd, pp) for sd in sds]
    torch.set_num_threads(8)
    drop = O.DropoutSource("philox" if n_mc > 1 else "off", pp.dropout_rate, seed, img)
    outs = [O.head_outputs(feats, hws[0], pp, drop, sample=s) for s in range(max(n_mc, 1))]
    level_hw = [tuple(f.shape[-2:]) for f in feats]
    anchors = O.make_anchors(level_hw, pp)
    ref = O.anchorwise(outs, anchors, pp, seed, img)

    def stack(key):
        if outs[0][key] is None:
            return None
        return torch.stack([torch.cat([o[key][l][0] for l in range(len(feats))], 0) for o in outs], 0)[None].contiguous().cuda()

    raw = {"logits": stack("box_cls"), "deltas": stack("box_delta"), "logvar": stack("box_cls_var"),
           "regvar": stack("box_reg_var")}
    level_off = [0]
    for a in anchors:
        level_off.append(level_off[-1] + a.shape[0])
    pc = engine.PathConfig(num_classes=pp.num_classes, dropout_rate=pp.dropout_rate, cls_var=pp.cls_var,
                           bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims, cls_var_num_samples=pp.cls_var_num_samples)
    eng = engine.HeadEngine(pc, [], "cuda")
    cand = eng.candidates(raw, level_off, torch.cat(anchors).cuda().contiguous(), seed, img)
    M = int(cand["count"][0])
    assert M == ref.boxes.shape[0]
    assert np.array_equal(cand["anchor"][0, :M].cpu().numpy().astype(np.int64), ref.anchor_ids)
    assert torch.equal(cand["classes"][0, :M].cpu().long(), ref.classes)
    assert torch.allclose(cand["scores"][0, :M].cpu(), ref.scores, rtol=1e-4, atol=1e-7)
    assert torch.allclose(cand["probs"][0, :M].cpu(), ref.probs, rtol=1e-4, atol=1e-7)
    assert torch.allclose(cand["boxes"][0, :M].cpu(), ref.boxes, rtol=1e-4, atol=2e-3)
    if isinstance(ref.cov, torch.Tensor):
        assert cand["has_cov"]
        assert _cov_close(cand["cov"][0, :M].cpu().numpy(), ref.cov.numpy(), 1e-4 if name != "fullcov_mc_n3" else 2e-4)
    else:
        assert not cand["has_cov"]


# ------------------------------------------------------------------------------------------ NMS / BayesOD
@pytest.mark.parametrize("tag", ["small", "large"])
def test_nms_and_bayesod_on_planted_candidates(tag):
    g = np.load(os.path.join(GOLDEN, "planted_%s.npz" % tag))
    pp = O.PathParams(cls_var=True, bbox_cov=True)
    cand = O.Candidates(torch.from_numpy(g["in_boxes"]), torch.from_numpy(g["in_cov"]), torch.from_numpy(g["in_scores"]),
                        torch.from_numpy(g["in_classes"]), torch.from_numpy(g["in_probs"]),
                        np.arange(g["in_boxes"].shape[0]), [g["in_boxes"].shape[0]])
    cd = G.cand_to_dict(cand)
    # standard NMS: survivors bit-exact; the oracle statement reproduces the reference fixture, and the
    # rescaled / clipped result (inference_utils.py:374-425) is compared after detector_postprocess
    ref_det = O.standard_nms_post(cand, pp, (720, 1280))
    assert np.array_equal(ref_det.boxes.numpy(), g["std_boxes"]) and np.array_equal(ref_det.cov.numpy(), g["std_cov"])
    ref_fin = O.detector_postprocess(ref_det, 720, 1280)
    for variant in (ops.NMS_AUTO, ops.NMS_VANILLA if tag == "large" else ops.NMS_TRICK):
        det = ops.nms_fuse(cd, 0, 0.5, 0.9, 100, (720, 1280), (720, 1280), nms_variant=variant)
        nk = int(det["keep_count"][0])
        assert nk == g["std_boxes"].shape[0]
        assert np.array_equal(det["keep"][0, :nk].cpu().numpy().astype(np.int64), ref_det.keep.numpy())
        n = int(det["count"][0])
        assert n == ref_fin.boxes.shape[0]
        assert np.array_equal(det["src"][0, :n].cpu().numpy().astype(np.int64), ref_fin.keep.numpy())
        assert np.array_equal(det["boxes"][0, :n].cpu().numpy(), ref_fin.boxes.numpy())
        assert np.array_equal(det["scores"][0, :n].cpu().numpy(), ref_fin.scores.numpy())
        assert np.array_equal(det["classes"][0, :n].cpu().numpy().astype(np.int64), ref_fin.classes.numpy())
        assert np.array_equal(det["probs"][0, :n].cpu().numpy(), ref_fin.probs.numpy())
        assert np.allclose(det["cov"][0, :n].cpu().numpy(), ref_fin.cov.numpy(), rtol=1e-6, atol=0)
    # both torchvision variants agree with their own oracle statement on the same boxes
    for variant, sl in ((ops.NMS_TRICK, slice(0, 900)), (ops.NMS_VANILLA, slice(0, None))):
        sub = O.Candidates(cand.boxes[sl], cand.cov[sl], cand.scores[sl], cand.classes[sl], cand.probs[sl],
                           cand.anchor_ids[sl], [cand.boxes[sl].shape[0]])
        if variant == ops.NMS_TRICK and sub.boxes.numel() > 4000:
            continue
        if variant == ops.NMS_VANILLA and sub.boxes.numel() <= 4000:
            continue
        det = ops.nms_fuse(G.cand_to_dict(sub), 0, 0.5, 0.9, 100, (720, 1280), (720, 1280), nms_variant=variant)
        r = O.standard_nms_post(sub, pp, (720, 1280), nms_impl="loop")
        n = int(det["keep_count"][0])
        assert np.array_equal(det["keep"][0, :n].cpu().numpy().astype(np.int64), r.keep.numpy())
    # anchor statistics (mode 2), with and without candidate covariances
    for use_cov, key in ((True, "ast_cov_"), (False, "ast_nocov_")):
        cd2 = dict(cd)
        cd2["has_cov"] = use_cov
        det = ops.nms_fuse(cd2, 2, 0.5, 0.9, 100, (720, 1280), (720, 1280))
        n = int(det["count"][0])
        assert n == g[key + "boxes"].shape[0], key
        assert np.array_equal(det["classes"][0, :n].cpu().numpy().astype(np.int64), g[key + "classes"]), key
        assert np.allclose(det["scores"][0, :n].cpu().numpy(), g[key + "scores"], rtol=1e-5), key
        assert np.allclose(det["probs"][0, :n].cpu().numpy(), g[key + "probs"], rtol=1e-5, atol=1e-8), key
        assert np.allclose(det["boxes"][0, :n].cpu().numpy(), g[key + "boxes"], rtol=1e-6, atol=1e-3), key
        assert _cov_close(det["cov"][0, :n].cpu().numpy(), g[key + "cov"], 1e-4), key
    # BayesOD, all merge-mode combinations
    for cm, ck in (("max_score", "ms"), ("bayesian_inference", "avg")):
        for bm, bk in (("bayesian_inference", "bi"), ("covariance_intersection", "ci")):
            pp.cls_merge, pp.box_merge = cm, bm
            det = ops.nms_fuse(cd, 1, 0.5, 0.9, 100, (720, 1280), (720, 1280), box_merge=0 if bk == "bi" else 1,
                               cls_merge=0 if ck == "ms" else 1)
            n = int(det["count"][0])
            key = "bod_%s_%s_" % (ck, bk)
            assert n == g[key + "boxes"].shape[0], key
            r64 = O.detector_postprocess(O.bayes_od_post(cand, pp, (720, 1280), dtype=np.float64), 720, 1280)
            assert np.array_equal(det["classes"][0, :n].cpu().numpy().astype(np.int64), g[key + "classes"]), key
            assert np.allclose(det["scores"][0, :n].cpu().numpy(), g[key + "scores"], rtol=1e-5), key
            assert np.allclose(det["probs"][0, :n].cpu().numpy(), g[key + "probs"], rtol=1e-5, atol=1e-8), key
            assert np.allclose(det["boxes"][0, :n].cpu().numpy(), r64.boxes.numpy(), rtol=1e-5, atol=1e-3), key
            assert _cov_close(det["cov"][0, :n].cpu().numpy(), r64.cov.numpy(), 1e-5), key
            # against the fp32 reference fixture: the measured envelope (profiles/r2a_bayesod_envelope.txt) is 4.9e-4 px /
            # 1.9e-6, all of it the reference's own fp32-LAPACK rounding (the same distance separates it from fp64)
            assert np.allclose(det["boxes"][0, :n].cpu().numpy(), g[key + "boxes"], rtol=1e-6, atol=2e-3), key
            assert _cov_close(det["cov"][0, :n].cpu().numpy(), g[key + "cov"], 2e-5), key


def _match_detections(got_boxes, ref_boxes, tol_px):
    """Greedy nearest-box pairing; returns (pairs, unmatched_got, unmatched_ref)."""
    used, pairs = set(), []
    for i, b in enumerate(got_boxes):
        d = np.abs(ref_boxes - b[None]).max(1)
        for j in np.argsort(d):
            if d[j] > tol_px:
                break
            if j not in used:
                used.add(j)
                pairs.append((i, int(j)))
                break
    return pairs, len(got_boxes) - len(pairs), len(ref_boxes) - len(pairs)


def test_cluster_merge_on_oracle_runs():
    """Stage-isolated post-NMS merge (inference_utils.py:165-289): identical per-run detections in, the
    clustered + re-NMSed result must reproduce the oracle."""
    g = torch.Generator().manual_seed(17)
    pp = O.PathParams(cls_var=True, bbox_cov=True)
    runs = []
    for r in range(4):
        boxes, cov, scores, classes, probs = S.make_planted_candidates(30 + r if r else 30, 10, (1, 6))
        # the same ground-truth layout jittered differently per run, like MC samples of one image
        cand = O.Candidates(boxes, cov, scores, classes, probs, np.arange(boxes.shape[0]), [boxes.shape[0]])
        runs.append(O.standard_nms_post(cand, pp, (720, 1280)))
    ref = O.detector_postprocess(O.black_box_post(runs, pp, (720, 1280)), 720, 1280)
    D, K = 100, 7
    det = {"boxes": torch.zeros((4, D, 4)), "cov": torch.zeros((4, D, 4, 4)), "scores": torch.zeros((4, D)),
           "classes": torch.zeros((4, D), dtype=torch.int32), "probs": torch.zeros((4, D, K)),
           "count": torch.zeros((4,), dtype=torch.int32)}
    for r, d in enumerate(runs):
        n = d.boxes.shape[0]
        det["boxes"][r, :n] = d.boxes; det["cov"][r, :n] = d.cov; det["scores"][r, :n] = d.scores
        det["classes"][r, :n] = d.classes.int(); det["probs"][r, :n] = d.probs; det["count"][r] = n
    det = {k: v.cuda() for k, v in det.items()}
    clusters = ops.cluster_merge(det, 4, 0.9)
    out = ops.nms_fuse(clusters, 0, 0.5, 0.9, 100, (720, 1280), (720, 1280))
    n = int(out["count"][0])
    assert n == ref.boxes.shape[0]
    assert np.array_equal(out["classes"][0, :n].cpu().numpy().astype(np.int64), ref.classes.numpy())
    assert np.allclose(out["scores"][0, :n].cpu().numpy(), ref.scores.numpy(), rtol=1e-5)
    assert np.allclose(out["boxes"][0, :n].cpu().numpy(), ref.boxes.numpy(), rtol=1e-6, atol=1e-3)
    assert np.allclose(out["probs"][0, :n].cpu().numpy(), ref.probs.numpy(), rtol=1e-5, atol=1e-8)
    assert _cov_close(out["cov"][0, :n].cpu().numpy(), ref.cov.numpy(), 1e-4)


@pytest.mark.parametrize("name", ["mcdrop_post_n3", "ensembles_post_e3"])
def test_end_to_end_post_nms_merge(name):
    """Post-NMS merge modes end to end.  Three selection stages are chained (per-run NMS, IoU >= 0.9
    clustering, final NMS), so detections are paired by box proximity and a few boundary flips are allowed."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds if len(sds) > 1 else sds[0])
    res = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
    g = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    gb, rb = res.pred_boxes.tensor.cpu().numpy(), g["final_boxes"]
    pairs, ug, ur = _match_detections(gb, rb, 0.05)
    assert ug <= 3 and ur <= 3, (ug, ur, len(gb), len(rb))
    ig = np.array([p[0] for p in pairs]); ir = np.array([p[1] for p in pairs])
    assert np.array_equal(res.pred_classes.cpu().numpy()[ig], g["final_classes"][ir])
    assert np.allclose(res.scores.cpu().numpy()[ig], g["final_scores"][ir], rtol=2e-4, atol=1e-6)
    assert np.allclose(res.pred_cls_probs.cpu().numpy()[ig], g["final_probs"][ir], rtol=2e-4, atol=1e-6)
    assert _cov_close(res.pred_boxes_covariance.cpu().numpy()[ig], g["final_cov"][ir], 2e-3)


def test_nms_edge_cases():
    K = 7
    # empty candidate list
    cd = {"boxes": torch.zeros((2, 8, 4), device="cuda"), "cov": torch.zeros((2, 8, 4, 4), device="cuda"),
          "scores": torch.zeros((2, 8), device="cuda"), "classes": torch.zeros((2, 8), dtype=torch.int32, device="cuda"),
          "probs": torch.zeros((2, 8, K), device="cuda"), "count": torch.tensor([0, 3], dtype=torch.int32, device="cuda"),
          "has_cov": True}
    # image 1: IoU exactly == threshold is kept (strict '>'), equal scores -> lower index first,
    # zero-area box neither suppresses nor is suppressed, clipped-away box is dropped by nonempty()
    cd["boxes"][1, 0] = torch.tensor([0.0, 0.0, 10.0, 10.0])
    cd["boxes"][1, 1] = torch.tensor([0.0, 0.0, 10.0, 5.0])       # IoU with box 0 = 0.5 exactly
    cd["boxes"][1, 2] = torch.tensor([-30.0, -30.0, -20.0, -20.0])  # entirely outside -> empty after clip
    cd["scores"][1, :3] = torch.tensor([0.9, 0.9, 0.8])
    cd["probs"][1, :3, 0] = cd["scores"][1, :3]
    cd["cov"][1, :3] = torch.eye(4, device="cuda")
    det = ops.nms_fuse(cd, 0, 0.5, 0.9, 100, (100, 100), (100, 100))
    assert det["count"].cpu().tolist() == [0, 2]
    assert det["keep_count"].cpu().tolist() == [0, 3]
    assert det["keep"][1, :3].cpu().tolist() == [0, 1, 2]
    ref = O.Candidates(cd["boxes"][1, :3].cpu(), cd["cov"][1, :3].cpu(), cd["scores"][1, :3].cpu(),
                       cd["classes"][1, :3].cpu().long(), cd["probs"][1, :3].cpu(), np.arange(3), [3])
    r = O.detector_postprocess(O.standard_nms_post(ref, O.PathParams(), (100, 100)), 100, 100)
    assert np.array_equal(det["boxes"][1, :2].cpu().numpy(), r.boxes.numpy())
    assert np.allclose(det["cov"][1, :2].cpu().numpy(), r.cov.numpy(), rtol=1e-6)
    # rescale to another output resolution
    det = ops.nms_fuse(cd, 0, 0.5, 0.9, 100, (100, 100), (50, 200))
    r = O.detector_postprocess(O.standard_nms_post(ref, O.PathParams(), (100, 100)), 50, 200)
    assert np.array_equal(det["boxes"][1, :2].cpu().numpy(), r.boxes.numpy())
    assert np.allclose(det["cov"][1, :2].cpu().numpy(), r.cov.numpy(), rtol=1e-6)


# ------------------------------------------------------------------------------------------ whole path
NEAR_TIE = 2e-5     # score gaps below this may legitimately reorder (head differs from fp32 CPU by ~1e-6)


def _align_by_id(ids_got, ids_ref, scores_ref, boundary_scores, group_of=None):
    """Candidates / detections are identified by their global anchor id.  Returns index arrays
    (ig, ir) pairing the common ids, after checking that any id present on one side only sits on a
    selection boundary (score within NEAR_TIE of the 0.05 threshold or of a level's k-th score), and that
    within a group (FPN level: candidates are level-major, descending score inside a level) the GPU order
    differs from the reference's only among near-equal scores."""
    ids_got, ids_ref = np.asarray(ids_got, np.int64), np.asarray(ids_ref, np.int64)
    assert len(set(ids_got.tolist())) == len(ids_got)
    pos_ref = {int(a): i for i, a in enumerate(ids_ref)}
    pos_got = {int(a): i for i, a in enumerate(ids_got)}
    for a in set(pos_ref) ^ set(pos_got):
        sc = boundary_scores(a)
        assert sc is not None and sc <= NEAR_TIE, "anchor %d selected on one side only (margin %s)" % (a, sc)
    common = [a for a in ids_ref.tolist() if a in pos_got]
    ig = np.array([pos_got[a] for a in common], dtype=np.int64)
    ir = np.array([pos_ref[a] for a in common], dtype=np.int64)
    # relative order may differ only among near-equal scores: walking the common candidates in the GPU's order, the
    # reference scores must be non-increasing (up to NEAR_TIE) inside every group, and groups must not interleave
    sr = np.asarray(scores_ref, np.float64)[ir]
    order_got = np.argsort(ig, kind="stable")
    grp = np.array([group_of(int(a)) for a in common], dtype=np.int64) if group_of is not None else np.zeros(len(common), np.int64)
    g_seq, s_seq = grp[order_got], sr[order_got]
    assert (np.diff(g_seq) >= 0).all(), "candidates of different levels interleave"
    same = np.diff(g_seq) == 0
    assert (np.diff(s_seq)[same] <= NEAR_TIE).all(), "GPU candidate order differs from the reference beyond near-ties"
    return ig, ir


def _compare_path(res, cand, det, ref_final, ref_cand, ref_det, pp, bayes, prob_atol=1e-7, prob_rtol=1e-4, box_atol=2e-3):
    M = int(cand["count"][0])
    ids_got = cand["anchor"][0, :M].cpu().numpy()
    sizes = np.cumsum([0] + [int(x.shape[0]) for x in ref_cand.level_scores])

    def boundary(a):
        lvl = int(np.searchsorted(sizes, a, side="right") - 1)
        sc = ref_cand.level_scores[lvl]
        v = float(sc[a - sizes[lvl]])
        k = min(pp.topk, sc.shape[0])
        kth = float(torch.sort(sc, descending=True)[0][k - 1])
        return min(abs(v - pp.score_thresh), abs(v - kth))

    level_of = lambda a: int(np.searchsorted(sizes, a, side="right") - 1)
    ig, ir = _align_by_id(ids_got, ref_cand.anchor_ids, ref_cand.scores.numpy(), boundary, group_of=level_of)
    assert len(ig) >= 0.98 * len(ref_cand.anchor_ids)
    g = lambda k: cand[k][0, :M].cpu().numpy()[ig]
    assert np.array_equal(g("classes").astype(np.int64), ref_cand.classes.numpy()[ir])
    assert _close(g("scores"), ref_cand.scores.numpy()[ir], prob_rtol, prob_atol, "candidate scores")
    assert _close(g("probs"), ref_cand.probs.numpy()[ir], prob_rtol, prob_atol, "candidate probability vectors")
    assert _close(g("boxes"), ref_cand.boxes.numpy()[ir], 1e-4, box_atol, "candidate boxes")
    if isinstance(ref_cand.cov, torch.Tensor):
        assert _cov_close(g("cov"), ref_cand.cov.numpy()[ir], 2e-4)
    # detections: identified by the anchor id of the NMS survivor they come from
    nk = int(det["keep_count"][0])
    keep_ids_got = ids_got[det["keep"][0, :nk].cpu().numpy()]
    keep_ids_ref = ref_cand.anchor_ids[ref_det.keep.numpy()]
    if len(ig) == len(ref_cand.anchor_ids) == M:
        assert set(keep_ids_got.tolist()) == set(keep_ids_ref.tolist())
    n = len(res)
    assert abs(n - ref_final.boxes.shape[0]) <= (0 if len(ig) == M else 2)
    fin_ids_got = ids_got[det["src"][0, :n].cpu().numpy()]
    fin_ids_ref = ref_cand.anchor_ids[ref_final.keep.numpy()]
    pos = {int(a): i for i, a in enumerate(fin_ids_ref)}
    sel = [(i, pos[int(a)]) for i, a in enumerate(fin_ids_got) if int(a) in pos]
    assert len(sel) >= max(n, len(fin_ids_ref)) - (0 if len(ig) == M else 2)
    i_g = np.array([x[0] for x in sel], dtype=np.int64)
    i_r = np.array([x[1] for x in sel], dtype=np.int64)
    assert np.array_equal(res.pred_classes.cpu().numpy()[i_g], ref_final.classes.numpy()[i_r])
    assert _close(res.scores.cpu().numpy()[i_g], ref_final.scores.numpy()[i_r], prob_rtol, prob_atol, "scores")
    assert _close(res.pred_cls_probs.cpu().numpy()[i_g], ref_final.probs.numpy()[i_r], prob_rtol, prob_atol, "probability vectors")
    # BayesOD end to end: measured worst case 7.8e-4 px / 1.4e-5 at cond(sum P) = 4e2 (profiles/r2a_bayesod_envelope.txt);
    # the fused covariance is bounded at the north-star's 1e-4 (scaled by the matrix' largest entry), boxes at 3e-3 px
    assert _close(res.pred_boxes.tensor.cpu().numpy()[i_g], ref_final.boxes.numpy()[i_r], 1e-4, max(box_atol, 3e-3) if bayes else box_atol, "boxes")
    assert _cov_close(res.pred_boxes_covariance.cpu().numpy()[i_g], ref_final.cov.numpy()[i_r], 1e-4 if bayes else 2e-4)
    # output order: descending score up to near-ties
    sc = res.scores.cpu().numpy()
    if not bayes or pp.cls_merge == "max_score":
        assert (np.diff(sc) <= NEAR_TIE).all()


def _check_final(inst, g, bayes):
    """Direct (position-wise) comparison with a reference fixture; valid when no near-tie reorders."""
    n = len(inst)
    assert n == g["final_boxes"].shape[0]
    assert np.array_equal(inst.pred_classes.cpu().numpy(), g["final_classes"])
    assert np.allclose(inst.scores.cpu().numpy(), g["final_scores"], rtol=1e-4, atol=1e-7)
    assert np.allclose(inst.pred_cls_probs.cpu().numpy(), g["final_probs"], rtol=1e-4, atol=1e-7)
    assert np.allclose(inst.pred_boxes.tensor.cpu().numpy(), g["final_boxes"], rtol=1e-4, atol=3e-3 if bayes else 2e-3)
    assert _cov_close(inst.pred_boxes_covariance.cpu().numpy(), g["final_cov"], 1e-4 if bayes else 2e-4)


@pytest.mark.parametrize("name", [n for n in C.CASES if not C.is_post_nms(n)])
def test_end_to_end_matches_oracle(name):
    """features -> build_predictor(cfg).infer_from_features -> Instances, against the oracle
    (oracle/podref.py, pinned bit-for-bit to the reference fixtures by tests/test_oracle_golden.py)
    on the same seeded inputs.  Candidates and detections are paired by anchor id because scores
    closer than the head's ~1e-6 numerical difference may legitimately swap positions."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds if len(sds) > 1 else sds[0])
    res, _, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_candidates=True)
    torch.set_num_threads(8)
    hws = [O.unpack_head(sd, pp) for sd in sds]
    ref_final, ref_cand, ref_det = O.predict(feats, hws, pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed, image=img,
                                             return_candidates=True, keep_diag=True, mc_single=C.is_mc_single(name))
    # the oracle here reproduces the committed reference fixture
    g = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    assert np.allclose(ref_final.boxes.numpy(), g["final_boxes"], rtol=1e-4, atol=1e-3)
    _compare_path(res[0], cand, det, ref_final, ref_cand, ref_det, pp, mode in ("bayes_od", "anchor_statistics"))


@pytest.mark.parametrize("name", ["mcdrop_pre_n4", "droponly_pre_n3", "bayesod_mc_n3", "ensembles_e3"])
def test_unread_last_sample_outputs_are_skipped_without_changing_results(name):
    """The reference never reads box_cls / box_cls_var / box_reg_var of the last MC sample / ensemble member
    (probabilistic_inference.py:216-267 loops over range(len-1)).  Leaving out the tower passes that
    feed only those outputs (predictor.skip_unread_outputs, default on) must give bit-identical detections, must
    leave every output that IS read bit-identical, and must really skip the work (fewer maps evaluated)."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    tri = lambda fs: [torch.cat([f, f * 0.5, f * 1.5], 0) for f in fs]   # batch of 3: group indexing per image
    feats3 = [tri(fs) for fs in feats] if isinstance(feats[0], (list, tuple)) else tri(feats)
    out = {}
    for skip in (True, False):
        pred = build_predictor(cfg)
        pred.load_weight_sets(sds if len(sds) > 1 else sds[0])
        pred.skip_unread_outputs = skip
        ops.PROFILE = []
        try:
            res, raw, cand, det = pred.infer_from_features(feats3, hw, out_hw, image0=img, seed=seed, return_raw=True)
            torch.cuda.synchronize()
            flop = sum(f for (_, _, f, tag) in ops.PROFILE if tag in ("tower256", "out"))
        finally:
            ops.PROFILE = None
        out[skip] = (res, {k: (v.clone() if v is not None else None) for k, v in raw.items()}, flop)
    (res_s, raw_s, flop_s), (res_f, raw_f, flop_f) = out[True], out[False]
    assert flop_s < flop_f
    S_ = raw_f["deltas"].shape[1]
    assert torch.equal(raw_s["deltas"], raw_f["deltas"])                     # every sample's deltas are read (:326-331)
    for k in ("logits", "logvar", "regvar"):
        if raw_f[k] is not None:
            assert torch.equal(raw_s[k][:, :S_ - 1], raw_f[k][:, :S_ - 1]), k
    for a, b in zip(res_s, res_f):
        assert len(a) == len(b)
        assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor)
        assert torch.equal(a.scores, b.scores)
        assert torch.equal(a.pred_classes, b.pred_classes)
        assert torch.equal(a.pred_cls_probs, b.pred_cls_probs)
        assert torch.equal(a.pred_boxes_covariance, b.pred_boxes_covariance)


@pytest.mark.parametrize("name", ["mcdrop_pre_n4", "regclsvar_std", "bayesod_mc_n3", "mcdrop_post_n3"])
def test_no_anchor_above_threshold_gives_empty_instances(name):
    """with the reference's own initialisation (cls_score bias -log(99), tiny logit weights) every score is
    ~0.01 < SCORE_THRESH_TEST = 0.05, so no anchor survives.  The whole path must return empty, well-formed
    Instances (and the oracle agrees) instead of failing on zero-sized work, in every mode and in a batch."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg = C.build_cfg(name)
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(seeds[0], num_classes=pp.num_classes, use_dropout=pp.use_dropout, cls_var=pp.cls_var,
                                bbox_cov=pp.bbox_cov, cov_dims=pp.cov_dims, logit_scale=0.01, logit_bias=-4.595)
    feats = S.make_features(0, img, hw[0], hw[1])
    feats2 = [torch.cat([f, f], 0) for f in feats]
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    res = pred.infer_from_features(feats2, hw, out_hw, image0=img, seed=seed)
    assert len(res) == 2
    for inst in res:
        assert len(inst) == 0
        assert tuple(inst.pred_boxes.tensor.shape) == (0, 4) and tuple(inst.pred_boxes_covariance.shape) == (0, 4, 4)
        assert tuple(inst.pred_cls_probs.shape) == (0, pp.num_classes) and inst.pred_classes.dtype == torch.int64
    if not C.is_post_nms(name):
        torch.set_num_threads(8)
        ref = O.predict(feats, [O.unpack_head(sd, pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed, image=img)
        assert ref.boxes.shape[0] == 0


def test_batched_equals_single_image():
    """B images in one call == B single-image calls (a batch is B independent problems)."""
    name = "mcdrop_pre_n4"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, _ = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    per_img = [S.make_features(0, i, hw[0], hw[1]) for i in range(3)]
    batch = [torch.cat([f[l] for f in per_img], 0) for l in range(5)]
    res_b = pred.infer_from_features(batch, hw, out_hw, image0=10, seed=seed)
    for i in range(3):
        r1 = pred.infer_from_features(per_img[i], hw, out_hw, image0=10 + i, seed=seed)[0]
        assert len(r1) == len(res_b[i])
        assert torch.equal(r1.pred_boxes.tensor, res_b[i].pred_boxes.tensor)
        assert torch.equal(r1.scores, res_b[i].scores)
        assert torch.equal(r1.pred_boxes_covariance, res_b[i].pred_boxes_covariance)


def test_chunked_pipelined_call_equals_one_call():
    """infer_from_features(chunk_images=c) evaluates the batch chunk by chunk (bounded activation memory) and uploads
    host-resident features of the next chunk on a copy stream meanwhile; results must be bit-identical to the
    single call, for pinned-host and for device-resident features, and for a ragged last chunk."""
    name = "mcdrop_pre_n4"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, _ = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    per_img = [S.make_features(0, i, hw[0], hw[1]) for i in range(5)]
    host = [torch.cat([f[l] for f in per_img], 0).pin_memory() for l in range(5)]
    ref, _, ref_cand, ref_det = pred.infer_from_features(host, hw, out_hw, image0=20, seed=seed, return_candidates=True)
    for feats in (host, [f.cuda() for f in host]):
        for chunk in (2, 1):
            got, _, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=20, seed=seed, return_candidates=True,
                                                         chunk_images=chunk)
            assert len(got) == len(ref) == 5
            assert torch.equal(det["count"], ref_det["count"]) and torch.equal(cand["count"], ref_cand["count"])
            for a, b in zip(got, ref):
                assert torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor)
                assert torch.equal(a.scores, b.scores)
                assert torch.equal(a.pred_classes, b.pred_classes)
                assert torch.equal(a.pred_boxes_covariance, b.pred_boxes_covariance)
    only = pred.infer_from_features(host, hw, out_hw, image0=20, seed=seed, chunk_images=3)      # list of Instances only
    assert len(only) == 5 and torch.equal(only[4].scores, ref[4].scores)
    # automatic chunking: with a budget of two images' activations the same call runs as chunks of two
    pred.max_activation_bytes = 2.5 * pred._activation_bytes_per_image(host)
    auto = pred.infer_from_features(host, hw, out_hw, image0=20, seed=seed)
    assert len(auto) == 5 and all(torch.equal(a.scores, b.scores) and torch.equal(a.pred_boxes.tensor, b.pred_boxes.tensor)
                                  for a, b in zip(auto, ref))


def test_reference_call_surface():
    """predictor(input_im) with the reference's input dict; invalid meta-architecture / mode raise
    ValueError as in probabilistic_inference.py:29-33,100-103."""
    name = "regclsvar_std"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    pred.rng_seed = seed
    input_im = [{"image": torch.zeros((3, hw[0], hw[1]), dtype=torch.uint8), "height": out_hw[0], "width": out_hw[1],
                 "image_id": img, "features": feats}]
    inst = pred(input_im)
    g = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    assert abs(len(inst) - g["final_boxes"].shape[0]) <= 2
    assert inst.has("pred_boxes_covariance") and inst.image_size == tuple(out_hw)
    bad = cfg.clone()
    bad.defrost()
    bad.MODEL.META_ARCHITECTURE = "GeneralizedRCNN"
    with pytest.raises(ValueError):
        build_predictor(bad)
    pred.inference_mode = "nonsense"
    with pytest.raises(ValueError):
        pred(input_im)


def test_full_size_path_properties_and_parity():
    """BASELINE geometry (1280x720 -> 184 140 anchors), MC-dropout N=2 so the CPU oracle finishes in
    seconds: parity of the whole path paired by anchor id, plus size-independent properties --
    scores sorted, boxes inside the image, covariances symmetric positive definite, NMS idempotent
    (re-running NMS on the survivors keeps all of them), results independent of batch position."""
    import bench
    H, W, N = 720, 1280, 2
    cfg = bench.build_cfg(N)
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(0, num_classes=7, use_dropout=True, cls_var=True, bbox_cov=True)
    per_img = [S.make_features(0, i, H, W) for i in range(2)]
    batch = [torch.cat([f[l] for f in per_img], 0) for l in range(5)]
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    res, _, cand, det = pred.infer_from_features(batch, (H, W), (H, W), image0=0, seed=3, return_candidates=True)
    assert int(cand["boxes"].shape[1]) == 4540 and int(cand["count"][0]) > 2000   # top-k binding on the big levels
    for b in range(2):
        r = res[b]
        sc = r.scores.cpu().numpy()
        assert len(r) == 100 and (np.diff(sc) <= 0).all()
        bx = r.pred_boxes.tensor.cpu().numpy()
        assert (bx[:, 0] >= 0).all() and (bx[:, 2] <= W).all() and (bx[:, 1] >= 0).all() and (bx[:, 3] <= H).all()
        cv = r.pred_boxes_covariance.cpu().numpy().astype(np.float64)
        assert np.allclose(cv, cv.transpose(0, 2, 1), rtol=1e-6, atol=1e-9)
        assert (np.linalg.eigvalsh(cv) > 0).all()
    # NMS idempotence on the survivors of image 0
    n = len(res[0])
    surv = {"boxes": det["boxes"][:1].contiguous(), "cov": det["cov"][:1].contiguous(), "scores": det["scores"][:1].contiguous(),
            "classes": det["classes"][:1].contiguous(), "probs": det["probs"][:1].contiguous(),
            "count": det["count"][:1].contiguous(), "has_cov": True}
    again = ops.nms_fuse(surv, 0, 0.5, 0.9, 100, (H, W), (H, W))
    assert int(again["keep_count"][0]) == n and again["keep"][0, :n].cpu().tolist() == list(range(n))
    # batch position independence
    r1 = pred.infer_from_features(per_img[1], (H, W), (H, W), image0=1, seed=3)[0]
    assert torch.equal(r1.pred_boxes.tensor, res[1].pred_boxes.tensor) and torch.equal(r1.scores, res[1].scores)
    # parity with the oracle for image 0
    torch.set_num_threads(os.cpu_count())
    ref_final, ref_cand, ref_det = O.predict(per_img[0], [O.unpack_head(sd, pp)], pp, "mc_dropout_ensembles", (H, W), n_mc=N,
                                             seed=3, image=0, return_candidates=True, keep_diag=True)
    one = {k: (v[:1] if isinstance(v, torch.Tensor) else v) for k, v in cand.items()}
    one_det = {k: (v[:1] if isinstance(v, torch.Tensor) else v) for k, v in det.items()}
    _compare_path(res[0], one, one_det, ref_final, ref_cand, ref_det, pp, False)


def test_predictor_from_raw_images_with_backbone():
    """predictor(input_im) from a raw image through the torch ResNet-50-FPN backbone (upstream of the
    rebuilt path) equals infer_from_features on the backbone's maps."""
    from pod_compare_b200 import backbone as BB
    name = "regclsvar_std"
    cfg, pp, sds, _ = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    pred.load_backbone(BB.random_state_dict(1))
    img = S.make_image(0, 0, 96, 160)
    inst = pred([{"image": img, "height": 96, "width": 160, "image_id": 0}])
    feats = pred.backbone([img])
    assert [tuple(f.shape[-2:]) for f in feats] == [(12, 20), (6, 10), (3, 5), (2, 3), (1, 2)]      # 96x160 is a multiple of 32
    ref, _, cand, det = pred.infer_from_features(feats, (96, 160), (96, 160), image0=0, return_candidates=True)
    assert len(inst) == len(ref[0]) and torch.equal(inst.pred_boxes.tensor, ref[0].pred_boxes.tensor)
    # the random backbone's maps are far outside the unit range (max|x| in the hundreds): the per-call activation scale
    # (engine.HeadEngine.feature_scale) keeps every tower layer inside the fp16 split range, and the result still
    # matches the fp32 oracle on the same maps
    amax = max(float(f.abs().max()) for f in feats)
    assert amax > 8.0
    opts, mode, n_mc, seeds, hw, out_hw, seed, _ = C.CASES[name]
    torch.set_num_threads(8)
    cpu_feats = [f.cpu() for f in feats]
    ref_final, ref_cand, ref_det = O.predict(cpu_feats, [O.unpack_head(sds[0], pp)], pp, mode, (96, 160), seed=pred.rng_seed,
                                             image=0, return_candidates=True, keep_diag=True)
    # These maps are ~15x larger than unit-variance features, and so are the logits (|logit| up to ~40).  A probability
    # moves by dp/p = (1 - p) * d(logit): the head's ~1e-5 relative accuracy of the largest logit is an
    # ABSOLUTE logit error of up to ~4e-4 here, i.e. up to 4e-4 relative on a probability -- the same conditioning limits
    # any fp32 evaluation, the reference's included.  Measured worst case 2.0e-4 (r2f); bound 5e-4, floor 2e-6.
    # The regression deltas are ~15x larger too (boxes move by hundreds of pixels, many clamp at exp(4.135)): the same
    # relative accuracy is 1e-2 px here instead of 2e-3 (measured worst case 2.8e-3, r2j).
    _compare_path(ref[0], cand, det, ref_final, ref_cand, ref_det, pp, False, prob_atol=2e-6, prob_rtol=5e-4, box_atol=1e-2)


def test_large_feature_magnitudes_are_rescaled():
    """Feature maps beyond the default fp16 split range (|x|*16 > 65504) get a smaller power-of-two scale
    instead of overflowing: results equal the oracle on the same (large) inputs."""
    name = "regclsvar_std"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    feats = [f * 3000.0 for f in feats]
    sd = {k: (v / 3000.0 if k.endswith(".0.weight") and "subnet" in k else v) for k, v in sds[0].items()}
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    res, _, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_candidates=True)
    torch.set_num_threads(8)
    ref_final, ref_cand, ref_det = O.predict(feats, [O.unpack_head(sd, pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed,
                                             image=img, return_candidates=True, keep_diag=True)
    assert torch.isfinite(res[0].scores).all()
    _compare_path(res[0], cand, det, ref_final, ref_cand, ref_det, pp, False)


# ------------------------------------------------------------------------------------------ product safety
def test_device_side_errors_reach_the_caller():
    """A bounded mbarrier wait that expires, or an activation outside the fp16 split range, must surface as PodError
    from the product call -- never as plausible-looking detections (the predictor polls pod_status once per call)."""
    name = "mcdrop_pre_n4"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    good = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
    # (1) fault injection: the TMA producer of the first CTA / CTA pair never loads; with a 1 ms wait budget its
    #     consumer waits expire, the kernel bails out and records the code of the wait
    ops.set_conv_wait_limit(2_000_000)
    ops.set_conv_debug_fault(True)
    try:
        with pytest.raises(PodError, match="device-side error"):
            pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)
    finally:
        ops.set_conv_debug_fault(False)
        ops.set_conv_wait_limit(0)
    assert ops.status() == 0                                   # the word was cleared by the failed call
    again = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
    assert torch.equal(again.pred_boxes.tensor, good.pred_boxes.tensor) and torch.equal(again.scores, good.scores)
    # (2) saturation: a tower layer scaled so that its activations exceed 65504 / 16 in the fp16 split pair
    sd = dict(sds[0])
    sd["head.cls_subnet.3.weight"] = sd["head.cls_subnet.3.weight"] * 3.0e4
    bad = build_predictor(cfg)
    bad.load_weight_sets(sd)
    with pytest.raises(PodError, match="fp16 split range"):
        bad.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)
    # (3) non-finite input features
    nf = [f.clone() for f in feats]
    nf[2][0, 5, 1, 1] = float("nan")
    with pytest.raises(PodError):
        pred.infer_from_features(nf, hw, out_hw, image0=img, seed=seed)
    assert ops.status() == 0
    ok = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
    assert torch.equal(ok.scores, good.scores)


def test_ensemble_members_read_their_own_feature_maps():
    """a15: feats[e][l] -- every member of the reference is a full model with its own backbone
    (probabilistic_inference.py:58-77,499-501).  Per-member maps must change the result relative to shared maps, a
    flat list must equal E copies of it, and raw member outputs must come from the member's own maps."""
    name = "ensembles_e3"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    assert isinstance(feats[0], list) and len(feats) == 3
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds)
    pred.skip_unread_outputs = False
    own, raw_own, _, _ = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_raw=True)
    shared, raw_sh, _, _ = pred.infer_from_features(feats[0], hw, out_hw, image0=img, seed=seed, return_raw=True)
    tiled, raw_t, _, _ = pred.infer_from_features([feats[0]] * 3, hw, out_hw, image0=img, seed=seed, return_raw=True)
    assert torch.equal(raw_sh["logits"], raw_t["logits"]) and torch.equal(shared[0].scores, tiled[0].scores)
    assert torch.equal(raw_own["logits"][:, 0], raw_sh["logits"][:, 0])          # member 0 reads set 0 in both
    assert not torch.equal(raw_own["logits"][:, 1], raw_sh["logits"][:, 1])      # member 1 reads its own set
    # member 1 alone on its own maps == slice 1 of the per-member evaluation
    torch.set_num_threads(8)
    hw1 = O.unpack_head(sds[1], pp)
    o1 = O.head_outputs(feats[1], hw1, pp, O.DropoutSource("off", 0.0))
    ref_logits = torch.cat([o1["box_cls"][l][0] for l in range(5)], 0)
    assert torch.allclose(raw_own["logits"][0, 1].cpu(), ref_logits, rtol=1e-4, atol=2e-5)
    with pytest.raises(Exception):
        pred.infer_from_features(feats[:2], hw, out_hw, image0=img, seed=seed)    # 2 sets for 3 members
    # the reference-style call with per-member features in the input dict: features[e][l]
    pred.rng_seed = seed
    inst = pred([{"image_hw": hw, "height": out_hw[0], "width": out_hw[1], "image_id": img, "features": feats}])
    assert len(inst) == len(own[0]) and torch.equal(inst.scores, own[0].scores)
    assert torch.equal(inst.pred_boxes_covariance, own[0].pred_boxes_covariance)
    # 'ensembles' together with MC_DROPOUT.ENABLE is not a combination the reference composes (its pre-NMS branch would
    # silently run MC-dropout on the un-loaded base model instead): an explicit error here, never a quiet guess
    cfg2 = cfg.clone()
    cfg2.defrost()
    cfg2.PROBABILISTIC_INFERENCE.MC_DROPOUT.ENABLE = True
    pred2 = build_predictor(cfg2)
    pred2.load_weight_sets(sds)
    with pytest.raises(PodError, match="ensembles"):
        pred2.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)


# ------------------------------------------------------------------------------------------ the BASELINE.json configs
_FULL = {}


def _full_size_oracle(n_mc, image, seed):
    """Oracle head outputs + anchor-wise candidates of ONE 1280x720 image at N MC samples (~26 s of host time at
    N=30); shared by the configs[2] and configs[3] tests, which differ only in the post-processing."""
    import bench
    key = (n_mc, image, seed)
    if key not in _FULL:
        cfg = bench.build_cfg(n_mc)
        pp = O.PathParams.from_cfg(cfg)
        sd = S.make_head_state_dict(0, num_classes=7, use_dropout=True, cls_var=True, bbox_cov=True)
        feats = S.make_features(0, image, 720, 1280)
        torch.set_num_threads(os.cpu_count())
        hwt = O.unpack_head(sd, pp)
        drop = O.DropoutSource("philox", pp.dropout_rate, seed, image)
        with torch.no_grad():
            outs = [O.head_outputs(feats, hwt, pp, drop, sample=s) for s in range(n_mc)]
            anchors = O.make_anchors([tuple(f.shape[-2:]) for f in feats], pp)
            cand = O.anchorwise(outs, anchors, pp, seed, image, keep_diag=True)
        _FULL[key] = (pp, sd, feats, cand)
    return _FULL[key]


@pytest.mark.parametrize("workload,mode", [("mc_pre", "mc_dropout_ensembles"), ("bayes_od_mc", "bayes_od")])
def test_baseline_config_mc_dropout_n30_full_size(workload, mode):
    """BASELINE.json configs[2] and configs[3] AS STATED: reg_cls_var_dropout head, 1280x720, N=30 MC-dropout samples,
    standard NMS (pre-NMS merge) resp. BayesOD fusion, in a batch of 17 evaluated in chunks of 16 -- the image under
    test is the one that crosses the chunk boundary.  Paired with the oracle by anchor id."""
    import bench
    N, B, H, W, seed, pos = 30, 17, 720, 1280, 5, 16
    cfg = bench.build_cfg(N, workload)
    pp_o, sd, feats_o, ref_cand = _full_size_oracle(N, pos, seed)
    pp = O.PathParams.from_cfg(cfg)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    per = [S.make_features(0, i % 3, H, W) for i in range(3)]
    batch = [torch.cat([(feats_o if i == pos else per[i % 3])[l] for i in range(B)], 0) for l in range(5)]
    res, _, cand, det = pred.infer_from_features(batch, (H, W), (H, W), image0=0, seed=seed, return_candidates=True,
                                                 chunk_images=16)
    assert len(res) == B and all(len(r) > 0 for r in res)
    if mode == "bayes_od":
        ref_det = O.bayes_od_post(ref_cand, pp, (H, W))
    else:
        ref_det = O.standard_nms_post(ref_cand, pp, (H, W))
    ref_final = O.detector_postprocess(ref_det, H, W)
    one = {k: (v[pos:pos + 1] if isinstance(v, torch.Tensor) else v) for k, v in cand.items()}
    one_det = {k: (v[pos:pos + 1] if isinstance(v, torch.Tensor) else v) for k, v in det.items()}
    assert int(one["count"][0]) > 3000                     # top-k is binding on the large levels
    _compare_path(res[pos], one, one_det, ref_final, ref_cand, ref_det, pp, mode == "bayes_od")
    # the same image evaluated alone (batch position / chunking independence at full size)
    alone = pred.infer_from_features(feats_o, (H, W), (H, W), image0=pos, seed=seed)[0]
    assert torch.equal(alone.pred_boxes.tensor, res[pos].pred_boxes.tensor) and torch.equal(alone.scores, res[pos].scores)
    assert torch.equal(alone.pred_boxes_covariance, res[pos].pred_boxes_covariance)


def test_baseline_config_ensembles_e5_full_size():
    """BASELINE.json configs[4]: reg_cls_var head, 5-member ensemble (5 weight sets, 5 feature sets: every member has
    its own backbone), pre-NMS merge, 1280x720."""
    import bench
    H, W, seed, img = 720, 1280, 6, 2
    cfg = bench.build_cfg(1, "ensembles5")
    pp = O.PathParams.from_cfg(cfg)
    sds = S.make_member_state_dicts(5, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True)   # as bench.py
    feats = S.make_member_features(5, img, H, W)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds)
    two = [[torch.cat([f, f.flip(-1)], 0) for f in fs] for fs in feats]          # batch of 2, image under test first
    res, _, cand, det = pred.infer_from_features(two, (H, W), (H, W), image0=img, seed=seed, return_candidates=True)
    torch.set_num_threads(os.cpu_count())
    ref_final, ref_cand, ref_det = O.predict(feats, [O.unpack_head(sd, pp) for sd in sds], pp, "ensembles", (H, W), seed=seed,
                                             image=img, return_candidates=True, keep_diag=True)
    assert ref_cand.boxes.shape[0] > 500
    one = {k: (v[:1] if isinstance(v, torch.Tensor) else v) for k, v in cand.items()}
    one_det = {k: (v[:1] if isinstance(v, torch.Tensor) else v) for k, v in det.items()}
    _compare_path(res[0], one, one_det, ref_final, ref_cand, ref_det, pp, False)


def test_baseline_config_loss_attenuation_batch8_full_size():
    """BASELINE.json configs[1]: reg_cls_var head (loss attenuation), standard NMS, single forward, batch 8 at
    1280x720; two of the eight images are checked against the oracle."""
    import bench
    H, W, seed = 720, 1280, 7
    cfg = bench.build_cfg(1, "loss_att")
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(0, num_classes=7, use_dropout=False, cls_var=True, bbox_cov=True)
    per = [S.make_features(0, i, H, W) for i in range(8)]
    batch = [torch.cat([f[l] for f in per], 0) for l in range(5)]
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    res, _, cand, det = pred.infer_from_features(batch, (H, W), (H, W), image0=40, seed=seed, return_candidates=True)
    assert len(res) == 8
    torch.set_num_threads(os.cpu_count())
    for b in (0, 7):
        ref_final, ref_cand, ref_det = O.predict(per[b], [O.unpack_head(sd, pp)], pp, "standard_nms", (H, W), seed=seed,
                                                 image=40 + b, return_candidates=True, keep_diag=True)
        one = {k: (v[b:b + 1] if isinstance(v, torch.Tensor) else v) for k, v in cand.items()}
        one_det = {k: (v[b:b + 1] if isinstance(v, torch.Tensor) else v) for k, v in det.items()}
        _compare_path(res[b], one, one_det, ref_final, ref_cand, ref_det, pp, False)


# ------------------------------------------------------------------------------------------ wire format (f3)
def _instances_from_golden(g, out_hw):
    from pod_compare_b200.structures import Boxes, Instances
    inst = Instances((int(out_hw[0]), int(out_hw[1])))
    inst.pred_boxes = Boxes(torch.from_numpy(g["final_boxes"]))
    inst.scores = torch.from_numpy(g["final_scores"])
    inst.pred_classes = torch.from_numpy(g["final_classes"])
    inst.pred_cls_probs = torch.from_numpy(g["final_probs"])
    inst.pred_boxes_covariance = torch.from_numpy(g["final_cov"])
    return inst


@pytest.mark.parametrize("name", ["regclsvar_std", "bayesod_plain", "baseline_std", "mcdrop_pre_n4"])
def test_wire_format_equals_reference_json(name):
    """Stage-isolated: the reference's final Instances (fixture) through the batched GPU writer must give EXACTLY the
    entries the reference's own instances_to_json wrote (tests/golden/json_*.json, oracle/make_golden.py): same order,
    same filtering by the category mapping, bit-identical floats (XYWH boxes, T Sigma T^T)."""
    import json
    from pod_compare_b200 import wire
    from pod_compare_b200.inference_utils import covar_xyxy_to_xywh, instances_to_json
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    g = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    want = json.load(open(os.path.join(GOLDEN, "json_%s.json" % name)))
    inst = _instances_from_golden(g, out_hw)
    bdd, kitti = wire.BDD_THING_DATASET_ID_TO_CONTIGUOUS_ID, wire.KITTI_THING_DATASET_ID_TO_CONTIGUOUS_ID
    maps = {"bdd": wire.build_category_mapping("bdd_train", "bdd_val", bdd, bdd),
            "kitti": wire.build_category_mapping("bdd_train", "kitti_val", bdd, kitti)}
    for key, m in maps.items():
        got = instances_to_json(inst, 1000 + img, m)
        assert got == want[key], key
        assert json.loads(json.dumps(got)) == want[key]
    # a batch of three (middle image empty) through one writer call
    from pod_compare_b200.structures import Boxes, Instances
    empty = Instances((int(out_hw[0]), int(out_hw[1])))
    empty.pred_boxes = Boxes(torch.zeros((0, 4))); empty.scores = torch.zeros((0,)); empty.pred_classes = torch.zeros((0,), dtype=torch.int64)
    empty.pred_cls_probs = torch.zeros((0, 7)); empty.pred_boxes_covariance = torch.zeros((0, 4, 4))
    w = wire.BatchJsonWriter(7, 100, maps["bdd"], "cuda")
    out = w.to_json(wire.det_from_instances([inst, empty, inst], 7, 100, "cuda"), [1000 + img, 5, 7])
    n = len(want["bdd"])
    assert out[:n] == want["bdd"] and len(out) == 2 * n and all(r["image_id"] == 7 for r in out[n:])
    ref_cov = O.covar_xyxy_to_xywh(torch.from_numpy(g["final_cov"]))
    assert torch.equal(covar_xyxy_to_xywh(torch.from_numpy(g["final_cov"])).cpu(), ref_cov)
    # round trip through the reader's transform (evaluation_utils.py:28-69): xywh -> xyxy boxes and covariances come back
    boxes, probs, covs = O.read_results_json(want["bdd"])
    k = 1000 + img
    assert torch.allclose(boxes[k], torch.from_numpy(g["final_boxes"]), rtol=1e-6, atol=1e-4)
    assert torch.equal(probs[k], torch.from_numpy(g["final_probs"]))
    scale = float(np.abs(g["final_cov"]).max()) if g["final_cov"].size else 1.0
    assert torch.allclose(covs[k], torch.from_numpy(g["final_cov"]), rtol=1e-5, atol=1e-6 * scale)


def test_predict_batch_json_end_to_end():
    """predictor.predict_batch_json == the reference's harness loop body (src/apply_net.py:88-98) for a batch: entries of
    every image equal instances_to_json of that image's Instances, and match the reference's JSON fixture within the
    head's numerical tolerance."""
    import json
    from pod_compare_b200 import wire
    from pod_compare_b200.inference_utils import instances_to_json
    name = "regclsvar_std"
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    pred.rng_seed = seed
    bdd = wire.BDD_THING_DATASET_ID_TO_CONTIGUOUS_ID
    cmap = wire.build_category_mapping("bdd_train", "bdd_val", bdd, bdd)
    other = S.make_features(0, img + 1, hw[0], hw[1])
    inputs = [[{"image_hw": hw, "height": out_hw[0], "width": out_hw[1], "image_id": img + i, "features": f}]
              for i, f in enumerate((feats, other))]
    entries = pred.predict_batch_json(inputs, cmap)
    insts = pred.predict_batch(inputs)
    per_image = [instances_to_json(inst, img + i, cmap) for i, inst in enumerate(insts)]
    assert entries == per_image[0] + per_image[1]
    want = json.load(open(os.path.join(GOLDEN, "json_%s.json" % name)))["bdd"]
    mine = [e for e in entries if e["image_id"] == img]
    assert abs(len(mine) - len(want)) <= 2
    wb = np.array([e["bbox"] for e in want])
    hits = 0
    for e in mine:
        d = np.abs(wb - np.array(e["bbox"])[None]).max(1)
        if d.min() > 0.05:
            continue
        r = want[int(d.argmin())]
        hits += 1
        assert e["category_id"] == r["category_id"]
        assert np.allclose(e["bbox"], r["bbox"], rtol=1e-4, atol=2e-3)
        assert np.allclose(e["cls_prob"], r["cls_prob"], rtol=1e-4, atol=1e-7)
        sc = np.abs(np.array(r["bbox_covar"])).max()
        assert np.allclose(e["bbox_covar"], r["bbox_covar"], rtol=0, atol=2e-4 * sc)
    assert hits >= len(want) - 3


# ------------------------------------------------------------------------------------------ fused Q1 sample accumulation
@pytest.mark.parametrize("how", ["stream", "epilogue"])
@pytest.mark.parametrize("name", ["mcdrop_pre_n4", "droponly_pre_n3", "bayesod_mc_n3", "fullcov_mc_n3"])
def test_fused_sample_mean_equals_per_sample_evaluation(name, how):
    """head_mc(fuse_q1): the last tower layer accumulates the reference's weighted sample sum in the conv epilogue and
    cls_score / cls_var / bbox_cov run once per image on the mean activation.  mean_s head(x_s) == head(mean_s x_s) for
    a linear head, so the Q1 means must agree with the per-sample evaluation to fp32 round-off, per-sample deltas must be
    bit-identical, fewer output convolutions must run -- and the detections must still match the oracle."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    feats3 = [torch.cat([f, f * 0.5, f * 1.5], 0) for f in feats]
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    pred.fuse_sample_mean = how
    eng = pred._engine
    dev = [f.cuda().contiguous() for f in feats3]
    out = {}
    for fuse in (False, how):
        ops.PROFILE = []
        try:
            raw, level_off = eng.head_mc(dev, n_mc, seed, img, skip_unread=True, fuse_q1=fuse)
            torch.cuda.synchronize()
            flop = sum(f for (_, _, f, tag) in ops.PROFILE if tag == "out")
        finally:
            ops.PROFILE = None
        assert ops.status() == 0
        out[fuse] = ({k: (v.clone() if v is not None else None) for k, v in raw.items()}, flop)
    (raw_u, flop_u), (raw_f, flop_f) = out[False], out[how]
    assert flop_f < flop_u
    assert torch.equal(raw_f["deltas"], raw_u["deltas"])                       # per-sample deltas: same kernels, same bits
    for k in ("logits", "logvar", "regvar"):
        if raw_u[k] is None:
            assert raw_f[k] is None
            continue
        assert raw_f[k].shape[1] == 1 and raw_u[k].shape[1] == n_mc
        want = ops.sample_mean_q1(raw_u[k])
        got = raw_f[k][:, 0]
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 2e-6 * scale, k             # fp32 round-off of a different summation order
    # and end to end against the oracle, through the default (fused) product path
    assert pred.fuse_sample_mean == how and pred.skip_unread_outputs
    res, _, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_candidates=True)
    torch.set_num_threads(8)
    ref_final, ref_cand, ref_det = O.predict(feats, [O.unpack_head(sds[0], pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc,
                                             seed=seed, image=img, return_candidates=True, keep_diag=True)
    _compare_path(res[0], cand, det, ref_final, ref_cand, ref_det, pp, mode in ("bayes_od", "anchor_statistics"))
    # batch composition does not change a single bit (fixed accumulation group size)
    solo = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed)[0]
    trio = pred.infer_from_features(feats3, hw, out_hw, image0=img, seed=seed)[0]
    assert torch.equal(solo.scores, trio.scores) and torch.equal(solo.pred_boxes.tensor, trio.pred_boxes.tensor)
    assert torch.equal(solo.pred_boxes_covariance, trio.pred_boxes_covariance)


def test_fused_sample_mean_with_more_samples_than_one_group():
    """N = 19 samples span three accumulation groups of 8 (partial sums 8 + 8 + 2 of the 18 averaged samples); odd tile
    counts per map make units of different length share a CTA pair."""
    name = "mcdrop_pre_n4"
    opts, mode, _, seeds, _, _, seed, img = C.CASES[name]
    cfg = C.build_cfg(name)
    cfg.defrost()
    cfg.PROBABILISTIC_INFERENCE.MC_DROPOUT.NUM_RUNS = 19
    cfg.freeze()
    pp = O.PathParams.from_cfg(cfg)
    sd = S.make_head_state_dict(seeds[0], num_classes=pp.num_classes, use_dropout=True, cls_var=True, bbox_cov=True)
    hw = (72, 104)                                          # level maps 9x13, 5x7, 3x4, 2x2, 1x1: ragged tiles everywhere
    feats = [torch.randn((2, 256, h, w), generator=torch.Generator().manual_seed(5 + i)) for i, (h, w) in
             enumerate([(9, 13), (5, 7), (3, 4), (2, 2), (1, 1)])]
    pred = build_predictor(cfg)
    pred.load_weight_sets(sd)
    eng = pred._engine
    dev = [f.cuda().contiguous() for f in feats]
    raw_u, _ = eng.head_mc(dev, 19, seed, img, skip_unread=True, fuse_q1=False)
    raw_u = {k: v.clone() for k, v in raw_u.items()}
    raw_f, _ = eng.head_mc(dev, 19, seed, img, skip_unread=True, fuse_q1="epilogue")
    raw_f = {k: v.clone() for k, v in raw_f.items()}
    raw_s, _ = eng.head_mc(dev, 19, seed, img, skip_unread=True, fuse_q1="stream")
    torch.cuda.synchronize()
    assert ops.status() == 0
    assert torch.equal(raw_f["deltas"], raw_u["deltas"])
    assert torch.equal(raw_s["deltas"], raw_u["deltas"])
    for k in ("logits", "logvar", "regvar"):
        want = ops.sample_mean_q1(raw_u[k])
        for got in (raw_f[k][:, 0], raw_s[k][:, 0]):
            assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max()), k


# ------------------------------------------------------------------------------------------ in-kernel input dropout
@pytest.mark.parametrize("name", ["mcdrop_pre_n4", "droponly_pre_n3", "mcdrop_single"])
def test_in_kernel_input_masking_is_bit_identical_to_mask_replication(name):
    """pod_conv_args.mask_in: the first masked tower layer applies the dropout mask of every (sample, pass) to its staged
    input tile in shared memory (warps 2-3 of the CTA-pair kernel) instead of reading N x passes masked copies written by
    pod_mask_expand_split.  Same operands into the same MMAs: every raw output must be bit-identical, for ragged tiles,
    odd tile counts, batches, the unread-pass selection and single-run dropout."""
    opts, mode, n_mc, seeds, hw, out_hw, seed, img = C.CASES[name]
    cfg, pp, sds, feats = _oracle_case(name)
    pred = build_predictor(cfg)
    pred.load_weight_sets(sds[0])
    eng = pred._engine
    shapes = [(13, 21), (8, 16), (5, 7), (2, 3), (1, 1)]
    dev = [torch.randn((3, 256, h, w), generator=torch.Generator().manual_seed(40 + i)).cuda().contiguous() for i, (h, w) in enumerate(shapes)]
    out = {}
    for inside in (False, True):
        eng.mask_in_kernel = inside
        for skip in ((False, True) if n_mc > 1 else (False,)):
            raw, _ = eng.head_mc(dev, max(n_mc, 1), seed, img, skip_unread=skip, fuse_q1=False)
            torch.cuda.synchronize()
            assert ops.status() == 0
            out[(inside, skip)] = {k: (v.clone() if v is not None else None) for k, v in raw.items()}
    for skip in ((False, True) if n_mc > 1 else (False,)):
        a, b = out[(False, skip)], out[(True, skip)]
        live = max(n_mc, 1) - (1 if skip else 0)
        assert torch.equal(a["deltas"], b["deltas"])
        for k in ("logits", "logvar", "regvar"):
            if a[k] is not None:
                assert torch.equal(a[k][:, :live], b[k][:, :live]), (k, skip)
    # and through the product path (fused sample mean, in-kernel masking still switched on) at the fixture geometry,
    # against the oracle
    res, _, cand, det = pred.infer_from_features(feats, hw, out_hw, image0=img, seed=seed, return_candidates=True)
    torch.set_num_threads(8)
    ref_final, ref_cand, ref_det = O.predict(feats, [O.unpack_head(sds[0], pp)], pp, mode, hw, out_hw=out_hw, n_mc=n_mc, seed=seed,
                                             image=img, return_candidates=True, keep_diag=True, mc_single=C.is_mc_single(name))
    _compare_path(res[0], cand, det, ref_final, ref_cand, ref_det, pp, False)
