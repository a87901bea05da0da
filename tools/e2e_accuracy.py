"""End-to-end accuracy at the benchmark geometry (1280x720, reg_cls_var_dropout, MC-dropout N small
enough for the CPU oracle) for accumulation-chunk / truncation-compensation settings of the tcgen05 convolution.
    python tools/e2e_accuracy.py [N]        (needs a B200; the oracle runs on the host cores)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from oracle import podref as O
from pod_compare_b200 import ops, synthetic as S
from pod_compare_b200.predictor import build_predictor

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H, W = 720, 1280
cfg = bench.build_cfg(N)
pp = O.PathParams.from_cfg(cfg)
sd = S.make_head_state_dict(0, num_classes=7, use_dropout=True, cls_var=True, bbox_cov=True)
feats = S.make_features(0, 0, H, W)
torch.set_num_threads(os.cpu_count())
t0 = time.time()
ref_final, ref_cand, ref_det = O.predict(feats, [O.unpack_head(sd, pp)], pp, "mc_dropout_ensembles", (H, W), n_mc=N, seed=7,
                                         image=0, return_candidates=True, keep_diag=True)
drop = O.DropoutSource("philox", pp.dropout_rate, 7, 0)
outs = [O.head_outputs(feats, O.unpack_head(sd, pp), pp, drop, sample=s) for s in range(N)]
print("oracle: %.1f s, %d candidates, %d detections" % (time.time() - t0, ref_cand.boxes.shape[0], ref_final.boxes.shape[0]))
pred = build_predictor(cfg)
pred.skip_unread_outputs = False      # the raw outputs of every sample are inspected below
pred.load_weight_sets(sd)
ids_ref = {int(a): i for i, a in enumerate(ref_cand.anchor_ids)}
for kb, comp in ((6, 0.0), (6, 0.27), (9, 0.27), (12, 0.0), (12, 0.27), (18, 0.27), (36, 0.0), (36, 0.27)):
    ops.set_conv_chunk_kblocks(kb)                 # K-blocks of 64 channels per TMEM chain (4 = one tap, 36 = single chain)
    ops.set_conv_trunc_comp(comp)                  # truncation compensation, ulps per MMA accumulation
    taps = -kb
    res, raw, cand, det = pred.infer_from_features(feats, (H, W), (H, W), image0=0, seed=7, return_raw=True)
    torch.cuda.synchronize()
    dl = max(float((raw["logits"][0, s].cpu() - torch.cat([o[0] for o in outs[s]["box_cls"]], 0)).abs().max()) for s in range(N))
    dd = max(float((raw["deltas"][0, s].cpu() - torch.cat([o[0] for o in outs[s]["box_delta"]], 0)).abs().max()) for s in range(N))
    M = int(cand["count"][0])
    ids = cand["anchor"][0, :M].cpu().numpy()
    common = [(i, ids_ref[int(a)]) for i, a in enumerate(ids) if int(a) in ids_ref]
    ig = np.array([c[0] for c in common]); ir = np.array([c[1] for c in common])
    sc = cand["scores"][0, :M].cpu().numpy()[ig]; scr = ref_cand.scores.numpy()[ir]
    bx = cand["boxes"][0, :M].cpu().numpy()[ig]; bxr = ref_cand.boxes.numpy()[ir]
    cv = cand["cov"][0, :M].cpu().numpy()[ig].astype(np.float64); cvr = ref_cand.cov.numpy()[ir].astype(np.float64)
    scale = np.abs(cvr).reshape(len(ir), -1).max(1).reshape(-1, 1, 1)
    print("chunk = %s, comp %.2f: max|dlogit| %.2e  max|ddelta| %.2e | candidates %d/%d common | score rel %.2e | box abs %.2e px | "
          "cov rel(matrix max) %.2e | detections %d vs %d"
          % (("%d tap(s)" % taps) if taps > 0 else ("%d K-blocks" % -taps), comp, dl, dd, len(common), ref_cand.boxes.shape[0], float(np.abs(sc / scr - 1).max()), float(np.abs(bx - bxr).max()),
             float((np.abs(cv - cvr) / scale).max()), len(res[0]), ref_final.boxes.shape[0]))
ops.set_conv_trunc_comp(0.27)
ops.set_conv_chunk_kblocks(12)
