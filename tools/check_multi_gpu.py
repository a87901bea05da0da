"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_multi_gpu.py
A global batch of images is sharded contiguously over the ranks, every rank runs the full path on its
slice, the fixed-size detection records are all-gathered (NCCL), and rank 0 verifies that the gathered
result equals the single-GPU result for the same global batch, bit for bit (the noise streams are keyed
by the global image id, so sharding must not change any output)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from pod_compare_b200 import distributed as D, synthetic as S
from pod_compare_b200.predictor import build_predictor

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
D.init_from_env("nccl")
H, W, N, G = 192, 320, 4, 8                       # global batch of 8 small images, MC-dropout N=4
cfg = bench.build_cfg(N, "bayes_od_mc")
pred = build_predictor(cfg)
pred.load_weight_sets(S.make_head_state_dict(0, num_classes=7, use_dropout=True, cls_var=True, bbox_cov=True))
per = [S.make_features(0, i, H, W) for i in range(G)]


def run(lo, hi):
    feats = [torch.cat([per[i][l] for i in range(lo, hi)], 0) for l in range(5)]
    _, _, _, det = pred.infer_from_features(feats, (H, W), (H, W), image0=lo, seed=5, return_candidates=True)
    return D.pack_records(det)


lo, hi = D.shard_range(G, rank, world)
gathered = D.all_gather_records(run(lo, hi))
torch.cuda.synchronize()
if rank == 0:
    single = run(0, G)
    ok = torch.equal(gathered, single)
    n_det = [int(r[0]) for r in single]
    print("world=%d global_batch=%d detections/image=%s gathered==single-GPU: %s" % (world, G, n_det, ok))
    assert ok
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
