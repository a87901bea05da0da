"""Host-side logic of the multi-GPU path on CPU: image sharding and the gather of fixed-size detection
records over a world_size-2 gloo group, plus record packing and the JSON wire format."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pod_compare_b200 import distributed as D
from pod_compare_b200.inference_utils import covar_xyxy_to_xywh, instances_to_json
from pod_compare_b200.structures import Boxes, Instances


def _fake_det(B, D_, K, seed):
    g = torch.Generator().manual_seed(seed)
    count = torch.randint(0, D_ + 1, (B,), generator=g, dtype=torch.int32)
    return {"boxes": torch.rand((B, D_, 4), generator=g), "scores": torch.rand((B, D_), generator=g),
            "classes": torch.randint(0, K, (B, D_), generator=g, dtype=torch.int32),
            "probs": torch.rand((B, D_, K), generator=g), "cov": torch.rand((B, D_, 4, 4), generator=g), "count": count}


def test_shard_range_partitions_the_batch():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    det = _fake_det(3, 100, 7, 0)
    rec = D.pack_records(det)
    assert rec.shape == (3, D.record_width(100, 7))
    out = D.unpack_records(rec, 100, 7)
    for b in range(3):
        n = int(det["count"][b])
        assert out[b]["boxes"].shape[0] == n
        assert torch.equal(out[b]["boxes"], det["boxes"][b, :n])
        assert torch.equal(out[b]["classes"], det["classes"][b, :n].long())
        assert torch.equal(out[b]["cov"], det["cov"][b, :n])
        assert float(rec[b, 1 + n * 29:].abs().sum()) == 0.0


def _worker(rank, world, port, q):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    D.init_from_env("gloo")
    det = _fake_det(2, 100, 7, 100 + rank)
    rec = D.pack_records(det)
    allrec = D.all_gather_records(rec)
    q.put((rank, allrec.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.cat([D.pack_records(_fake_det(2, 100, 7, 100 + r)) for r in range(2)], 0).numpy()
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)   # rank order == image order


def test_instances_to_json_schema_and_covariance_transform():
    inst = Instances((720, 1280))
    inst.pred_boxes = Boxes(torch.tensor([[10.0, 20.0, 110.0, 220.0], [0.0, 0.0, 5.0, 5.0]]))
    inst.scores = torch.tensor([0.9, 0.2])
    inst.pred_classes = torch.tensor([2, 5])
    inst.pred_cls_probs = torch.rand((2, 7))
    A = torch.rand((2, 4, 4))
    inst.pred_boxes_covariance = A @ A.transpose(1, 2)
    out = instances_to_json(inst, 42, {2: 3})          # class 5 has no dataset id -> dropped
    assert len(out) == 1
    r = out[0]
    assert set(r) == {"image_id", "category_id", "bbox", "score", "cls_prob", "bbox_covar"}
    assert r["image_id"] == 42 and r["category_id"] == 3 and r["bbox"] == [10.0, 20.0, 100.0, 200.0]
    T = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [-1.0, 0, 1.0, 0], [0, -1.0, 0, 1.0]])
    want = T @ inst.pred_boxes_covariance[0] @ T.T
    assert torch.allclose(torch.tensor(r["bbox_covar"]), want, atol=1e-6)
    assert torch.allclose(covar_xyxy_to_xywh(inst.pred_boxes_covariance)[0], want, atol=1e-6)
