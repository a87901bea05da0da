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
from pod_compare_b200 import wire


def _pack_cpu(det):
    """Layout of pod_wire_records(xywh=0) restated with torch ops (the product packs on the GPU; the CPU suite only
    needs records to push through the sharding / gather / unpack plumbing)."""
    B, D_ = det["scores"].shape
    body = torch.cat([det["boxes"], det["scores"].unsqueeze(-1), det["classes"].to(torch.float32).unsqueeze(-1),
                      det["probs"], det["cov"].reshape(B, D_, 16)], dim=2)
    valid = (torch.arange(D_)[None, :] < det["count"][:, None]).unsqueeze(-1)
    body = torch.where(valid, body, torch.zeros((), dtype=body.dtype))
    return torch.cat([det["count"].to(torch.float32).unsqueeze(-1), body.reshape(B, -1)], dim=1).contiguous()


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
    rec = _pack_cpu(det)
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
    rec = _pack_cpu(det)
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
    want = torch.cat([_pack_cpu(_fake_det(2, 100, 7, 100 + r)) for r in range(2)], 0).numpy()
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)   # rank order == image order


def test_records_to_json_schema_and_filtering():
    """Host half of the wire format: records already in the JSON layout (what pod_wire_records(xywh=1) writes) ->
    the reference's result dicts (inference_utils.py:486-500): schema, image order, category -1 dropped."""
    K, D_ = 7, 5
    w = 22 + K
    rec = np.zeros((3, 1 + D_ * w), np.float32)
    rng = np.random.RandomState(0)
    rec[0, 0] = 2
    rec[0, 1:1 + 2 * w] = rng.rand(2 * w).astype(np.float32)
    rec[0, 1 + 5] = 3.0               # category ids
    rec[0, 1 + w + 5] = -1.0          # no id in the test dataset -> dropped
    rec[2, 0] = 1
    rec[2, 1:1 + w] = rng.rand(w).astype(np.float32)
    rec[2, 1 + 5] = 1.0
    out = wire.records_to_json(rec, ["a", "b", "c"], K, D_)
    assert [r["image_id"] for r in out] == ["a", "c"]
    r = out[0]
    assert set(r) == {"image_id", "category_id", "bbox", "score", "cls_prob", "bbox_covar"}
    assert r["category_id"] == 3 and isinstance(r["category_id"], int)
    assert r["bbox"] == rec[0, 1:5].tolist() and r["score"] == float(rec[0, 5])
    assert r["cls_prob"] == rec[0, 7:7 + K].tolist() and np.array_equal(np.array(r["bbox_covar"], np.float32).ravel(), rec[0, 7 + K:1 + w])
    import json
    json.dumps(out)                     # plain Python numbers only


def test_category_mapping_follows_apply_net():
    """src/apply_net.py:53-79 with the reference's own tables (src/core/datasets/metadata.py)."""
    bdd, kitti = wire.BDD_THING_DATASET_ID_TO_CONTIGUOUS_ID, wire.KITTI_THING_DATASET_ID_TO_CONTIGUOUS_ID
    same = wire.build_category_mapping("bdd_train", "bdd_val", bdd, bdd)
    assert same == {i: i + 1 for i in range(7)}
    cross = wire.build_category_mapping("bdd_train", "kitti_val", bdd, kitti)
    assert cross == {0: 1, 3: 2}        # BDD car -> KITTI id 1, BDD person -> KITTI id 2; every other class is dropped
    m = wire.category_map_tensor(cross, 7, "cpu")
    assert m.tolist() == [1, -1, -1, 2, -1, -1, -1]
    import pytest
    with pytest.raises(ValueError):
        wire.build_category_mapping("lyft_train", "kitti_val", bdd, kitti)
