"""Image-sharded data parallelism: one process per GPU, no data-path collective inside the path;
one NCCL all-gather of fixed-size per-image detection records at the end.  The reference
is single-GPU (src/apply_net.py:113-114); this is the only parallel strategy the path needs because
images are independent problems."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # asynchronous NCCL errors abort the communicator and raise (instead of hanging the other ranks)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_range(n_items, rank, world):
    """Contiguous slice of a global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def record_width(max_dets, K):
    return 1 + max_dets * (4 + 1 + 1 + K + 16)


def pack_records(det):
    """det dict of ops.nms_fuse -> (B, 1 + max_dets*(22+K)) fp32: count, then per detection
    box(4) score(1) class(1) probs(K) cov(16); rows past `count` are zero.  One launch of pod_wire_records."""
    from . import ops
    return ops.wire_records(det, xywh=False)


def unpack_records(rec, max_dets, K):
    """Inverse of pack_records -> list of dicts of tensors (one per image)."""
    out = []
    w = 4 + 1 + 1 + K + 16
    for row in rec:
        n = int(row[0].item())
        body = row[1:].reshape(max_dets, w)[:n]
        out.append({"boxes": body[:, 0:4], "scores": body[:, 4], "classes": body[:, 5].to(torch.int64),
                    "probs": body[:, 6:6 + K], "cov": body[:, 6 + K:].reshape(n, 4, 4)})
    return out


def all_gather_records(rec):
    """(B_local, w) -> (world*B_local, w) on every rank; identity when not distributed.
    Equal B_local on all ranks (pad the last shard)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    out = torch.empty((dist.get_world_size() * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    # NCCL reports communicator faults asynchronously (ncclCommGetAsyncError, polled by ProcessGroupNCCL's watchdog,
    # which init_from_env arms): wait() on the work object is where such a fault -- or a timeout of a dead peer --
    # surfaces as an exception, at the one collective of the path instead of at some later unrelated CUDA call
    try:
        work = dist.all_gather_into_tensor(out, rec, async_op=True)
        work.wait()
    except Exception as e:  # noqa: BLE001
        raise RuntimeError("all-gather of the detection records failed on rank %d: %s" % (dist.get_rank(), e)) from e
    return out
