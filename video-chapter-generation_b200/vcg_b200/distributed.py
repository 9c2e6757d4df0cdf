"""Multi-GPU scoring: candidate clips are independent, so the global clip list is cut into contiguous per-rank
shards (no data-path collective) and the only exchange is one all-gather of the [n,2] boundary logits, after which
every rank holds all scores and derives identical labels / cut points (the reference itself has no
inference-time collective, its NCCL use is training-only: train_video_segment_ddp.py:64-86).
One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_size(n_items, world_size):
    return (n_items + world_size - 1) // world_size


def shard_range(n_items, rank, world_size):
    """Contiguous shard [lo, hi) of rank `rank`; every shard has ceil(n/W) items except the (possibly empty) tail."""
    per = shard_size(n_items, world_size)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def allgather_scores(local_scores, n_total, group=None):
    """local_scores [n_local, C] (this rank's shard, in order) -> [n_total, C] on every rank.

    The send buffer is padded to ceil(n_total/W) rows so that a single fixed-size all-gather serves ragged shards.
    """
    world = dist.get_world_size(group)
    per = shard_size(n_total, world)
    C = local_scores.shape[1]
    send = local_scores
    if local_scores.shape[0] != per:
        send = torch.zeros(per, C, dtype=local_scores.dtype, device=local_scores.device)
        send[:local_scores.shape[0]] = local_scores
    recv = torch.empty(world * per, C, dtype=local_scores.dtype, device=local_scores.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    else:
        chunks = list(recv.view(world, per, C).unbind(0))
        dist.all_gather(chunks, send.contiguous(), group=group)
    return recv[:n_total]


def allgather_scores_async(local_scores, n_total, group=None):
    """Same exchange, not waited for: returns (result [n_total, C], work).  The collective runs on the backend's own
    stream, so the next scoring pass can start underneath it; call ``work.wait()`` before reading the result (it makes
    the current stream wait, it does not block the host on NCCL)."""
    world = dist.get_world_size(group)
    per = shard_size(n_total, world)
    C = local_scores.shape[1]
    send = local_scores
    if local_scores.shape[0] != per:
        send = torch.zeros(per, C, dtype=local_scores.dtype, device=local_scores.device)
        send[:local_scores.shape[0]] = local_scores
    recv = torch.empty(world * per, C, dtype=local_scores.dtype, device=local_scores.device)
    if dist.get_backend(group) == "nccl":
        work = dist.all_gather_into_tensor(recv, send.contiguous(), group=group, async_op=True)
    else:
        chunks = list(recv.view(world, per, C).unbind(0))
        work = dist.all_gather(chunks, send.contiguous(), group=group, async_op=True)
    return recv[:n_total], work


def score_sharded(score_fn, n_clips, group=None):
    """score_fn(lo, hi) -> logits [hi-lo, 2] for this rank's shard; returns all logits [n_clips, 2] on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_clips, rank, world)
    local = score_fn(lo, hi)
    assert local.shape[0] == hi - lo
    return allgather_scores(local, n_clips, group)
