"""Batch sharding over the GPUs of one node.

Every operator of the SA path is independent per scene and inference BatchNorm uses moving statistics, so
the path shards by scene with no activation exchange: rank r of G processes scenes [r*B/G, (r+1)*B/G).  The
only collective is ONE all-gather per step of the fixed-size per-scene detection block ([100, 9] fp32 + a
count word); at <= 64 scenes x 3.6 KB it is latency-bound, so a single NCCL call is the right tool.
The reference has no inference sharding (evaluation is single-GPU bs=1, lib/core/evaluator.py:145-147).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(total, rank, world):
    """Contiguous scene range of `rank`; the first total % world ranks take one extra scene."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    lo, hi = shard_bounds(batch.shape[0], rank, world)
    return batch[lo:hi]


def gather_detections(block, count, total_scenes=None):
    """All-gather the per-scene detection blocks: block [b_local, 100, 9] fp32, count [b_local] int32 ->
    ([B, 100, 9], [B]) on every rank.  One collective: the count travels as a 10th column."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return block, count
    world = dist.get_world_size()
    b_local = block.shape[0]
    if total_scenes is None:
        total_scenes = b_local * world
    b_max = (total_scenes + world - 1) // world
    packed = torch.zeros((b_max, block.shape[1], block.shape[2] + 1), dtype=torch.float32, device=block.device)
    packed[:b_local, :, : block.shape[2]] = block
    packed[:b_local, 0, block.shape[2]] = count.to(torch.float32)
    out = torch.empty((world * b_max,) + tuple(packed.shape[1:]), dtype=torch.float32, device=block.device)
    dist.all_gather_into_tensor(out, packed)
    pieces, counts = [], []
    for r in range(world):
        lo, hi = shard_bounds(total_scenes, r, world)
        seg = out[r * b_max: r * b_max + (hi - lo)]
        pieces.append(seg[:, :, : block.shape[2]])
        counts.append(seg[:, 0, block.shape[2]].to(torch.int32))
    return torch.cat(pieces, dim=0), torch.cat(counts, dim=0)
