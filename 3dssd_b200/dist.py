"""Batch sharding over the GPUs of one node.

Every operator of the SA path is independent per scene and inference BatchNorm uses moving statistics, so
the path shards by scene with no activation exchange: rank r of G processes scenes [r*B/G, (r+1)*B/G).  The
only collective is ONE all-gather per step of the fixed-size per-scene detection block ([100, 9] fp32 + a
count word); at <= 64 scenes x 3.6 KB it is latency-bound, so a single NCCL call is the right tool.
The reference has no inference sharding (evaluation is single-GPU bs=1, lib/core/evaluator.py:145-147).

DetectionGather keeps the whole exchange inside the captured step: the NMS kernel writes its block and count
straight into a preallocated send buffer, ncclAllGather on that buffer is captured into the step's CUDA graph, and the
gathered result is read through views of the receive buffer -- no pack / unpack kernels and no host work per step
beyond graph.replay().
"""
import os

import torch
import torch.distributed as dist

MAX_OUTPUT_NUM = 100
BLOCK_COLS = 9


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


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class DetectionGather:
    """Send / receive buffers of the per-step all-gather, owned by one step pipeline.

    total_scenes scenes are sharded with shard_bounds; every rank sends b_max = ceil(total / world) slots (unused
    slots stay zero).  Layout of a rank's slice (bytes): [b_max * 100 * 9 fp32 block | b_max int32 count].
      out()      -> (block [b_local,100,9] fp32, count [b_local] int32) views of the SEND buffer: hand them to the NMS
      gather()   -> enqueue the all-gather on the current stream (CUDA-graph capturable; a no-op for one rank)
      result()   -> (blocks [total,100,9], counts [total]) of all scenes, assembled from views of the RECEIVE buffer
      raw        -> the receive buffer itself (uint8), what a host reads back in one copy
    group: the process group (= NCCL communicator) to use.  Step pipelines that run concurrently on different streams
    must not share one -- collectives of one communicator may not overlap -- so each pipeline passes its own group."""

    def __init__(self, total_scenes, device, group=None, max_output=MAX_OUTPUT_NUM):
        self.world, self.rank = _world(group)
        self.group = group
        self.total = int(total_scenes)
        self.max_output = int(max_output)
        lo, hi = shard_bounds(self.total, self.rank, self.world)
        self.b_local = hi - lo
        self.b_max = (self.total + self.world - 1) // self.world
        self.block_bytes = self.b_max * self.max_output * BLOCK_COLS * 4
        self.slice_bytes = (self.block_bytes + self.b_max * 4 + 15) // 16 * 16      # 16-byte multiple (vector copies)
        self._alloc(device)

    def _alloc(self, device):
        self.send = torch.zeros((self.slice_bytes,), dtype=torch.uint8, device=device)
        self.raw = self.send if self.world == 1 else torch.zeros((self.world * self.slice_bytes,), dtype=torch.uint8,
                                                                 device=device)

    def _views(self, buf):
        blk = buf[: self.block_bytes].view(torch.float32).view(self.b_max, self.max_output, BLOCK_COLS)
        cnt = buf[self.block_bytes: self.block_bytes + self.b_max * 4].view(torch.int32)
        return blk, cnt

    def out(self):
        blk, cnt = self._views(self.send)
        return blk[: self.b_local], cnt[: self.b_local]

    def gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.raw, self.send, group=self.group)

    def result(self):
        blocks, counts = [], []
        for r in range(self.world):
            lo, hi = shard_bounds(self.total, r, self.world)
            blk, cnt = self._views(self.raw[r * self.slice_bytes: (r + 1) * self.slice_bytes])
            blocks.append(blk[: hi - lo]); counts.append(cnt[: hi - lo])
        if self.world == 1:
            return blocks[0], counts[0]
        return torch.cat(blocks, dim=0), torch.cat(counts, dim=0)


def peer_layout(world, slice_bytes):
    """Byte offsets inside one pipeline's share of the symmetric allocation (csrc/peer_gather.cu):
    ((recv_off parity 0, parity 1), (flag_off parity 0, parity 1), total bytes)."""
    recv = world * slice_bytes
    flags = (world * 4 + 15) // 16 * 16
    return (0, recv), (2 * recv, 2 * recv + flags), 2 * recv + 2 * flags


class PeerArena:
    """ONE symmetric allocation (mapped into every rank of `group`) shared by all step pipelines of a process: a single
    rendezvous instead of one per pipeline.  take(nbytes) hands out 256-byte aligned shares, in the same order on every rank."""

    def __init__(self, nbytes, device, group=None):
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.buf = symm.empty((int(nbytes),), dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()                                  # flags start at 0, before any peer can write
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        self.peer_bases = [int(p) for p in self.handle.buffer_ptrs]
        self.used = 0

    @staticmethod
    def share_bytes(world, slice_bytes):
        return (peer_layout(world, slice_bytes)[2] + 255) // 256 * 256

    def take(self, nbytes):
        off = self.used
        self.used += (int(nbytes) + 255) // 256 * 256
        if self.used > self.buf.numel():
            raise ValueError("PeerArena exhausted")
        return off


class PeerGather(DetectionGather):
    """DetectionGather with the exchange done by ONE kernel of this library over NVLink peer memory
    (tf_ops.peer_allgather, csrc/peer_gather.cu) instead of ncclAllGather: every rank stores its slice straight into
    the peers' symmetric buffers, publishes a flag and waits for theirs.  Same interface and the same `raw` layout;
    capturable; pipelines need no process group of their own (nothing here is a communicator).  Every rank must call
    gather() the same number of times."""

    def __init__(self, total_scenes, device, arena, max_output=MAX_OUTPUT_NUM):
        self.arena = arena
        super().__init__(total_scenes, device, group=arena.group, max_output=max_output)

    def _alloc(self, device):
        from . import tf_ops                                   # (tf_ops does not import this module)
        self._tf_ops = tf_ops
        super()._alloc(device)
        self.raw = torch.zeros((self.world * self.slice_bytes,), dtype=torch.uint8, device=device)
        self.recv_off, self.flag_off, total = peer_layout(self.world, self.slice_bytes)
        base = self.arena.take(total)
        self.peers = torch.tensor([b + base for b in self.arena.peer_bases], dtype=torch.int64, device=device)
        self.state = torch.zeros((4,), dtype=torch.int32, device=device)   # replay count, CTA count, time-outs

    def timeouts(self):
        """Waits that gave up (a peer that never delivered): 0 in a healthy run.  Synchronises."""
        return int(self.state[2].item())

    def gather(self):
        self._tf_ops.peer_allgather(self.send, self.peers, self.world, self.rank, self.slice_bytes, self.recv_off, self.flag_off,
                                    self.state, self.raw)


def gather_detections(block, count, total_scenes, group=None):
    """One-shot form: all-gather per-scene detection blocks block [b_local, 100, 9] fp32 / count [b_local] int32 of a
    batch of `total_scenes` scenes sharded with shard_bounds -> ([total, 100, 9], [total]) on every rank.
    total_scenes is required: with uneven shards the ranks cannot derive a common slot count from their own b_local."""
    world, rank = _world(group)
    if world == 1:
        return block, count
    lo, hi = shard_bounds(int(total_scenes), rank, world)
    if block.shape[0] != hi - lo or count.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d scenes but shard_bounds(%d, %d, %d) assigns it %d"
                         % (rank, block.shape[0], total_scenes, rank, world, hi - lo))
    g = DetectionGather(total_scenes, block.device, group=group, max_output=block.shape[1])
    b, c = g.out()
    b.copy_(block); c.copy_(count)
    g.gather()
    return g.result()
