"""Multi-GPU plumbing: keyframe pairs are independent ICP problems, so they are sharded across ranks with
no data-path collective; one all-gather of the fixed-size 160-byte result records brings the relative
transforms, fitness and RMSE back.  torch.distributed is plumbing only: `nccl` on the
GPUs (NVLink 5 / NVSwitch), `gloo` in the CPU tests.

Reference analogue: none — run_scanmatcher.py:196-213 and loopclosing.py:80-99 loop over pairs one at a time.
"""
import numpy as np

from .engine import RESULT_DTYPE


def shard_bounds(n_items, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`: consecutive pairs share scans, so contiguous blocks let a
    rank preprocess each of its scans once (one-scan halo at block edges)."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scans_of_pairs(tgt_ids, src_ids):
    """Sorted unique scan ids a list of pairs touches (what one rank has to upload and preprocess)."""
    return np.unique(np.concatenate([np.asarray(tgt_ids, dtype=np.int64), np.asarray(src_ids, dtype=np.int64)]))


def sort_pairs_for_cache(tgt_ids, src_ids):
    """Loop-closure pairs: order by (target, source) so that a rank's scan cache is reused.  Returns the permutation."""
    t = np.asarray(tgt_ids, dtype=np.int64)
    s = np.asarray(src_ids, dtype=np.int64)
    return np.lexsort((s, t))


def gather_records(records, device=None, group=None, counts=None):
    """All-gather per-rank result records (numpy RESULT_DTYPE, possibly different counts per rank); every rank
    returns the concatenation in rank order.  Without an initialised process group: identity.
    `counts` (records per rank, e.g. from shard_bounds) saves the exchange of the counts: one collective per batch."""
    import torch
    import torch.distributed as dist
    records = np.ascontiguousarray(records, dtype=RESULT_DTYPE)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    if counts is None:
        n = torch.tensor([len(records)], dtype=torch.int64, device=dev)
        all_n = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_n, n, group=group)
        counts = [int(c) for c in all_n.cpu().tolist()]
    elif len(counts) != world or int(counts[dist.get_rank(group)]) != len(records):
        raise ValueError("gather_records: counts do not match this rank's records")
    counts = [int(c) for c in counts]
    cap = max(max(counts), 1)
    words = RESULT_DTYPE.itemsize // 8
    buf = torch.zeros((cap, words), dtype=torch.float64, device=dev)
    if len(records):
        buf[:len(records)] = torch.from_numpy(records.view(np.float64).reshape(len(records), words)).to(dev)
    out = torch.empty((world * cap, words), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    host = out.cpu().numpy().reshape(world, cap, words)
    parts = [host[r, :counts[r]].reshape(-1).view(RESULT_DTYPE) for r in range(world)]
    return np.concatenate(parts) if parts else records


class _DeviceBytes:
    """A raw device allocation owned by the engine, exposed through __cuda_array_interface__ so that torch can wrap it
    without a copy (plumbing only: torch.distributed needs a tensor to hand to NCCL)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


class DeviceGather:
    """All-gather of the 160-byte result records straight from the engine's device memory.

    `start(ticket)` is called right after `icp_batch_async`: on the engine's own stream - i.e. ordered after the last
    kernel of the batch, with no host synchronisation - the rank's records are copied into the fixed-size send slot,
    all-gathered over NCCL and, on `dst_rank`, copied to a pinned host buffer.  `finish()` waits for that one copy.
    Because nothing blocks the host in between, the caller can enqueue the next batch before it collects this one:
    the gather of batch k overlaps the kernels of batch k + 1."""

    def __init__(self, engine, device, capacity, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.eng, self.group = torch, dist, engine, group
        self.device = torch.device(device)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap = max(int(capacity), 1)              # most records one rank contributes to one gather
        self.words = RESULT_DTYPE.itemsize // 8
        self.stream = torch.cuda.ExternalStream(engine.stream_handle(), device=self.device)
        # two buffer sets, used alternately: gather k + 1 may be in flight while the host still reads gather k
        self.slots = [{"send": torch.zeros((self.cap, self.words), dtype=torch.float64, device=self.device),
                       "recv": torch.empty((self.world * self.cap, self.words), dtype=torch.float64, device=self.device),
                       "host": torch.empty((self.world * self.cap, self.words), dtype=torch.float64).pin_memory(),
                       "event": torch.cuda.Event()} for _ in range(2)]
        self.started = 0

    def start(self, ticket, counts):
        """`counts`: records every rank contributes (known to all ranks from the sharding).  Returns a handle for finish().
        At most two gathers may be outstanding."""
        torch = self.torch
        counts = [int(c) for c in counts]
        ptr, n = self.eng.icp_batch_device_records(ticket)
        if len(counts) != self.world or n != counts[self.rank] or max(counts) > self.cap:
            raise ValueError("DeviceGather: this rank's batch has %d pairs, counts say %s (capacity %d)" % (n, counts, self.cap))
        slot = self.slots[self.started % 2]
        slot["counts"] = counts
        self.started += 1
        with torch.cuda.stream(self.stream):
            if n:
                src = torch.as_tensor(_DeviceBytes(ptr, n * self.words), device=self.device).view(n, self.words)
                slot["send"][:n].copy_(src, non_blocking=True)
            self.dist.all_gather_into_tensor(slot["recv"], slot["send"], group=self.group)
            slot["host"].copy_(slot["recv"], non_blocking=True)
            slot["event"].record(self.stream)
        return slot

    def finish(self, slot):
        """Concatenation of all ranks' records in rank order (numpy RESULT_DTYPE)."""
        slot["event"].synchronize()
        host = slot["host"].numpy().reshape(self.world, self.cap, self.words)
        return np.concatenate([host[r, :slot["counts"][r]].reshape(-1).view(RESULT_DTYPE).copy() for r in range(self.world)])

    def close(self):
        """Drop the device buffers while the engine's stream still exists (torch records an event on every stream a block
        was used on when it frees the block)."""
        self.torch.cuda.synchronize(self.device)
        self.slots = []


def balanced_bounds(costs, world_size):
    """Contiguous shard bounds [world_size + 1] with (nearly) equal summed cost: scan locality of the sorted pair list is
    kept, the shards differ in length instead of in work."""
    c = np.cumsum(np.asarray(costs, dtype=np.float64))
    n = len(c)
    if n == 0:
        return [0] * (world_size + 1)
    inner = [int(np.searchsorted(c, c[-1] * r / world_size)) for r in range(1, world_size)]
    b = [0] + inner + [n]
    for k in range(1, len(b)):                       # monotone, and no shard past the end
        b[k] = min(max(b[k], b[k - 1]), n)
    return b


def batch_counts(bounds, b, batch):
    """Records every rank contributes to device batch `b` (batches of <= `batch` pairs) of its shard [bounds[r], bounds[r+1])."""
    return [min(batch, max(0, (bounds[r + 1] - bounds[r]) - b * batch)) for r in range(len(bounds) - 1)]


def assemble_global(parts, bounds, batch):
    """The gathered batches (parts[b] = all ranks' records of batch b, in rank order) back in the order of the global list."""
    world, n = len(bounds) - 1, bounds[-1]
    out = np.zeros(n, dtype=RESULT_DTYPE)
    for b, part in enumerate(parts):
        off = 0
        for r, c in enumerate(batch_counts(bounds, b, batch)):
            lo = bounds[r] + b * batch
            out[lo:lo + c] = part[off:off + c]
            off += c
        if off != len(part):
            raise ValueError("assemble_global: batch %d holds %d records, the bounds say %d" % (b, len(part), off))
    return out


def rebalanced_bounds(passes, bounds, busy_ms, base_cost=3.0):
    """Run-time load balancer of the sharded pair list.  `passes[k]`: correspondence passes pair k needed in the batch that
    has just been gathered (every rank holds all records); `busy_ms[r]`: time rank r's GPU spent on its shard.  A pair's
    cost estimate is (base_cost + passes) x the time its rank needed per such unit; the returned contiguous bounds give
    every rank the same summed estimate for the NEXT batch (scan locality of the sorted list is kept)."""
    cost = base_cost + np.asarray(passes, dtype=np.float64)
    for r in range(len(bounds) - 1):
        seg = cost[bounds[r]:bounds[r + 1]]
        if len(seg) and seg.sum() > 0 and busy_ms[r] > 0:
            cost[bounds[r]:bounds[r + 1]] = seg * (float(busy_ms[r]) / seg.sum())
    return balanced_bounds(cost, len(bounds) - 1)
