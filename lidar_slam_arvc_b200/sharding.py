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
