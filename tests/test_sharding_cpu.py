"""Multi-GPU host logic on CPU: contiguous pair sharding and the all-gather of result records over `gloo`, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from lidar_slam_arvc_b200 import sharding
from lidar_slam_arvc_b200.engine import RESULT_DTYPE


def test_shard_bounds_partition():
    for n in (0, 1, 7, 99, 10000):
        for w in (1, 2, 3, 4, 8):
            b = [sharding.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_scans_of_pairs_and_cache_order():
    tg, sr = [5, 0, 5, 2], [9, 1, 6, 3]
    np.testing.assert_array_equal(sharding.scans_of_pairs(tg, sr), [0, 1, 2, 3, 5, 6, 9])
    order = sharding.sort_pairs_for_cache(tg, sr)
    assert [(tg[k], sr[k]) for k in order] == [(0, 1), (2, 3), (5, 6), (5, 9)]


def _make_records(rank, n):
    rec = np.zeros(n, dtype=RESULT_DTYPE)
    for k in range(n):
        rec[k]["pair"] = 1000 * rank + k
        rec[k]["updates"] = k + 1
        rec[k]["T"] = np.eye(4) * (rank + 1) + k
        rec[k]["fitness"] = 0.5 + 0.01 * k
        rec[k]["rmse"] = 0.1 * (rank + 1)
    return rec


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = sharding.gather_records(_make_records(rank, counts[rank]))
        known = sharding.gather_records(_make_records(rank, counts[rank]), counts=list(counts))     # counts known: one collective
        assert known.tobytes() == out.tobytes()
        try:
            sharding.gather_records(_make_records(rank, counts[rank]), counts=[c + 1 for c in counts])
            raise AssertionError("wrong counts accepted")
        except ValueError:
            pass
        q.put((rank, out.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(3, 3), (4, 1), (0, 2)])
def test_gather_records_gloo_world2(counts):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.concatenate([_make_records(r, counts[r]) for r in range(2)])
    for r in range(2):
        out = np.frombuffer(got[r], dtype=RESULT_DTYPE)
        assert len(out) == sum(counts)
        assert out.tobytes() == expect.tobytes()          # bit-exact, rank order


def test_gather_without_process_group_is_identity():
    rec = _make_records(0, 3)
    assert sharding.gather_records(rec).tobytes() == rec.tobytes()


def test_balanced_bounds_equalise_cost_and_keep_order():
    rng = np.random.default_rng(0)
    cost = rng.gamma(2.0, 5.0, size=10000)
    cost[2000:3000] *= 3.0                                        # an expensive stretch of the sorted list
    for w in (1, 2, 4, 8):
        b = sharding.balanced_bounds(cost, w)
        assert b[0] == 0 and b[-1] == len(cost) and all(x <= y for x, y in zip(b[:-1], b[1:]))
        loads = [cost[b[r]:b[r + 1]].sum() for r in range(w)]
        assert max(loads) / np.mean(loads) < 1.01
        even = [cost[lo:hi].sum() for lo, hi in (sharding.shard_bounds(len(cost), w, r) for r in range(w))]
        assert max(loads) <= max(even) + 1e-9
    assert sharding.balanced_bounds([], 3) == [0, 0, 0, 0]
    assert sharding.balanced_bounds([1.0], 4)[-1] == 1


def test_batches_and_reassembly_of_the_global_list():
    bounds, B = [0, 5, 5, 12], 3                                   # three ranks: 5, 0 and 7 pairs; batches of <= 3
    assert [sharding.batch_counts(bounds, b, B) for b in range(3)] == [[3, 0, 3], [2, 0, 3], [0, 0, 1]]
    glob = np.zeros(12, dtype=RESULT_DTYPE)
    glob["pair"] = np.arange(12)
    glob["rmse"] = np.arange(12) * 0.5
    parts = []
    for b in range(3):                                             # what the all-gather of batch b delivers: rank order
        parts.append(np.concatenate([glob[bounds[r] + b * B: bounds[r] + b * B + c] for r, c in enumerate(sharding.batch_counts(bounds, b, B))]))
    out = sharding.assemble_global(parts, bounds, B)
    assert out.tobytes() == glob.tobytes()
    with pytest.raises(ValueError):
        sharding.assemble_global([parts[0][:-1]] + parts[1:], bounds, B)


def _balance_worker(rank, world, port, q):
    """Two ranks, a list whose second half is three times as expensive per pass: count-based bounds first, then the bounds the
    balancer derives from the gathered passes and every rank's own measured time (exchanged over the process group)."""
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1000
        passes = np.full(n, 8)
        true_cost = (3.0 + passes) * np.where(np.arange(n) < n // 2, 1.0, 3.0)
        bounds = [sharding.shard_bounds(n, world, r)[0] for r in range(world)] + [n]
        for _ in range(4):
            lo, hi = bounds[rank], bounds[rank + 1]
            rec = np.zeros(hi - lo, dtype=RESULT_DTYPE)
            rec["passes"] = passes[lo:hi]
            every = sharding.gather_records(rec, counts=[bounds[r + 1] - bounds[r] for r in range(world)])
            mine = torch.tensor([true_cost[lo:hi].sum()], dtype=torch.float64)          # this rank's "measured busy time"
            busy = torch.zeros(world, dtype=torch.float64)
            dist.all_gather_into_tensor(busy, mine)
            bounds = sharding.rebalanced_bounds(every["passes"], bounds, busy.numpy())
        q.put((rank, bounds, [float(true_cost[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]))
    finally:
        dist.destroy_process_group()


def test_run_time_balancer_converges_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_balance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((r, (b, loads)) for r, b, loads in (q.get(timeout=120) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == got[1][0]                                  # every rank derives the same bounds
    loads = got[0][1]
    assert max(loads) / np.mean(loads) < 1.02, loads               # 2.0 before (count split): 1 : 3
    assert 660 <= got[0][0][1] <= 672                              # the cheap half grew: 500 + 500 / 3
