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
