"""N>1 host logic on CPU: world_size-2 gloo processes shard a pair stream, build records for their
block and gather them; every rank must see the complete, ordered table."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT  # noqa: F401  (sys.path)
from mfr_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = parallel.shard_range(n_pairs, rank, world)
    ids = list(range(a, b))
    g = torch.Generator().manual_seed(100)
    R_all = torch.randn(n_pairs, 3, 3, generator=g, dtype=torch.float64)
    t_all = torch.randn(n_pairs, 3, generator=g, dtype=torch.float64)
    R_all[3] = float("nan")                     # a failed pair travels as NaN
    rec = parallel.pack_records(ids, R_all[a:b], t_all[a:b], torch.arange(a, b))
    table = parallel.gather_records(rec, n_pairs)
    ok = (table.shape == (n_pairs, parallel.RECORD_WIDTH)
          and torch.equal(table[:, 0], torch.arange(n_pairs, dtype=torch.float64))
          and torch.equal(table[:, 1], torch.arange(n_pairs, dtype=torch.float64))
          and torch.allclose(table[:, 11:14], t_all)
          and torch.allclose(table[:, 2:11].reshape(-1, 3, 3), R_all, equal_nan=True))
    med = parallel.median_pose_errors(table, R_all.nan_to_num(0), t_all)
    q.put((rank, bool(ok), med[2]))
    dist.destroy_process_group()


def test_shard_range_partitions_stream():
    for n in (0, 1, 7, 16, 10000):
        for w in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_pairs = 11                                # ragged: 6 + 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, n_pairs - 1), (1, True, n_pairs - 1)]


def test_median_pose_errors_formula():
    from mfr_b200 import synth
    Rs = np.stack([synth.rodrigues([0.01 * k, 0, 0]) for k in range(1, 6)])
    ts = np.zeros((5, 3))
    table = parallel.pack_records(list(range(5)), torch.tensor(Rs), torch.tensor(ts) + 0.002, torch.ones(5))
    ang, dt, n = parallel.median_pose_errors(table, np.stack([np.eye(3)] * 5), ts)
    assert n == 5 and abs(ang - 0.03) < 1e-9 and abs(dt - 0.002 * 3 ** 0.5) < 1e-12


class _FakePipe:
    """One-deep pipeline stand-in: result of a batch = a deterministic function of its pair ids."""

    def __init__(self):
        self.prev = None
        self.calls = []

    @staticmethod
    def _result(ids):
        i = torch.tensor(ids, dtype=torch.float64)
        R = torch.eye(3, dtype=torch.float64)[None].repeat(len(ids), 1, 1) * (i[:, None, None] + 1)
        return R, torch.stack([i, 2 * i, 3 * i], 1), torch.tensor(ids, dtype=torch.int32) * 10

    def submit(self, ids):
        assert len(ids) == 4                       # always full engine batches
        self.calls.append(list(ids))
        out, self.prev = self.prev, self._result(ids)
        return out

    def drain(self):
        return self.prev


def _stream_worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipe = _FakePipe()
    table = parallel.run_stream(n_pairs, 4, pipe.submit, pipe.drain, rank, world)
    i = torch.arange(n_pairs, dtype=torch.float64)
    ok = (table.shape == (n_pairs, parallel.RECORD_WIDTH) and torch.equal(table[:, 0], i) and torch.equal(table[:, 1], 10 * i)
          and torch.equal(table[:, 2], i + 1) and torch.equal(table[:, 11:14], torch.stack([i, 2 * i, 3 * i], 1)))
    a, b = parallel.shard_range(n_pairs, rank, world)
    padded_ok = all(c[0] >= a and max(c) < b for c in pipe.calls) and len(pipe.calls) == -(-(b - a) // 4)
    q.put((rank, bool(ok), bool(padded_ok)))
    dist.destroy_process_group()


def test_run_stream_world2_gloo():
    """Contiguous shards, ragged last batch padded inside the shard, one-call-late results, ordered gathered table."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, 2, port, 23, q)) for r in range(2)]     # 12 + 11 pairs, batches of 4
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]


def test_run_stream_single_rank_and_empty():
    pipe = _FakePipe()
    table = parallel.run_stream(6, 4, pipe.submit, pipe.drain)
    assert table.shape == (6, parallel.RECORD_WIDTH) and pipe.calls == [[0, 1, 2, 3], [4, 5, 5, 5]]
    assert parallel.run_stream(0, 4, pipe.submit, pipe.drain).shape == (0, parallel.RECORD_WIDTH)
