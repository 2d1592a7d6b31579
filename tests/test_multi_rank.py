"""world_size-2 gloo test of the N>1 host logic (sharding + count/payload all-gather + compaction).
Per-rank conversion is done by the oracle here (CPU box, no GPU); on the GPU box the same
plan_shards / all_gather_records code runs over NCCL in bench.py --gpus N."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from mesh2splat_b200 import _abi, synth
    from mesh2splat_b200.shard import all_gather_records, estimate_cost, plan_shards

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tri = synth.displaced_sphere(24, 12, seed=2)
        s = _abi.Scene(tri, [_abi.Primitive(0, len(tri), (1, 1, 1, 1), 0, -1, -1)], [synth.random_texture(32, 32, 1)])
        s.compute_bboxes()
        R, layout = 64, _abi.LAYOUT_PACKED56
        cost = estimate_cost(s.triangles, s.primitives[0].bbox_min, s.primitives[0].bbox_max, R)
        shards = plan_shards(s.triangle_count, world, cost)
        first, count = shards[rank]
        rec, keys, total = oracle.convert(s, R, layout, first_triangle=first, triangle_count=max(count, 0) or 0) \
            if count else (np.zeros(0, _abi.record_dtype(layout)), np.zeros(0, np.uint64), 0)
        local = torch.from_numpy(np.frombuffer(rec.tobytes(), np.uint8).copy())
        buf, counts = all_gather_records(local, len(rec), _abi.STRIDES[layout], dist, torch)
        whole, wkeys, wtotal = oracle.convert(s, R, layout)
        got = buf.numpy().view(_abi.record_dtype(layout))
        # rank-major concatenation of contiguous triangle ranges == the single-process output order
        ok = (sum(counts) == wtotal == len(got)) and got.tobytes() == whole.tobytes()
        q.put((rank, ok, counts, shards))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_gather_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, counts, shards in res:
        assert ok, f"rank {rank}: gathered buffer differs from the single-process result"
        assert len(counts) == 2 and all(c > 0 for c in counts)
        assert shards[0][0] == 0 and shards[0][0] + shards[0][1] == shards[1][0]


def test_plan_shards_properties():
    sys.path.insert(0, ROOT)
    from mesh2splat_b200.shard import plan_shards
    for T in (0, 1, 7, 1000, 70074):
        for w in (1, 2, 3, 8):
            sh = plan_shards(T, w)
            assert len(sh) == w and sh[0][0] == 0 and sum(c for _, c in sh) == T
            assert all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in sh) - min(c for _, c in sh) <= 1
    cost = np.ones(100); cost[:10] = 10.0   # the first 10 triangles carry about half of the work
    sh = plan_shards(100, 2, cost)
    assert 8 <= sh[0][1] <= 12 and sh[0][1] + sh[1][1] == 100
    with pytest.raises(ValueError):
        plan_shards(10, 0)
