"""Multi-GPU fused gather (m2s_convert_gather_enqueue over torch symmetric memory): needs >= 2 GPUs.
Skipped on single-GPU boxes.  Each rank runs in its own process (spawn), NCCL only bootstraps the
symmetric-memory rendezvous; the records travel by peer stores from the fragment kernel."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from mesh2splat_b200 import _abi, synth
    from mesh2splat_b200.api import Context
    from mesh2splat_b200.shard import PeerGather, estimate_cost, plan_shards

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok, msg = True, ""
    try:
        tri = synth.displaced_sphere(64, 32, seed=3)
        s = _abi.Scene(tri, [_abi.Primitive(0, len(tri), (1, 0.9, 0.8, 1), 0, 1, 2)], synth.make_material_textures(128, 4))
        s.compute_bboxes()
        ctx = Context(rank)
        ds = ctx.upload(s)
        for layout in (_abi.LAYOUT_PACKED56, _abi.LAYOUT_REF96):
            stride = _abi.STRIDES[layout]
            R = 160
            cap = 6 * R * R
            pg = PeerGather(ctx, cap, stride, dist, torch)
            cost = estimate_cost(s.triangles, s.primitives[0].bbox_min, s.primitives[0].bbox_max, R)
            first, count = plan_shards(s.triangle_count, world, cost)[rank]
            p = _abi.make_params(R, layout, 0.65, 0, _abi.FLAG_UNCAPPED, first, count)
            for _ in range(3):  # repeated calls: epochs must pair up
                pg.convert_enqueue(ds, p)
            got, n = pg.records(layout)
            from mesh2splat_b200._lib import lib
            if lib().m2s_ctx_status(ctx.handle) != 0:
                ok, msg = False, "m2s_ctx_status: a gather wait timed out"
            whole = ctx.convert(ds, R, layout, flags=_abi.FLAG_UNCAPPED, capacity=cap)
            want = whole.numpy()
            a = np.sort(np.frombuffer(got.tobytes(), np.dtype((np.void, stride))))
            b = np.sort(np.frombuffer(want.tobytes(), np.dtype((np.void, stride))))
            if n != whole.total or len(a) != len(b) or not np.array_equal(a, b):
                ok, msg = False, f"layout {layout}: gathered {n} records vs single-GPU {whole.total}"
            dist.barrier()
        ds.free()
        ctx.close()
    except Exception as e:  # noqa: BLE001
        ok, msg = False, repr(e)
    finally:
        q.put((rank, ok, msg))
        dist.destroy_process_group()


def _run_world(world: int):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, f"rank {rank}: {msg}"


def test_fused_gather_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_world(2)


def test_fused_gather_all_gpus():
    """The same check on every GPU of the box (4 or 8 ranks): the gathered buffer is the single-GPU multiset on every rank."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    _run_world(n)
