"""N > 1 host logic on CPU: world_size 2, gloo backend (127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    r, l, w = vd.init(backend="gloo")
    assert (r, w) == (rank, world)
    d = 7
    blocks = vd.frame_blocks(d, world)
    b, e = blocks[rank]
    local = np.arange(d * 6, dtype=np.float32).reshape(d, 3, 2)[b:e] * 2.0          # this rank's frames of a (d,3,2) result
    whole = vd.gather_frames(local, d)
    units, secs = vd.reduce_throughput(10.0 * (rank + 1), 0.5 + rank)                # SUM of units, MAX of seconds
    dist.barrier()
    q.put((rank, whole, units, secs))
    dist.destroy_process_group()


def test_frame_blocks_and_chain_plan():
    from videomorphing_b200 import dist as vd
    for d in (1, 2, 7, 120, 121):
        for world in (1, 2, 4, 8):
            bl = vd.frame_blocks(d, world)
            assert len(bl) == world and bl[0][0] == 0 and bl[-1][1] == d
            assert all(bl[i][1] == bl[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in bl]
            assert max(sizes) - min(sizes) <= 1
            plan = vd.chain_plan(d, world)
            mid = d // 2                                                             # morph.cu:1374
            assert plan["mid"] == mid
            assert plan["forward"][1] == list(range(mid + 1, d)) and plan["backward"][1] == list(range(mid - 1, -1, -1))
            assert sorted(plan["forward"][1] + plan["backward"][1] + [mid]) == list(range(d))
            assert plan["forward"][0] == 0 and plan["backward"][0] == (1 if world > 1 else 0)


def test_world2_gloo_gather_and_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(7 * 6, dtype=np.float32).reshape(7, 3, 2) * 2.0
    for rank, whole, units, secs in res:
        np.testing.assert_array_equal(whole, want)
        assert units == 30.0 and secs == 1.5
