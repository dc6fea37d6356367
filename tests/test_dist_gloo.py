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
    # frame-sharded QuadraticPath: every rank solves its block (stand-in solver), everyone gets all frames back in order
    vecs = np.arange(d * 4 * 5 * 2, dtype=np.float32).reshape(d, 4, 5, 2)
    fake = lambda v: (v * np.float32(3.0) + np.float32(1.0), np.full((v.shape[0], 2), 7 + rank, np.int32))
    qp, its = vd.quadratic_path_sharded(vecs, solve=fake)
    assert qp.shape == vecs.shape and np.array_equal(qp, vecs * np.float32(3.0) + np.float32(1.0))
    assert [int(x) for x in its[:, 0]] == [7 + r for r, (bb, ee) in enumerate(blocks) for _ in range(ee - bb)]
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


# ------------------------------------------------------------------------------------------ level pipeline (dist.run_pipeline)
class _FakeEngine:
    """Stand-in for dist.MorphEngine on CPU tensors: every operation is a cheap deterministic function of exactly the
    inputs the real one reads (the coarser level's frame, the chain neighbour's final frame), and asserts that those
    inputs are final when it runs -- so a wrong hand-off order or a missing exchange changes the result or trips."""
    P = 5

    def __init__(self, depths):
        import torch
        self.depths = list(depths)
        n = len(depths)
        self.v = [torch.full((depths[l], self.P), float("nan"), dtype=torch.float64) for l in range(n)]
        self.final = [set() for _ in range(n)]          # frames of level l whose v is final
        self.ready = [set() for _ in range(n)]          # frames prolonged + initialised
        self.temp = {}

    def coarse_solve(self):
        import torch
        l = len(self.depths) - 1
        self.v[l] = torch.arange(self.depths[l] * self.P, dtype=torch.float64).reshape(self.depths[l], self.P) * 0.25 + 1
        self.final[l] = set(range(self.depths[l]))

    def _up(self, l, i):
        dc, df = self.depths[l + 1], self.depths[l]
        if dc == df:
            need = [i]
            val = self.v[l + 1][i] * 2 + 1
        else:                                            # temporal in-fill: both coarse neighbours
            a, b = i // 2, min((i + 1) // 2, dc - 1)
            need = [a, b]
            val = self.v[l + 1][a] * 2 + self.v[l + 1][b] * 0.5
        assert all(f in self.final[l + 1] for f in need), f"level {l} frame {i}: coarse frames {need} not final"
        self.v[l][i] = val
        self.ready[l].add(i)

    def upsample(self, l):
        for i in range(self.depths[l]):
            dc, df = self.depths[l + 1], self.depths[l]
            need = [i] if dc == df else [i // 2, min((i + 1) // 2, dc - 1)]
            if all(f in self.final[l + 1] for f in need):
                self._up(l, i)                           # frames of the other chain may be missing on this rank: never used
    def initialize(self, l): pass
    def upsample_frames(self, l, i): self._up(l, i)
    def initialize_frames(self, l, i): assert i in self.ready[l]

    def init_temp(self, l, i, direction):
        assert (i + direction) in self.final[l], f"level {l} frame {i}: neighbour {i + direction} not final"
        self.temp[(l, i)] = self.v[l][i + direction] * 3 + 0.125

    def optimize_frame(self, l, i, flag, max_iter):
        assert i in self.ready[l] and i not in self.final[l]
        self.v[l][i] = self.v[l][i] * 1.5 + (self.temp.pop((l, i)) if flag else 0.0) + max_iter * 0.001 + l
        self.final[l].add(i)
        return 1

    def optimize_chains(self, l, max_iter, chains):
        d = self.depths[l]; mid = d // 2
        self.optimize_frame(l, mid, False, max_iter)
        for i in (range(mid + 1, d) if chains & 1 else []):
            self.init_temp(l, i, -1); self.optimize_frame(l, i, True, max_iter)
        for i in (range(mid - 1, -1, -1) if chains & 2 else []):
            self.init_temp(l, i, 1); self.optimize_frame(l, i, True, max_iter)

    def new_pages(self, l, n=1):
        import torch
        return torch.zeros(n * self.P * 8, dtype=torch.uint8)
    def get_pages(self, l, a, b):
        assert all(f in self.final[l] for f in range(a, b))
        return self.v[l][a:b].contiguous().view(-1).view(dtype=__import__("torch").uint8).clone()
    def set_pages(self, l, a, t):
        import torch
        k = t.numel() // (self.P * 8)
        self.v[l][a:a + k] = t.view(dtype=torch.float64).reshape(k, self.P)
        self.final[l].update(range(a, a + k)); self.ready[l].update(range(a, a + k))
    def sync(self): pass


def _pipeline_reference(depths, max_iter0, drop):
    e = _FakeEngine(depths)
    n = len(depths)
    e.coarse_solve()
    mi = np.float32(max_iter0)
    for l in range(n - 2, 0, -1):
        e.upsample(l); e.initialize(l); e.optimize_chains(l, float(mi), 3)
        mi = np.float32(mi / np.float32(drop))
    return e.v[1].numpy().copy()


def _pipeline_worker(rank, world, port, depths, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    vd.init(backend="gloo")
    e = _FakeEngine(depths)
    plan = vd.run_pipeline(e, 1000.0, 1.5, rank, world)
    dist.barrier()
    q.put((rank, plan["nstages"], e.v[1].numpy().copy() if rank in (0, 1) else None))
    dist.destroy_process_group()


def test_pipeline_plan_shapes():
    from videomorphing_b200 import dist as vd
    cfg4 = [120, 120, 120, 120, 120, 120, 61, 31, 16]                 # 720p x 120, cap lifted
    p8 = vd.pipeline_plan(cfg4, 8)
    assert p8["nstages"] == 4 and sorted(p8["ranks"]) == list(range(8))
    assert [p8["ranks"][r]["levels"] for r in (0, 2, 4, 6)] == [[1], [2], [3], [7, 6, 5, 4]]
    assert all(p8["ranks"][r]["send_to"] == (r - 2 if r >= 2 else None) for r in range(8))
    assert all(p8["ranks"][r]["recv_from"] == (r + 2 if r < 6 else None) for r in range(8))
    assert vd.pipeline_plan(cfg4, 4)["nstages"] == 2 and vd.pipeline_plan(cfg4, 2)["nstages"] == 1 and vd.pipeline_plan(cfg4, 1)["nstages"] == 1
    # every optimised level is owned by exactly one pair, for any world size
    for world in (1, 2, 3, 4, 6, 8, 16):
        for depths in (cfg4, [9, 9, 9, 9, 5], [16, 16, 16, 9, 5, 3], [1, 1, 1, 1], [5, 5, 3]):
            pl = vd.pipeline_plan(depths, world)
            owned = sorted(l for r, e in pl["ranks"].items() if e["dir"] == 0 for l in e["levels"])
            assert owned == list(range(1, len(depths) - 1))
            for r, e in pl["ranks"].items():
                if e["recv_from"] is not None:           # streamed-into levels have the depth of the level above
                    assert len(e["levels"]) == 1 and depths[e["levels"][0]] == depths[e["levels"][0] + 1]
    assert vd.chain_frames(7, 0) == [3, 4, 5, 6] and vd.chain_frames(7, 1) == [3, 2, 1, 0] and vd.chain_frames(1, 0) == [0]


@pytest.mark.parametrize("world,depths", [(4, [9, 9, 9, 9, 5]), (6, [16, 16, 16, 16, 9, 5, 3]), (4, [6, 6, 6, 4, 3])])
def test_level_pipeline_hand_offs_on_gloo(world, depths):
    """dist.run_pipeline on `world` CPU ranks with a stand-in engine: the frames each stage reads are final when it reads
    them, and ranks 0 and 1 end with exactly the level-1 field of the sequential schedule."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, depths, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _pipeline_reference(depths, 1000.0, 1.5)
    assert np.isfinite(want).all()
    for rank, nst, v1 in res:
        assert nst == world // 2
        if rank in (0, 1):
            np.testing.assert_array_equal(v1, want)
