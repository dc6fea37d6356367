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


# ------------------------------------------------------------------------------------------ wavefront over ranks (dist.run_wavefront)
class _FakeEngine:
    """Stand-in for dist.MorphEngine on CPU tensors: every operation is a cheap deterministic function of exactly the
    inputs the real one reads (the coarser level's frame, the chain neighbour's final frame), and asserts that those
    inputs are final when it runs -- so a wrong hand-off order or a missing exchange changes the result or trips."""
    P = 5

    def __init__(self, depths, max_iter0=1000.0, drop=1.5):
        import torch
        from videomorphing_b200 import dist as vd
        self.depths = list(depths)
        n = len(depths)
        self.dims = {l: (16 << (n - l), 9 << (n - l)) for l in range(n)}
        self.max_iters = vd.level_max_iters(max_iter0, drop, n)
        self.v = [torch.full((depths[l], self.P), float("nan"), dtype=torch.float64) for l in range(n)]
        self.final = [set() for _ in range(n)]          # frames of level l whose v is final
        self.ready = [set() for _ in range(n)]          # frames prolonged + initialised
        self.temp = {}
        self.launches = []

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

    def _init_temp(self, l, i, direction):
        assert (i + direction) in self.final[l], f"level {l} frame {i}: neighbour {i + direction} not final"
        self.temp[(l, i)] = self.v[l][i + direction] * 3 + 0.125

    def _optimize(self, l, i, flag, max_iter):
        assert i in self.ready[l] and i not in self.final[l], (l, i)
        self.v[l][i] = self.v[l][i] * 1.5 + (self.temp.pop((l, i)) if flag else 0.0) + max_iter * 0.001 + l
        self.final[l].add(i)

    def _whole_level(self, l):
        d = self.depths[l]; mid = d // 2
        for i in range(d):
            self._up(l, i)
        self._optimize(l, mid, False, self.max_iters[l])
        for i in range(mid + 1, d):
            self._init_temp(l, i, -1); self._optimize(l, i, True, self.max_iters[l])
        for i in range(mid - 1, -1, -1):
            self._init_temp(l, i, 1); self._optimize(l, i, True, self.max_iters[l])

    def prepare(self):
        import torch
        from videomorphing_b200 import dist as vd
        n = len(self.depths)
        l = n - 1
        self.v[l] = torch.arange(self.depths[l] * self.P, dtype=torch.float64).reshape(self.depths[l], self.P) * 0.25 + 1
        self.final[l] = set(range(self.depths[l]))
        K = vd.wavefront_head(self.depths)
        for l in range(n - 2, K, -1):
            self._whole_level(l)
        for i in range(self.depths[K]):
            self._up(K, i)
        return K

    def prep_frame(self, l, i, head, first, tdir):
        if not head:
            self._up(l, i)
        assert i in self.ready[l]                        # (the real engine initialises the frame here, head level included)
        if not first:
            self._init_temp(l, i, tdir)

    def enqueue_jobs(self, jobs):
        assert len({(l, f) for l, f, _, _ in jobs}) == len(jobs) <= 16
        for l, f, flag, mi in jobs:                      # the jobs of a launch are independent of each other
            assert not flag or (l, f) in self.temp
        for l, f, flag, mi in jobs:
            self._optimize(l, f, flag, mi)
        self.launches.append(len(jobs))

    def collect(self): pass

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


def _sequential_reference(depths):
    e = _FakeEngine(depths)
    K = e.prepare()
    for l in range(K, 0, -1):
        if l < K:
            for i in range(depths[l]):
                e._up(l, i)
        d = depths[l]; mid = d // 2
        e._optimize(l, mid, False, e.max_iters[l])
        for i in range(mid + 1, d):
            e._init_temp(l, i, -1); e._optimize(l, i, True, e.max_iters[l])
        for i in range(mid - 1, -1, -1):
            e._init_temp(l, i, 1); e._optimize(l, i, True, e.max_iters[l])
    return e.v[1].numpy().copy()


def _wavefront_worker(rank, world, port, depths, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    vd.init(backend="gloo")
    e = _FakeEngine(depths)
    plan = vd.run_wavefront(e, rank, world)
    dist.barrier()
    q.put((rank, plan["K"], e.v[1].numpy().copy(), e.launches))
    dist.destroy_process_group()


def test_wavefront_plan_shapes():
    from videomorphing_b200 import dist as vd
    cfg4 = [120, 120, 120, 120, 120, 120, 61, 31, 16]                 # 720p x 120, cap lifted
    dims = {l: (1280 >> (l - 1), 720 >> (l - 1)) for l in range(1, 9)}
    mi = vd.level_max_iters(1000, 2, 9)
    assert vd.wavefront_head(cfg4) == 5 and vd.wavefront_head([1, 1, 1, 1]) == 2 and vd.wavefront_head([9, 9, 9, 9, 5]) == 3
    assert [mi[l] for l in (7, 6, 5, 4, 3, 2, 1)] == [1000.0, 500.0, 250.0, 125.0, 62.5, 31.25, 15.625]
    for world in (2, 3, 4, 6, 8):
        for depths in (cfg4, [9, 9, 9, 9, 5], [16, 16, 16, 9, 5, 3], [5, 5, 3]):
            dm = {l: (64 << (len(depths) - l), 36 << (len(depths) - l)) for l in range(len(depths))}
            pl = vd.wavefront_plan(depths, dm, vd.level_max_iters(1000, 2, len(depths)), world)
            K = pl["K"]
            assert sorted(pl["owner"]) == sorted((l, dr) for l in range(1, K + 1) for dr in (0, 1))     # every chain has exactly one owner
            g0 = (world + 1) // 2
            assert all((r < g0) == (dr == 0) for (l, dr), r in pl["owner"].items())                     # a rank serves one direction
            assert pl["owner"][(1, 0)] == 0 and pl["owner"][(1, 1)] == g0                                # level 1 on the directions' first ranks
            for dr in (0, 1):                                                                            # contiguous level groups, finest first
                owners = [pl["owner"][(l, dr)] for l in range(1, K + 1)]
                assert owners == sorted(owners)
    p8 = vd.wavefront_plan(cfg4, dims, mi, 8)
    assert sorted(p8["groups"]) == list(range(8))                                                        # 720p x 120 on 8 GPUs: nobody idles
    assert vd.chain_plan(7, 2)["mid"] == 3


@pytest.mark.parametrize("world,depths", [(2, [9, 9, 9, 9, 5]), (4, [9, 9, 9, 9, 5]), (6, [16, 16, 16, 16, 9, 5, 3]), (3, [6, 6, 6, 4, 3]), (8, [7, 7, 7, 7, 7, 7, 4])])
def test_wavefront_hand_offs_on_gloo(world, depths):
    """dist.run_wavefront on `world` CPU ranks with a stand-in engine: the frames each chain reads are final when it reads
    them, every launch holds independent jobs only, and EVERY rank ends with exactly the level-1 field of the sequential
    schedule."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wavefront_worker, args=(r, world, port, depths, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _sequential_reference(depths)
    assert np.isfinite(want).all()
    for rank, K, v1, launches in res:
        np.testing.assert_array_equal(v1, want)


class _FakeArrays:
    """Stand-in for dist.LevelArrays: byte arrays in host memory; a rank starts with its own frame block only."""

    def __init__(self, sizes, d, blocks, rank):
        import torch
        self.want = {k: (torch.arange(n, dtype=torch.int64) * (7 + k[0]) % 251).to(torch.uint8) for k, n in sizes.items()}
        self.data = {}
        a, b = blocks[rank]
        for k, n in sizes.items():
            per = n // d
            t = torch.full((n,), 255, dtype=torch.uint8)                 # frames this rank did not build: garbage
            t[a * per: b * per] = self.want[k][a * per: b * per]
            self.data[k] = t

    def nbytes(self, l, name): return self.data[(l, name)].numel()
    def read(self, l, name, off, t): t.copy_(self.data[(l, name)][off: off + t.numel()])
    def write(self, l, name, off, t): self.data[(l, name)][off: off + t.numel()] = t


def _exchange_worker(rank, world, port, d, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from videomorphing_b200 import dist as vd
    vd.init(backend="gloo")
    blocks = vd.frame_blocks(d, world)
    sizes = {(1, "img0"): d * 40, (1, "f0"): d * 80, (2, "img0"): d * 12, (2, "keep0"): d * 36}
    arr = _FakeArrays(sizes, d, blocks, rank)
    vd.exchange_frames(arr, list(sizes), d, blocks, rank, "cpu")
    ok = all(bool((arr.data[k] == arr.want[k]).all()) for k in sizes)
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,d", [(2, 8), (3, 7), (4, 9)])
def test_frame_sharded_build_exchange_on_gloo(world, d):
    """dist.exchange_frames (the all-gather of dist.build_pyramid): equal and ragged frame blocks, every rank ends with every frame."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res) and len(res) == world


def test_c_abi_wavefront_plan_equals_the_python_plan():
    """vm_wavefront_plan (what tools/vmorph_video.cpp and any non-Python host use) == dist.wavefront_plan."""
    import ctypes as C
    from videomorphing_b200 import _lib, dist as vd
    L = _lib.load()
    shapes = [(1280, 720, 120, 8, 1 << 62), (1280, 720, 120, 8, 14000000), (3840, 2160, 16, 8, 1 << 62), (96, 64, 9, 4, 1 << 62), (640, 360, 33, 8, 1 << 62)]
    for w, h, d, sr, cap in shapes:
        whd = (C.c_int32 * (3 * 32))()
        n = L.vm_level_schedule(w, h, d, sr, cap, 32, whd, None)
        assert n >= 3
        depths = [whd[3 * l + 2] for l in range(n)]
        dims = {l: (whd[3 * l], whd[3 * l + 1]) for l in range(n)}
        mi = vd.level_max_iters(1000, 2, n)
        mia = (C.c_float * n)(*[mi.get(l, 0.0) for l in range(n)])
        for world in (1, 2, 3, 4, 5, 8):
            owner = (C.c_int32 * (2 * n))()
            K = L.vm_wavefront_plan(n, whd, mia, world, owner)
            pl = vd.wavefront_plan(depths, dims, mi, world)
            assert K == pl["K"]
            got = {(l, dr): owner[2 * l + dr] for l in range(n) for dr in (0, 1) if owner[2 * l + dr] >= 0}
            assert got == pl["owner"], (w, h, d, world)
