"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

What shards on this path:
  * image pairs (configs 1-3) do not shard: N GPUs = N independent replicas, no data-path collective;
  * the frame-parallel stages of a video (pyramid image levels, render, QuadraticPath) split into contiguous frame
    blocks, no exchange; results are gathered only when one host wants all frames;
  * the optimizer is two sequential frame chains per level (forward / backward from the middle frame,
    morph.cu:1374-1439).  Exact mode (same arithmetic as one GPU, bit-identical result): the direction x level WAVEFRONT that
    vm_morph_run executes on one GPU (frame i of level l needs frame i of level l+1 and frame i -/+ 1 of level l, so the
    levels work one chain position behind each other and the two directions are independent) is split over the ranks by
    (direction, contiguous group of levels); every tick each rank runs ONE multi-job launch over its chains and hands the
    frames it finished to the rank that owns the next finer level of the same direction -- NCCL send / recv of one `v`
    page per frame and link, stream-ordered, no collective on the path (wavefront_plan / run_wavefront).
No collective is used unless a stage really exchanges data; timing reductions are scalar (MAX of seconds, SUM of units).
"""
import datetime
import os

import numpy as np
import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None, device_id=None):
    """init_process_group from the torchrun environment (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT). No-op for world 1."""
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"timeout": datetime.timedelta(seconds=int(os.environ.get("VMORPH_DIST_TIMEOUT_S", "300")))}   # fail fast instead of hanging a GPU box
        if backend == "nccl" and device_id is not None:
            kw["device_id"] = torch.device("cuda", device_id)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def frame_blocks(d, world):
    """Contiguous frame blocks [(begin, end), ...] per rank for the frame-parallel stages; sizes differ by at most one."""
    base, rem = divmod(d, world)
    out, b = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((b, b + n))
        b += n
    return out


def chain_plan(d, world):
    """Exact-mode optimizer plan for a level of depth d: {"mid": frame, "forward": (rank, [frames]), "backward": (rank, [frames])}.
    The middle frame is optimised by every chain owner (deterministic => identical), each owner then walks its chain;
    afterwards the owners exchange the `v` pages of their frames.  Ranks >= 2 own no chain (the chain is sequential)."""
    mid = d // 2
    fwd = list(range(mid + 1, d))
    bwd = list(range(mid - 1, -1, -1))
    return {"mid": mid, "forward": (0, fwd), "backward": (min(1, world - 1), bwd)}


def reduce_throughput(units, seconds, device=None):
    """Whole-job throughput: SUM of the units all ranks processed / MAX over ranks of the time.  Returns (units, seconds)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units), float(seconds)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())


def gather_frames(local, d, device=None):
    """All ranks contribute their frame block (frame_blocks order) of a (n_local, ...) array; every rank gets the (d, ...) whole."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.ascontiguousarray(local)
    world = dist.get_world_size()
    blocks = frame_blocks(d, world)
    nmax = max(e - b for b, e in blocks)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    loc = np.ascontiguousarray(local)
    pad = np.zeros((nmax,) + loc.shape[1:], loc.dtype)
    pad[: loc.shape[0]] = loc
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[: e - b] for o, (b, e) in zip(outs, blocks)], 0)


def render_frames_sharded(w, h, ex, ext0, ext1, vectors, qpaths=None, color_from=1, device=0, gather=True):
    """RenderWidget-style pass over all d frames (geo_fa = color_fa = smoothstep(frame / (d-1)), UI/RenderWidget.cpp:93-96),
    each rank rendering its frame block on its own GPU.  ext0/ext1: (d, h+2ex, w+2ex, 4) u8; vectors: (d,h,w,2)."""
    from . import api, synth
    d = vectors.shape[0]
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b, e = frame_blocks(d, world)[rank]
    out = np.zeros((e - b, h, w, 3), np.uint8)
    for z in range(b, e):
        fa = float(synth.smoothstep(z / max(1, d - 1)))
        out[z - b] = api.render_halfway_image(w, h, ex, fa, fa, color_from, ext0[z], ext1[z], vectors[z],
                                              None if qpaths is None else qpaths[z], device=device)
    return gather_frames(out, d) if gather else out


def quadratic_path_sharded(vectors, max_iter=10000, tol=1e-12, device=0, gather=True, solve=None):
    """CQuadraticPath::optimize (QuadraticPath.cpp:24-223 loops z over the frames; the frames are independent): each rank
    solves its contiguous frame block on its own GPU, no exchange; with gather=True every rank gets all d frames back.
    Returns (qpaths, iterations (d, 2)).  `solve` replaces the GPU call in the CPU tests."""
    from . import api
    solve = solve or (lambda v: api.quadratic_path_frames(v, max_iter, tol, device=device))
    d = vectors.shape[0]
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    b, e = frame_blocks(d, world)[rank]
    if e > b:
        q, it = solve(np.ascontiguousarray(vectors[b:e]))
    else:
        q, it = np.zeros((0,) + vectors.shape[1:], np.float32), np.zeros((0, 2), np.int32)
    if not gather:
        return q, it
    return gather_frames(q, d), gather_frames(np.asarray(it, np.int32), d)


# ------------------------------------------------------------------------------------------ Pyramid::build, sharded by frame
class LevelArrays:
    """Byte ranges of a vm.Pyramid's device arrays, for the exchange of build_pyramid (the CPU tests use a stand-in)."""

    def __init__(self, pyramid, device):
        self.p, self.device = pyramid, device
        from . import _lib
        self._lib, self.L = _lib, _lib.load()

    def nbytes(self, l, name):
        return self.p.dev_ptr(l, name)[1]

    def read(self, l, name, off, t):                 # device array [off, off + len(t)) -> t (on this GPU, or pinned host memory)
        base = self.p.dev_ptr(l, name)[0]
        f = self.L.vm_dev_copy if t.is_cuda else self.L.vm_dev_download
        self._lib.check(f(self.device, t.data_ptr(), base + off, t.numel(), None))
        if not t.is_cuda:
            self._lib.check(self.L.vm_stream_sync(self.device, None))

    def write(self, l, name, off, t):
        base = self.p.dev_ptr(l, name)[0]
        f = self.L.vm_dev_copy if t.is_cuda else self.L.vm_dev_upload
        self._lib.check(f(self.device, base + off, t.data_ptr(), t.numel(), None))
        if not t.is_cuda:
            self._lib.check(self.L.vm_stream_sync(self.device, None))


def exchange_frames(arrays, items, d, blocks, rank, staging):
    """Every array in `items` [(level, name), ...] holds d contiguous frames of which this rank built blocks[rank]; afterwards
    every rank holds all of them.  NCCL: one in-place all-gather per array over NVLink (equal blocks) or one broadcast per
    block; gloo (CPU tests, several ranks on one GPU): broadcasts of host staging buffers."""
    world = len(blocks)
    a, b = blocks[rank]
    equal = len({e - s for s, e in blocks}) == 1
    for l, name in items:
        n = arrays.nbytes(l, name)
        per = n // d
        t = torch.empty(n, dtype=torch.uint8, device=staging, pin_memory=(staging == "cpu" and torch.cuda.is_available()))
        if b > a:
            arrays.read(l, name, a * per, t[a * per: b * per])
        if equal and t.is_cuda:
            dist.all_gather_into_tensor(t, t[a * per: b * per])
        else:
            for r, (s, e) in enumerate(blocks):
                if e > s:
                    dist.broadcast(t[s * per: e * per], r)
        if a > 0:
            arrays.write(l, name, 0, t[: a * per])
        if b < d:
            arrays.write(l, name, b * per, t[b * per:])


def build_pyramid(pyramid, video0, video1, flows=None, start_res=8, voxel_cap=None, device=0, stream=None):
    """Pyramid::build (pyramid.cu:166-485) on all ranks' GPUs: the levels that keep every frame are per-frame work
    (pyramid.cu:267-403), so each rank uploads and resamples its contiguous frame block only, the blocks are all-gathered
    (gray images, the four flow fields, and the linear-light planes of the last such level), and the temporally halved
    levels (pyramid.cu:406-459: frame t composes the flows of frames 2t and 2t +- 1) are then built by every rank from
    the complete level below -- they are tiny.  Every rank ends with the pyramid a one-GPU build gives, bit for bit."""
    from . import api
    cap = api.REFERENCE_VOXEL_CAP if voxel_cap is None else voxel_cap
    world = dist.get_world_size() if dist.is_initialized() else 1
    d = video0.shape[0]
    if world == 1 or d < world:
        return pyramid.build(video0, video1, flows, start_res=start_res, voxel_cap=cap, stream=stream)
    rank = dist.get_rank()
    blocks = frame_blocks(d, world)
    a, b = blocks[rank]
    K = pyramid.build_frames(video0, video1, flows, a, b - a, start_res=start_res, voxel_cap=cap, stream=stream)
    names = ["img0", "img1"] + (["f0", "f1", "b0", "b1"] if flows is not None else [])
    items = [(l, nm) for l in range(1, K + 1) for nm in names] + [(K, "keep0"), (K, "keep1")]
    staging = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
    arrays = LevelArrays(pyramid, device)
    arrays._lib.check(arrays.L.vm_stream_sync(device, stream))       # the exchange runs on the default stream
    exchange_frames(arrays, items, d, blocks, rank, staging)
    return pyramid.build_finish(stream=stream)


# ------------------------------------------------------------------------------------------ optimizer, exact mode
def wavefront_head(depths):
    """Head level K of the wavefront: the coarsest level whose finer levels all have its depth (no temporal in-fill between
    them, upsample.cu:297-338); at most 8 stages.  depths[l] = frames of pyramid level l (l = 0 .. n-1)."""
    n = len(depths)
    K = 1
    while K + 1 <= n - 2 and depths[K] == depths[K + 1]:
        K += 1
    return min(K, 8)


def level_max_iters(max_iter0, drop, n):
    """_max_iter of every optimised level (morph.cu:131,163): float32 like the reference."""
    out, mi = {}, np.float32(max_iter0)
    for l in range(n - 2, 0, -1):
        out[l] = float(mi)
        mi = np.float32(mi / np.float32(drop))
    return out


def _group_cost(levels, dims, max_iters):
    """Relative cost of one tick of a lock-step launch over `levels` (one frame each): the throughput terms add up, the
    latency terms (rounds of the slowest level) overlap.  Fitted to the 720p measurements (profiles/r2_wavefront.md)."""
    thr = sum(dims[l][0] * dims[l][1] * min(max_iters[l], 16.0) * 1e-6 for l in levels)
    lat = max(min(max_iters[l], 50.0) * 16 * 0.02 for l in levels)
    return thr + lat


def wavefront_plan(depths, dims, max_iters, world):
    """Who runs what.  Chains are (level, direction) for the levels K .. 1; direction 0 = the middle frame and the frames
    after it, 1 = the frames before it.  Direction 0 goes to the first ceil(world / 2) ranks, direction 1 to the others
    (one rank: both); within a direction the levels are split into contiguous groups, one per rank, minimising the most
    expensive group (_group_cost).  Returns {"K", "owner": {(level, dir): rank}, "groups": {rank: (dir, [levels])}}."""
    K = wavefront_head(depths)
    levels = list(range(K, 0, -1))
    g0 = (world + 1) // 2
    ranks_of = {0: list(range(g0)), 1: list(range(g0, world)) if world > 1 else [0]}

    def split(g):                                   # best contiguous partition of `levels` into at most g groups
        g = max(1, min(g, len(levels)))
        best = None

        def rec(start, left, acc):
            nonlocal best
            if left == 1:
                cand = acc + [levels[start:]]
                cost = sorted((_group_cost(c, dims, max_iters) for c in cand), reverse=True)   # most expensive group first, then the next ...
                if best is None or cost < best[0]:
                    best = (cost, cand)
                return
            for end in range(start + 1, len(levels) - left + 2):
                rec(end, left - 1, acc + [levels[start:end]])
        rec(0, g, [])
        return best[1]
    owner, groups = {}, {}
    for dr in (0, 1):
        rk = ranks_of[dr]
        parts = split(len(rk))
        # the finest group goes to the direction's first rank: ranks 0 and ceil(world / 2) end up owning level 1
        for r, part in zip(rk, reversed(parts)):
            for l in part:
                owner[(l, dr)] = r
            groups.setdefault(r, []).append((dr, part))
    return {"K": K, "owner": owner, "groups": groups}


class MorphEngine:
    """The operations the wavefront needs, on a vm.Morph / vm.Pyramid of this rank's GPU."""

    def __init__(self, morph, pyramid, device, params=None):
        self.m, self.p, self.device = morph, pyramid, device
        from . import _lib
        self._lib, self.L = _lib, _lib.load()
        n = pyramid.num_levels
        self.depths = [pyramid.info(l)["d"] for l in range(n)]
        self.dims = {l: (pyramid.info(l)["w"], pyramid.info(l)["h"]) for l in range(n)}
        prm = params if params is not None else morph.params
        self.max_iters = level_max_iters(prm.max_iter, prm.max_iter_drop_factor, n)

    def prepare(self):
        """Everything above the wavefront, redundantly on every rank (deterministic => identical, no exchange): coarse solve,
        the temporally subsampled levels, head level prolonged.  Returns the head level K."""
        return self.m.wavefront_prepare()

    def prep_frame(self, l, i, head, first, tdir):
        if not head:                                  # the head level was prolonged whole by prepare() (temporal in-fill)
            self.m.upsample_frames(l, i, 1)
        self.m.initialize_frames(l, i, 1)
        if not first:
            self.m.initialize_temp(l, i, tdir)

    def enqueue_jobs(self, jobs): self.m.enqueue_jobs(jobs)
    def collect(self): self.m.collect()

    def _page(self, l):
        base, _ = self.p.dev_ptr(l, "v")
        return base, self.p.info(l)["pagestride"] * 8

    @property
    def staging(self):
        """Where hand-off buffers live: on this GPU for NCCL (send / recv over NVLink), in pinned host memory for gloo
        (the CPU tests and the several-ranks-on-one-GPU test: gloo has no CUDA send / recv)."""
        if dist.is_available() and dist.is_initialized() and dist.get_backend() != "nccl":
            return "cpu"
        return f"cuda:{self.device}"

    def new_pages(self, l, n=1):
        st = self.staging
        return torch.empty(n * self._page(l)[1], dtype=torch.uint8, device=st, pin_memory=(st == "cpu" and torch.cuda.is_available()))

    def get_pages(self, l, a, b):
        """Copy of the `v` pages [a, b) of level l (one contiguous byte tensor on this GPU, or in host memory under gloo)."""
        base, pb = self._page(l)
        t = self.new_pages(l, b - a)
        if t.is_cuda:
            self._lib.check(self.L.vm_dev_copy(self.device, t.data_ptr(), base + a * pb, (b - a) * pb, None))
        else:
            self._lib.check(self.L.vm_dev_download(self.device, t.data_ptr(), base + a * pb, (b - a) * pb, None))
            self._lib.check(self.L.vm_stream_sync(self.device, None))
        return t

    def set_pages(self, l, a, t):
        base, pb = self._page(l)
        if t.is_cuda:
            self._lib.check(self.L.vm_dev_copy(self.device, base + a * pb, t.data_ptr(), t.numel(), None))
        else:
            self._lib.check(self.L.vm_dev_upload(self.device, base + a * pb, t.data_ptr(), t.numel(), None))
            self._lib.check(self.L.vm_stream_sync(self.device, None))
        self._lib.check(self.L.vm_level_mark_v_valid(self.p.h, l))

    def sync(self):
        self._lib.check(self.L.vm_stream_sync(self.device, None))
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()


def run_wavefront(eng, rank, world):
    """The direction x level wavefront on this rank (see wavefront_plan).  `eng` is a MorphEngine (or, in the CPU tests, a
    stand-in with the same methods).  Chain (l, dr) works on position c (frame mid + c forward, mid - c backward) at tick
    (K - l) + c: the frame of level l + 1 it prolongs and its chain neighbour were finished one tick earlier.  A backward
    chain whose level's forward chain lives on another rank optimises the middle frame itself (position 0; deterministic
    => the same bits), so the two directions never exchange anything.  Afterwards EVERY rank holds the whole level-1 field."""
    depths = eng.depths
    plan = wavefront_plan(depths, eng.dims, eng.max_iters, world)
    K, owner = plan["K"], plan["owner"]
    assert world >= 2 and all(len({dr for dr, _ in g}) == 1 for g in plan["groups"].values())
    assert eng.prepare() == K
    d = depths[K]
    mid = d // 2
    npos = {0: d - mid, 1: mid + 1}                              # chain positions incl. the middle frame
    mine = sorted([c for c, r in owner.items() if r == rank], key=lambda c: (-c[0], c[1]))

    def active(l, dr, c):                                        # does chain (l, dr) run position c?  (every rank serves one direction,
        return 0 <= c < npos[dr]                                 #  so a backward chain always optimises the middle frame itself)
    frame = lambda dr, c: mid + c if dr == 0 else mid - c
    sends = []
    for T in range(K - 1 + max(npos.values())):
        work = [(l, dr, T - (K - l)) for l, dr in mine if active(l, dr, T - (K - l))]
        # ---- receive the frames of the next coarser level that another rank finished one tick ago
        ops, incoming = [], []
        for l, dr, c in work:
            if l < K and owner[(l + 1, dr)] != rank:
                t = eng.new_pages(l + 1)
                ops.append(dist.P2POp(dist.irecv, t, owner[(l + 1, dr)]))
                incoming.append((l + 1, frame(dr, c), t))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        for l1, f, t in incoming:
            eng.set_pages(l1, f, t)
        # ---- prolong / initialise / temporal reference, then ONE lock-step launch over this rank's chains
        jobs = []
        for l, dr, c in work:
            first = c == 0
            eng.prep_frame(l, frame(dr, c), l == K, first, -1 if dr == 0 else 1)
            jobs.append((l, frame(dr, c), not first, eng.max_iters[l]))
        if jobs:
            eng.enqueue_jobs(jobs)
        # ---- hand the finished frames to the owner of the next finer level of the same direction
        ops = []
        for l, dr, c in work:
            if l > 1 and owner[(l - 1, dr)] != rank and active(l - 1, dr, c):
                t = eng.get_pages(l, frame(dr, c), frame(dr, c) + 1)
                ops.append(dist.P2POp(dist.isend, t, owner[(l - 1, dr)]))
                sends.append(t)
        sends += (dist.batch_isend_irecv(ops) if ops else [])
    for w in sends:
        if hasattr(w, "wait"):
            w.wait()
    eng.collect()
    # ---- every rank gets the whole level-1 field: the two level-1 owners broadcast their halves
    o0, o1 = owner[(1, 0)], owner[(1, 1)]
    for src, a, b in ((o0, mid, d), (o1, 0, mid)):
        if b <= a:
            continue
        if world > 1:
            t = eng.get_pages(1, a, b) if rank == src else eng.new_pages(1, b - a)
            dist.broadcast(t, src)
            if rank != src:
                eng.set_pages(1, a, t)
    eng.sync()
    return plan


def optimize_video(morph, pyramid, params, device=0):
    """Morph::calculate_halfway_parametrization (morph.cu:150-168) for a video on 1 .. 8 GPUs, exact mode (the same
    arithmetic as one GPU; every rank ends with the bit-identical level-1 field).  One rank: vm_morph_run (the wavefront
    inside one GPU).  Several ranks: the same wavefront split by direction and level group (run_wavefront).  Every rank
    holds the whole pyramid (build_pyramid: built by frame block, all-gathered)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        morph.run()
        return morph
    run_wavefront(MorphEngine(morph, pyramid, device, params), rank, world)
    return morph
