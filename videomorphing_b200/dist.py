"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

What shards on this path:
  * image pairs (configs 1-3) do not shard: N GPUs = N independent replicas, no data-path collective;
  * the frame-parallel stages of a video (pyramid image levels, render, QuadraticPath) split into contiguous frame
    blocks, no exchange; results are gathered only when one host wants all frames;
  * the optimizer is two sequential frame chains per level (forward / backward from the middle frame,
    morph.cu:1374-1439).  Exact mode (same arithmetic as one GPU): with 2-3 ranks each chain gets a rank and the ranks swap
    their halves of `v` once per level; with 4+ ranks the levels of a chain additionally run as a wavefront on different
    ranks (pipeline_plan / run_pipeline): frame i of level l+1 is handed to the rank that owns level l as soon as it is
    final -- NCCL send / recv of one `v` page per frame and link, no collective.
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


# ------------------------------------------------------------------------------------------ optimizer, exact mode
def _optimize_video_two_chains(morph, pyramid, params, device=0):
    """Morph::calculate_halfway_parametrization (morph.cu:150-168) for a video with the two frame chains of every level on
    two GPUs (exact mode: the same arithmetic as one GPU, bit-identical result on ranks 0 and 1).

    Every rank holds the whole pyramid (built from the same frames).  Per level: upsample + initialise all frames (cheap,
    redundant), rank 0 optimises the middle frame and the forward chain, rank 1 the middle frame and the backward chain
    (the middle frame is deterministic, so both get the same bits), then they swap the `v` pages of their chains -- the
    only exchange on the path: one frame's vector field per frame and level, handed to the rank that needs it for the
    next level's prolongation.  Ranks >= 2 own no chain (the chain is sequential) and receive the level by broadcast.
    With one rank this is the ordinary run (both chains concurrently on two streams)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    n = pyramid.num_levels
    eng = MorphEngine(morph, pyramid, device)
    morph.cpu_optimize_level()
    max_iter = np.float32(params.max_iter)
    for l in range(n - 2, 0, -1):
        morph.upsample(l)
        morph.initialize_level(l)
        d = pyramid.info(l)["d"]
        plan = chain_plan(d, world)
        mid = plan["mid"]
        if world == 1:
            morph.optimize_chains(l, float(max_iter), 3)
        else:
            chains = (1 if rank == plan["forward"][0] else 0) | (2 if rank == plan["backward"][0] else 0)
            fwd, bwd = (mid + 1, d), (0, mid)
            if rank <= 1:
                morph.optimize_chains(l, float(max_iter), chains)
                _swap_pages(eng, l, fwd if rank == 0 else bwd, bwd if rank == 0 else fwd, rank ^ 1)
            if world > 2:
                t = eng.get_pages(l, 0, d) if rank == 0 else eng.new_pages(l, d)
                dist.broadcast(t, 0)
                if rank > 1:
                    eng.set_pages(l, 0, t)
                eng.sync()
        max_iter = np.float32(max_iter / np.float32(params.max_iter_drop_factor))     # float like the reference (morph.cu:163)
    return morph


# ------------------------------------------------------------------------------------------ optimizer, level pipeline
def pipeline_plan(depths, world):
    """Exact-mode plan for `world` ranks (direction x level wavefront).

    depths[l] = number of frames of pyramid level l (l = 0 .. n-1; levels 1 .. n-2 are optimised, n-1 is the dense solve).
    Ranks come in pairs: rank = 2*pair + direction (0 = forward chain mid+1.., 1 = backward chain mid-1..0).
    Pair 0 owns level 1, pair 1 level 2, ..., the last pair ("coarse") owns every remaining level: it runs them level by
    level as on two GPUs (pages swapped with its partner per level) and streams the frames of its finest level out one
    by one.  A pair that owns a single level receives frame i of the coarser level from the pair above (same
    direction), prolongs + initialises + optimises its own frame i, and hands it on: level l of frame i only needs level
    l+1 of frame i and level l of frame i -/+ 1, so the levels run as a wavefront behind each other.  A level can only be
    streamed into when it has the depth of the level above (no temporal in-fill between them); that bounds the number
    of stages.  Returns {"nstages", "ranks": {rank: {"pair", "dir", "levels", "recv_from", "send_to"}}}; ranks that
    own nothing are absent."""
    n = len(depths)
    nopt = n - 2                                     # optimised levels 1 .. n-2
    nst = 1
    while nst < world // 2 and nst < nopt and depths[nst] == depths[nst + 1]:
        nst += 1                                     # level `nst` may be streamed into from level nst+1
    if world < 2 or nopt < 1:
        nst = 1
    ranks = {}
    ndir = 2 if world >= 2 else 1
    for pair in range(nst):
        levels = [pair + 1] if pair < nst - 1 else list(range(nopt, pair, -1))       # coarse pair: n-2 .. nst
        for dr in range(ndir):
            r = 2 * pair + dr
            ranks[r] = {"pair": pair, "dir": dr, "levels": levels,
                        "recv_from": (r + 2) if pair < nst - 1 else None,
                        "send_to": (r - 2) if pair > 0 else None}
    return {"nstages": nst, "ranks": ranks}


def chain_frames(d, direction):
    """Frames of one chain in processing order, the middle frame first (morph.cu:1374-1439)."""
    mid = d // 2
    return [mid] + (list(range(mid + 1, d)) if direction == 0 else list(range(mid - 1, -1, -1)))


class MorphEngine:
    """The operations the level pipeline needs, on a vm.Morph / vm.Pyramid of this rank's GPU."""

    def __init__(self, morph, pyramid, device):
        self.m, self.p, self.device = morph, pyramid, device
        from . import _lib
        self._lib, self.L = _lib, _lib.load()
        self.depths = [pyramid.info(l)["d"] for l in range(pyramid.num_levels)]

    def coarse_solve(self): self.m.cpu_optimize_level()
    def upsample(self, l): self.m.upsample(l)
    def initialize(self, l): self.m.initialize_level(l)
    def optimize_chains(self, l, max_iter, chains): self.m.optimize_chains(l, max_iter, chains)
    def upsample_frames(self, l, i): self.m.upsample_frames(l, i, 1)
    def initialize_frames(self, l, i): self.m.initialize_frames(l, i, 1)
    def init_temp(self, l, i, direction): self.m.initialize_temp(l, i, direction)
    def optimize_frame(self, l, i, flag, max_iter): return self.m.optimize_frame(l, i, flag, max_iter)

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


def _swap_pages(eng, l, mine, theirs, peer):
    """Send this rank's pages `mine` = (a, b) of level l to `peer` and receive the peer's pages `theirs`."""
    ops, rt = [], None
    if mine[1] > mine[0]:
        ops.append(dist.P2POp(dist.isend, eng.get_pages(l, *mine), peer))
    if theirs[1] > theirs[0]:
        rt = eng.new_pages(l, theirs[1] - theirs[0])
        ops.append(dist.P2POp(dist.irecv, rt, peer))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    if rt is not None:
        eng.set_pages(l, theirs[0], rt)
    eng.sync()


_link_groups = {}


def _links(plan, world):
    """One 2-rank process group per hand-off link (downstream, upstream), created once per plan shape by ALL ranks in the
    same order.  Separate groups keep a stage's receive from the level above and its send to the level below on
    independent NCCL communicators / streams (P2P ops on one group are serialised in issue order)."""
    key = (world, plan["nstages"])
    if key not in _link_groups:
        g = {}
        for r in sorted(plan["ranks"]):
            to = plan["ranks"][r]["send_to"]
            if to is not None:
                g[(to, r)] = dist.new_group([to, r])
        _link_groups[key] = g
    return _link_groups[key]


def run_pipeline(eng, max_iter0, drop, rank, world):
    """The level pipeline on one rank (see pipeline_plan).  `eng` is a MorphEngine (or, in the CPU tests, a stand-in with the
    same methods).  Returns the plan; afterwards ranks 0 and 1 hold the complete level-1 field."""
    depths = eng.depths
    n = len(depths)
    plan = pipeline_plan(depths, world)
    links = _links(plan, world) if plan["nstages"] > 1 else {}
    me = plan["ranks"].get(rank)
    if me is None:
        return plan                                                      # this rank owns nothing
    g_up = links.get((rank, me["recv_from"]))
    g_down = links.get((me["send_to"], rank))
    nst, dr = plan["nstages"], me["dir"]
    partner = rank ^ 1 if world >= 2 else None
    max_iter, mi = {}, np.float32(max_iter0)
    for l in range(n - 2, 0, -1):                                        # morph.cu:163, float like the reference
        max_iter[l] = float(mi)
        mi = np.float32(mi / np.float32(drop))
    sends = []

    def stream_level(l, recv_from, send_to, prepared):
        """One chain of level l frame by frame; `prepared`: the whole level is already prolonged and initialised."""
        tdir = -1 if dr == 0 else 1                                      # the neighbour the temporal term looks at
        for i in chain_frames(depths[l], dr):
            if recv_from is not None:
                t = eng.new_pages(l + 1)
                dist.recv(t, recv_from, group=g_up)
                eng.set_pages(l + 1, i, t)
            if not prepared:
                eng.upsample_frames(l, i)
                eng.initialize_frames(l, i)
            mid = i == depths[l] // 2
            if not mid:
                eng.init_temp(l, i, tdir)
            eng.optimize_frame(l, i, not mid, max_iter[l])
            if send_to is not None:
                t = eng.get_pages(l, i, i + 1)
                sends.append((t, dist.isend(t, send_to, group=g_down)))

    if me["pair"] == nst - 1:                                            # coarse pair
        eng.coarse_solve()
        for l in me["levels"]:
            eng.upsample(l)
            eng.initialize(l)
            d = depths[l]
            mid = d // 2
            if l == nst and nst > 1:
                stream_level(l, None, me["send_to"], True)
            else:
                eng.optimize_chains(l, max_iter[l], 3 if partner is None else (1 << dr))
                if partner is not None and d > 1:
                    fwd, bwd = (mid + 1, d), (0, mid)
                    _swap_pages(eng, l, fwd if dr == 0 else bwd, bwd if dr == 0 else fwd, partner)
    else:
        stream_level(me["levels"][0], me["recv_from"], me["send_to"], False)
    for t, w in sends:
        w.wait()
    if me["pair"] == 0 and nst > 1 and partner is not None:              # both level-1 owners end with the whole field
        d = depths[1]
        mid = d // 2
        fwd, bwd = (mid + 1, d), (0, mid)
        _swap_pages(eng, 1, fwd if dr == 0 else bwd, bwd if dr == 0 else fwd, partner)
    eng.sync()
    return plan


def optimize_video(morph, pyramid, params, device=0):
    """Morph::calculate_halfway_parametrization (morph.cu:150-168) for a video on 1 .. 8 GPUs, exact mode (the same
    arithmetic as one GPU; ranks 0 and 1 end with the bit-identical level-1 field).  2-3 ranks: one frame chain per rank.
    4+ ranks: direction x level pipeline (pipeline_plan).  Every rank holds the whole pyramid, built from the same frames."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    depths = [pyramid.info(l)["d"] for l in range(pyramid.num_levels)]
    if pipeline_plan(depths, world)["nstages"] < 2:
        return _optimize_video_two_chains(morph, pyramid, params, device)
    run_pipeline(MorphEngine(morph, pyramid, device), float(params.max_iter), float(params.max_iter_drop_factor), rank, world)
    return morph
