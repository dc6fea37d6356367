"""Full-size checks at BASELINE.json sizes through size-independent properties (the oracle is too slow there):
determinism, incremental == direct state, energy monotonicity across the run, recovery of the synthetic warp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vm():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from videomorphing_b200 import build
    build.build()
    import videomorphing_b200 as vm
    return vm


def _levels_from_oracle(vm, oracle_lib, rgb0, rgb1):
    o = oracle_lib.Oracle()
    n = o.build(rgb0, rgb1)
    d, h, w, _ = rgb0.shape
    pyr = vm.Pyramid(0)
    assert pyr.alloc(w, h, d) == n
    for l in range(1, n - 1):
        pyr.set(l, "img0", o.get(l, "img0")); pyr.set(l, "img1", o.get(l, "img1"))
    return pyr, n


def test_cfg2_properties(vm, oracle_lib):
    # BASELINE.json configs[1]: 512x512 pair, 20 UI point constraints, full pyramid
    from videomorphing_b200 import synth
    w, h, d, s1, s2, amp = synth.CONFIGS["cfg2"]
    rgb0, rgb1, field = synth.image_pair(w, h, s1, s2, amp)
    cons = synth.point_pairs(20, w, h, 2003, field)
    pyr, n = _levels_from_oracle(vm, oracle_lib, rgb0, rgb1)
    runs = []
    for _ in range(2):
        m = vm.Morph(vm.Parameters(), pyr)
        m.set_constraints(*cons)
        m.run()
        runs.append((m.get_vectors(), m.iters_log().copy(), m.energy(1)[0]))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])            # deterministic (no float atomics anywhere)
    np.testing.assert_array_equal(runs[0][1], runs[1][1])
    log = runs[0][1]
    assert list(log[:, 0]) == list(range(n - 2, 0, -1))
    max_it = 1000.0
    for l, _, it in log:
        assert 1 <= it <= int(np.ceil(max_it))
        max_it /= 2
    # incremental state == direct re-initialisation from the final v (SURVEY section 4 cross-check 3)
    inc = {k: pyr.get(1, k) for k in ("mean", "var", "cross", "value", "tps_b")}
    e_inc = m.energy(1)[0]
    m.initialize_level(1)
    for k, a in inc.items():
        np.testing.assert_allclose(a, pyr.get(1, k), rtol=3e-4, atol=0.5 if k in ("var", "cross") else 5e-3, err_msg=k)
    assert abs(m.energy(1)[0] - e_inc) <= 1e-3 * abs(e_inc)
    # the optimizer recovers the synthetic warp: halfway vector ~ warp / 2
    err = np.abs(runs[0][0][0] - field / 2)
    assert err.mean() < 0.5 and np.median(err) < 0.25


def test_optimisation_lowers_energy_at_every_level(vm, oracle_lib):
    from videomorphing_b200 import synth
    rgb0, rgb1, _ = synth.image_pair(320, 200, 81, 82, 8.0)
    pyr, n = _levels_from_oracle(vm, oracle_lib, rgb0, rgb1)
    m = vm.Morph(vm.Parameters(max_iter=200), pyr)
    m.cpu_optimize_level()
    mi = 200.0
    for l in range(n - 2, 0, -1):
        m.upsample(l); m.initialize_level(l)
        e0 = m.energy(l)[0]
        m.optimize_frame(l, 0, False, mi)
        e1 = m.energy(l)[0]
        assert e1 <= e0 * (1 + 1e-6), (l, e0, e1)
        mi /= 2


def _two_gpu_worker(rank, port, q, world=2):
    import os
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VMORPH_DIST_TIMEOUT_S="120")
    import torch
    import torch.distributed as dist
    import videomorphing_b200 as vm
    from videomorphing_b200 import dist as vd, synth
    torch.cuda.set_device(rank)
    vd.init("nccl", device_id=rank)
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(rank); vd.build_pyramid(pyr, v0, v1, flows, start_res=4, device=rank)   # ragged frame blocks: one broadcast per block
    m = vm.Morph(prm, pyr)
    vd.optimize_video(m, pyr, prm, device=rank)
    vec = m.get_vectors()                                  # every rank ends with the whole level-1 field
    # equal frame blocks (8 frames): the in-place NCCL all-gather of dist.build_pyramid == Pyramid::build on this GPU
    fl8 = tuple(f[:8] for f in flows)
    pa = vm.Pyramid(rank); vd.build_pyramid(pa, v0[:8], v1[:8], fl8, start_res=4, device=rank)
    pb = vm.Pyramid(rank); pb.build(v0[:8], v1[:8], fl8, start_res=4)
    assert _pyramid_digest(pa) == _pyramid_digest(pb)
    dist.barrier()
    q.put((rank, vec, m.iters_log().copy()))
    dist.destroy_process_group()


def test_two_gpu_chain_split_is_bit_identical_to_one_gpu(vm):
    # exact multi-GPU mode: the forward half of the wavefront on GPU 0, the backward half on GPU 1 (dist.run_wavefront over NCCL)
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from videomorphing_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (v, it)) for r, v, it in (q.get(timeout=240) for _ in range(2)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); pyr.build(v0, v1, flows, start_res=4)
    m = vm.Morph(prm, pyr); m.run()
    ref = m.get_vectors()
    np.testing.assert_array_equal(res[0][0], ref)
    np.testing.assert_array_equal(res[1][0], ref)
    # each rank logged the middle frame + its own chain of every level; together they cover the single-GPU log
    one = {(int(l), int(f)): int(i) for l, f, i in m.iters_log()}
    both = {}
    for r in (0, 1):
        for l, f, i in res[r][1]:
            assert one[(int(l), int(f))] == int(i)
            both[(int(l), int(f))] = int(i)
    assert both == one


def test_frame_by_frame_level_ops_match_the_whole_level_run(vm):
    """vm_level_upsample_frames / vm_level_initialize_frames / vm_level_init_temp / vm_level_optimize_frame (the wavefront's
    per-frame operators) called one frame at a time in chain order give the bits of vm_morph_run."""
    from videomorphing_b200 import dist as vd, synth
    v0, v1, flows, field = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); n = pyr.build(v0, v1, flows, start_res=4)
    # UI point pairs on several frames (level-0 pixel units, p.z = frame): exercises the per-frame UI splat
    rng = np.random.Generator(np.random.PCG64(43))
    k = 12
    lp = np.stack([rng.integers(8, 88, k), rng.integers(8, 56, k), rng.integers(0, 9, k), np.ones(k, np.int64)], 1).astype(np.int32)
    rp = lp.copy(); rp[:, 0] += rng.integers(-3, 4, k).astype(np.int32); rp[:, 1] += rng.integers(-3, 4, k).astype(np.int32)
    cons = (lp, np.ones(k, np.float32), rp, np.ones(k, np.float32))
    m = vm.Morph(prm, pyr)
    m.set_constraints(*cons)
    m.run()
    ref, ref_log = m.get_vectors(), {(int(l), int(f)): int(i) for l, f, i in m.iters_log()}
    depths = [pyr.info(l)["d"] for l in range(n)]
    K = 1
    while K < n - 2 and depths[K] == depths[K + 1]:
        K += 1
    assert K >= 3                                         # levels 1 .. K-1 take the frame-by-frame path (level K is prolonged as a whole)
    m2 = vm.Morph(prm, pyr)
    m2.set_constraints(*cons)
    m2.cpu_optimize_level()
    mi = np.float32(24)
    chain = lambda d, dr: [d // 2] + (list(range(d // 2 + 1, d)) if dr == 0 else list(range(d // 2 - 1, -1, -1)))
    for l in range(n - 2, 0, -1):
        if l >= K:
            m2.upsample(l); m2.initialize_level(l); m2.optimize_chains(l, float(mi), 3)
        else:
            mid = depths[l] // 2
            for dr in (0, 1):
                for i in chain(depths[l], dr):
                    if dr == 1 and i == mid:
                        continue
                    m2.upsample_frames(l, i); m2.initialize_frames(l, i)
                    if i != mid:
                        m2.initialize_temp(l, i, -1 if dr == 0 else 1)
                    m2.optimize_frame(l, i, i != mid, float(mi))
        mi = np.float32(mi / np.float32(2))
    np.testing.assert_array_equal(m2.get_vectors(), ref)
    assert {(int(l), int(f)): int(i) for l, f, i in m2.iters_log()} == ref_log


def test_four_gpu_level_pipeline_is_bit_identical_to_one_gpu(vm):
    # exact multi-GPU mode on 4 GPUs: two level groups per direction (dist.wavefront_plan), frames handed on over NCCL
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 4:
        pytest.skip("needs four CUDA devices")
    from videomorphing_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_gpu_worker, args=(r, port, q, 4)) for r in range(4)]
    for p in procs:
        p.start()
    res = dict((r, (v, it)) for r, v, it in (q.get(timeout=300) for _ in range(4)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); pyr.build(v0, v1, flows, start_res=4)
    m = vm.Morph(prm, pyr); m.run()
    ref = m.get_vectors()
    for r in range(4):
        np.testing.assert_array_equal(res[r][0], ref)      # every rank ends with the whole field
    one = {(int(l), int(f)): int(i) for l, f, i in m.iters_log()}
    allr = {}
    for r in range(4):
        for l, f, i in res[r][1]:
            assert one[(int(l), int(f))] == int(i)
            allr[(int(l), int(f))] = int(i)
    assert allr == one                                     # together the four ranks ran every (level, frame) of the one-GPU log


def _pyramid_digest(pyr):
    import hashlib
    out = {}
    for l in range(1, pyr.num_levels - 1):
        for nm in ("img0", "img1", "f0", "f1", "b0", "b1"):
            out[(l, nm)] = hashlib.sha1(pyr.get(l, nm).tobytes()).hexdigest()
    return out


def _one_gpu_pipeline_worker(rank, port, q, world):
    """Rank of the exact multi-GPU schedule with EVERY rank on cuda:0 (gloo carries the hand-offs through pinned host
    memory): the real kernels and the real run_pipeline / chain-split code on a one-GPU lease."""
    import os
    os.environ.update(RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VMORPH_DIST_TIMEOUT_S="240")
    import torch
    import torch.distributed as dist
    import videomorphing_b200 as vm
    from videomorphing_b200 import dist as vd, synth
    torch.cuda.set_device(0)
    vd.init("gloo")
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); vd.build_pyramid(pyr, v0, v1, flows, start_res=4, device=0)      # each rank builds its frame block
    digest = _pyramid_digest(pyr)
    m = vm.Morph(prm, pyr)
    vd.optimize_video(m, pyr, prm, device=0)
    vec = m.get_vectors()                                  # every rank ends with the whole level-1 field
    dist.barrier()
    q.put((rank, vec, m.iters_log().copy(), digest))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_exact_multi_rank_schedules_on_one_gpu_are_bit_identical(vm, world):
    """dist.optimize_video (the wavefront split by direction and level group, dist.run_wavefront) at world 2, 3, 4 and 8 with
    all ranks sharing ONE GPU: every rank ends with the bits of vm_morph_run, and the union of the ranks' iteration logs is
    the one-GPU log (the middle frames are optimised by both directions' owners and must agree).  (The same schedules on
    2 / 4 real GPUs over NCCL: test_two_gpu_chain_split_*, test_four_gpu_level_pipeline_*.)"""
    import socket
    import torch.multiprocessing as mp
    from videomorphing_b200 import dist as vd, synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_one_gpu_pipeline_worker, args=(r, port, q, world)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict((r, (v, it, dg)) for r, v, it, dg in (q.get(timeout=600) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); n = pyr.build(v0, v1, flows, start_res=4)
    if world == 8:                                         # 4 equal-depth levels: each of the 8 ranks owns one (level, direction) chain
        eng = vd.MorphEngine(vm.Morph(prm, pyr), pyr, 0, prm)
        assert sorted(vd.wavefront_plan(eng.depths, eng.dims, eng.max_iters, world)["groups"]) == list(range(8))
    digest = _pyramid_digest(pyr)
    m = vm.Morph(prm, pyr); m.run()
    ref = m.get_vectors()
    for r in range(world):
        assert res[r][2] == digest                         # dist.build_pyramid (frame blocks + exchange) == Pyramid::build on one GPU
        np.testing.assert_array_equal(res[r][0], ref)      # every rank ends with the whole field
    one = {(int(l), int(f)): int(i) for l, f, i in m.iters_log()}
    allr = {}
    for r in range(world):
        for l, f, i in res[r][1]:
            assert one[(int(l), int(f))] == int(i)
            allr[(int(l), int(f))] = int(i)
    assert allr == one
