"""Video optimizer in exact multi-GPU mode (dist.optimize_video): 2 ranks = one frame chain per rank, `v` halves swapped per
level; 4+ ranks = direction x level pipeline (frames handed from level to level over NCCL send / recv).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/video_dist_bench.py --width 1280 --height 720 --frames 120
With one process it runs the single-GPU schedule (both chains concurrently on two streams)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", dest="w", type=int, default=1280); ap.add_argument("--height", dest="h", type=int, default=720); ap.add_argument("--frames", dest="d", type=int, default=120)
    ap.add_argument("--max-iter", type=int, default=1000); ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import videomorphing_b200 as vm
    from videomorphing_b200 import dist as vd, synth
    rank, local, world = vd.env_world()
    torch.cuda.set_device(local)
    vd.init("nccl", device_id=local)
    v0, v1, flows, field = synth.video_pair(args.w, args.h, args.d, 4001, 4002, 8.0)
    prm = vm.Parameters(max_iter=args.max_iter)
    pyr = vm.Pyramid(local)
    pyr.build(v0, v1, flows, voxel_cap=1 << 62)
    for rep in range(args.reps):
        m = vm.Morph(prm, pyr)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        t = time.perf_counter()
        vd.optimize_video(m, pyr, prm, device=local)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        px = m.executed_pixel_iters
        px_all, t_max = vd.reduce_throughput(px, dt)
        # device time this rank spent inside sweep launches (CUDA events around every k_sweep): which stage bounds the pipeline
        busy = torch.tensor([m.sweep_time_ms()[0]], dtype=torch.float64, device="cuda")
        busy_all = [torch.zeros_like(busy) for _ in range(world)]
        if world > 1: dist.all_gather(busy_all, busy)
        else: busy_all = [busy]
        if rank == 0:
            vec = m.get_vectors()
            plan = vd.pipeline_plan([pyr.info(l)["d"] for l in range(pyr.num_levels)], world)
            sched = ("direction x level pipeline, %d stages: " % plan["nstages"] + "; ".join(
                "rank %d %s levels %s" % (r, "fwd" if e["dir"] == 0 else "bwd", e["levels"]) for r, e in sorted(plan["ranks"].items()))) \
                if plan["nstages"] > 1 else ("one frame chain per rank" if world > 1 else "both chains on one GPU")
            print(json.dumps({"workload": f"{args.w}x{args.h}x{args.d} video pair, exact mode over {world} GPU(s)", "schedule": sched, "rep": rep, "optimize_s": t_max,
                              "pixel_iters_all_ranks_incl_duplicate_mid_frames": px_all, "frames_per_s_optimize": args.d / t_max,
                              "sweep_busy_ms_per_rank": [round(float(b.item()), 1) for b in busy_all],
                              "checksum": float(np.abs(vec).sum())}), flush=True)
        m.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
