"""tools/vmorph_video (the C++ multi-GPU host) on the 720p x 120 video of bench.py (no UI tracks: the host reads them from a
settings.xml only): writes the raw inputs to a scratch directory, runs the host on 1 and on N GPUs, checks that the vector
fields are bit-identical.    python tools/video_host_bench.py [--gpus 2] [--frames 120]"""
import argparse, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2); ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--w", type=int, default=1280); ap.add_argument("--h", type=int, default=720)
    args = ap.parse_args()
    from videomorphing_b200 import build as vb, synth
    vb.build(); exe = vb.build_video_host()
    v0, v1, flows, _ = synth.video_pair(args.w, args.h, args.frames, 4001, 4002, 8.0)
    tmp = tempfile.mkdtemp(prefix="vmorph_video_")
    v0.astype(np.uint8).tofile(os.path.join(tmp, "v0.rgb")); v1.astype(np.uint8).tofile(os.path.join(tmp, "v1.rgb"))
    for n, f in zip(("f0", "f1", "b0", "b1"), flows):
        np.ascontiguousarray(f, np.float32).tofile(os.path.join(tmp, n + ".bin"))
    base = [exe, "--size", str(args.w), str(args.h), str(args.frames), "--video0", os.path.join(tmp, "v0.rgb"), "--video1", os.path.join(tmp, "v1.rgb"),
            "--flows"] + [os.path.join(tmp, n + ".bin") for n in ("f0", "f1", "b0", "b1")] + ["--voxel-cap", str(1 << 62), "--repeat", "3"]
    outs = {}
    for n in sorted({1, args.gpus}):
        devs = ",".join(str(i) for i in range(n))
        t = time.perf_counter()
        r = subprocess.run(base + ["--devices", devs, "--vectors", os.path.join(tmp, f"v{n}.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        info = json.loads(r.stdout.strip().splitlines()[-1]); info["wall_s_incl_file_reads"] = time.perf_counter() - t
        outs[n] = np.fromfile(os.path.join(tmp, f"v{n}.bin"), np.float32)
        info["checksum_sum_abs_v"] = float(np.abs(outs[n]).sum(dtype=np.float64))
        print(json.dumps(info), flush=True)
    if len(outs) == 2:
        a, b = outs.values()
        print(json.dumps({"bit_identical_1_vs_%d_gpus" % args.gpus: bool(np.array_equal(a, b))}))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
