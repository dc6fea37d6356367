"""tools/vmorph_headless.cpp: the headless C++ driver over the C ABI (plain g++, no CUDA headers)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe():
    from videomorphing_b200 import build as vb
    vb.build()
    return vb.build_headless()


def _ppm(path, rgb):
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(rgb, np.uint8).tobytes())


def test_builds_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _exe()
    r = subprocess.run([exe, "--version"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "vmorph" in r.stdout
    assert subprocess.run([exe], stderr=subprocess.PIPE).returncode == 2          # usage
    import torch
    if not torch.cuda.is_available():
        from videomorphing_b200 import synth
        a, b, _ = synth.image_pair(64, 48, 1, 2, 2.0)
        _ppm(tmp_path / "a.ppm", a[0]); _ppm(tmp_path / "b.ppm", b[0])
        r = subprocess.run([exe, "--img0", str(tmp_path / "a.ppm"), "--img1", str(tmp_path / "b.ppm")], stderr=subprocess.PIPE, text=True)
        assert r.returncode == 3 and "no CUDA device" in r.stderr                 # no CPU fallback


@pytest.mark.gpu
def test_headless_matches_the_python_host(tmp_path, oracle_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import videomorphing_b200 as vm
    from videomorphing_b200 import synth
    exe = _exe()
    w, h = 96, 64
    a, b, _ = synth.image_pair(w, h, 11, 12, 3.0)
    _ppm(tmp_path / "a.ppm", a[0]); _ppm(tmp_path / "b.ppm", b[0])
    r = subprocess.run([exe, "--img0", str(tmp_path / "a.ppm"), "--img1", str(tmp_path / "b.ppm"), "--frames", "3", "--max-iter", "24",
                        "--out", str(tmp_path / "m_%03d.ppm"), "--vectors", str(tmp_path / "v.bin")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["frames"] == 3 and info["pixel_iters"] > 0
    v = np.fromfile(tmp_path / "v.bin", np.float32).reshape(h, w, 2)
    o = oracle_lib.Oracle(dict(max_iter=24))
    o.build(a, b); o.run()
    np.testing.assert_array_equal(v, o.extract_vectors()[0])                       # same bits as the oracle
    ex = int(max(w, h) * 0.1)
    e0, e1 = synth.extended_rgba(a[0], ex), synth.extended_rgba(b[0], ex)
    for k in range(3):
        fa = float(synth.smoothstep(k / 2))
        with open(tmp_path / ("m_%03d.ppm" % k), "rb") as f:
            assert f.readline() == b"P6\n"; f.readline(); f.readline()
            img = np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)
        ref = oracle_lib.render_halfway(w, h, ex, fa, fa, 1, e0, e1, v)[:, :w]
        np.testing.assert_array_equal(img, ref)


# ------------------------------------------------------------------ tools/vmorph_video.cpp: the multi-GPU video host in C++
def _video_exe():
    from videomorphing_b200 import build as vb
    vb.build()
    return vb.build_video_host()


def _write_video(tmp_path, v0, v1, flows):
    v0.astype(np.uint8).tofile(tmp_path / "v0.rgb"); v1.astype(np.uint8).tofile(tmp_path / "v1.rgb")
    for name, f in zip(("f0", "f1", "b0", "b1"), flows):
        np.ascontiguousarray(f, np.float32).tofile(tmp_path / (name + ".bin"))
    d, h, w, _ = v0.shape
    return ["--size", str(w), str(h), str(d), "--video0", str(tmp_path / "v0.rgb"), "--video1", str(tmp_path / "v1.rgb"),
            "--flows"] + [str(tmp_path / (n + ".bin")) for n in ("f0", "f1", "b0", "b1")]


def test_video_host_builds_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _video_exe()
    r = subprocess.run([exe, "--version"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "vmorph" in r.stdout
    assert subprocess.run([exe], stderr=subprocess.PIPE).returncode == 2          # usage
    import torch
    if not torch.cuda.is_available():
        from videomorphing_b200 import synth
        v0, v1, flows, _ = synth.video_pair(48, 36, 3, 1, 2, 2.0)
        args = _write_video(tmp_path, v0, v1, flows)
        r = subprocess.run([exe] + args + ["--devices", "0,0"], stderr=subprocess.PIPE, text=True)
        assert r.returncode == 3 and "no CUDA device" in r.stderr                 # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0", "0,0", "0,0,0", "0,0,0,0", "0,0,0,0,0,0,0,0", "0,1", "0,1,2,3"])
def test_video_host_exact_mode_equals_one_gpu_run(tmp_path, devices):
    """tools/vmorph_video.cpp (threads + peer copies over the C ABI: frame-sharded Pyramid::build, the wavefront split by
    vm_wavefront_plan) gives the bits of vm_morph_run -- with the ranks sharing GPU 0 (the one-GPU lease) and on 2 / 4 GPUs."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    need = max(int(x) for x in devices.split(",")) + 1
    if torch.cuda.device_count() < need:
        pytest.skip(f"needs {need} GPUs")
    import videomorphing_b200 as vm
    from videomorphing_b200 import synth
    exe = _video_exe()
    v0, v1, flows, _ = synth.video_pair(96, 64, 9, 41, 42, 3.0)
    args = _write_video(tmp_path, v0, v1, flows)
    r = subprocess.run([exe] + args + ["--devices", devices, "--start-res", "4", "--max-iter", "24", "--voxel-cap", str(1 << 62), "--vectors", str(tmp_path / "v.bin")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["ranks"] == len(devices.split(",")) and info["frames"] == 9
    got = np.fromfile(tmp_path / "v.bin", np.float32).reshape(9, 64, 96, 2)
    prm = vm.Parameters(max_iter=24, start_res=4)
    pyr = vm.Pyramid(0); pyr.build(v0, v1, flows, start_res=4, voxel_cap=1 << 62)
    m = vm.Morph(prm, pyr); m.run()
    np.testing.assert_array_equal(got, m.get_vectors())
