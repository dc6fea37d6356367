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
