"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/vmorph.h declares,
its host-side logic (level schedule, stencil tables, argument validation) matches the oracle, and compute entry points
fail loudly when no CUDA device is present (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vm():
    from videomorphing_b200 import build
    build.build()
    import videomorphing_b200 as vm
    return vm


def test_exports_match_header(vm):
    hdr = open(os.path.join(ROOT, "include", "vmorph.h")).read()
    declared = set(re.findall(r"\b(vm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vm_conp", "vm_connect", "vm_params", "vm_tracks", "vm_level_info", "vm_pyramid", "vm_morph"}
    L = vm._lib.load()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/vmorph.h but not exported"
    assert set(vm._lib.EXPORTS) == declared


def test_struct_layouts(vm):
    assert C.sizeof(vm._lib.VmParams) == 40 and C.sizeof(vm._lib.VmConp) == 20 and C.sizeof(vm._lib.VmLevelInfo) == 44


def test_defaults(vm):
    p = vm.Parameters()          # UI/MdiEditor.cpp:131-140
    assert (p.w_ssim, p.ssim_clamp, p.w_ui, p.w_temp, p.max_iter, p.max_iter_drop_factor, p.start_res, p.bcond) == \
        (100.0, 0.0, 100000.0, 10.0, 1000, 2.0, 8, 0)
    assert abs(p.w_tps - 0.05) < 1e-9 and abs(p.eps - 0.01) < 1e-9


def test_schedule_bit_exact_vs_oracle(vm, oracle_lib):
    rng = np.random.Generator(np.random.PCG64(3))
    cases = [(256, 256, 1), (512, 512, 1), (1920, 1080, 1), (1280, 720, 120), (3840, 2160, 240), (600, 338, 100), (64, 64, 1)]
    cases += [(int(rng.integers(24, 2000)), int(rng.integers(24, 1200)), int(rng.integers(1, 130))) for _ in range(60)]
    for (w, h, d) in cases:
        for cap in (14000000, 10 ** 12):
            a = vm.level_schedule(w, h, d, 8, cap)
            b = oracle_lib.schedule(w, h, d, 8, cap)
            assert [(e["w"], e["h"], e["d"], e["factor_d"]) for e in a] == [(e["w"], e["h"], e["d"], e["factor_d"]) for e in b], (w, h, d, cap)


def test_stencils_bit_exact_vs_oracle(vm, oracle_lib):
    # product tables are derived from the dense-operator formulation; the oracle transcribes stencils.cpp literally
    for a, b in zip(vm.stencils(), oracle_lib.stencils()):
        np.testing.assert_array_equal(a, b)


def test_argument_validation(vm):
    L = vm._lib.load()
    assert L.vm_level_schedule(0, 10, 1, 8, 14000000, 64, None, None) == -1
    assert b"bad schedule" in L.vm_last_error()
    assert L.vm_params_default(None) == -1
    assert L.vm_render_halfway_dev(None, 0, 0, 0, 0, 0.5, 0.5, 1, None, None, None, None, None) == -1
    # round-2 entry points: the multi-GPU plan, the frame-sharded build, pinning
    own = (C.c_int32 * 8)()
    assert L.vm_wavefront_plan(2, None, None, 2, own) == -1 and b"wavefront_plan" in L.vm_last_error()
    whd = (C.c_int32 * 12)(64, 48, 9, 64, 48, 9, 32, 24, 9, 16, 12, 5)
    mi = (C.c_float * 4)(0, 50, 100, 0)
    assert L.vm_wavefront_plan(4, whd, mi, 0, own) == -1
    assert L.vm_wavefront_plan(4, whd, mi, 2, own) == 2 and list(own) == [-1, -1, 0, 1, 0, 1, -1, -1]     # levels 2, 1: forward on rank 0, backward on rank 1
    assert L.vm_pyramid_build_frames(None, None, None, None, None, None, None, 64, 48, 9, 4, 14000000, 0, 9, None) == -1
    assert L.vm_pyramid_build_finish(None, None) < 0
    assert L.vm_host_pin(None, 16) == -1
    assert L.vm_host_unpin(None) == 0


def test_no_cpu_fallback(vm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = vm._lib.load()
    assert L.vm_device_count() == 0
    with pytest.raises(vm._lib.VmError) as e:
        vm.Pyramid(0)
    assert "no CPU fallback" in str(e.value)
    out = np.zeros((4, 4, 3), np.uint8)
    ext = np.zeros((4, 4, 4), np.uint8)
    v = np.zeros((4, 4, 2), np.float32)
    with pytest.raises(vm._lib.VmError):
        vm.render_halfway_image(4, 4, 0, 0.5, 0.5, 1, ext, ext, v)


def test_product_does_not_reference_oracle():
    # the product package must never import / link / call anything under oracle/
    pkg = os.path.join(ROOT, "videomorphing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "vmo.h" not in txt, f
