"""ctypes loader for libvmorph.so (the C-ABI CUDA library, include/vmorph.h).

Fails loudly when the library is missing: there is no CPU or PyTorch fallback for any compute entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VMORPH_LIB") or os.path.join(_HERE, "libvmorph.so")   # VMORPH_LIB: development builds (e.g. libvmorph_trace.so)


class VmParams(C.Structure):
    _fields_ = [("w_ui", C.c_float), ("w_tps", C.c_float), ("w_ssim", C.c_float), ("w_temp", C.c_float),
                ("ssim_clamp", C.c_float), ("eps", C.c_float), ("max_iter", C.c_int32), ("start_res", C.c_int32),
                ("max_iter_drop_factor", C.c_float), ("bcond", C.c_int32)]


class VmConp(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("z", C.c_int32), ("w", C.c_int32), ("weight", C.c_float)]


class VmConnect(C.Structure):
    _fields_ = [("li_track", C.c_int32), ("li_idx", C.c_int32), ("ri_track", C.c_int32), ("ri_idx", C.c_int32)]


class VmLevelInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth", C.c_int32), ("rowstride", C.c_int32),
                ("pagestride", C.c_int32), ("impmask_rowstride", C.c_int32), ("impmask_pagestride", C.c_int32),
                ("has_images", C.c_int32), ("factor_t", C.c_int32), ("factor_d", C.c_float), ("inv_wh", C.c_float)]


class VmTracks(C.Structure):
    _fields_ = [("n_left", C.c_int32), ("n_right", C.c_int32), ("n_groups", C.c_int32),
                ("left_len", C.POINTER(C.c_int32)), ("right_len", C.POINTER(C.c_int32)), ("group_len", C.POINTER(C.c_int32)),
                ("left", C.POINTER(VmConp)), ("right", C.POINTER(VmConp)), ("connects", C.POINTER(VmConnect))]


# every symbol include/vmorph.h declares (tests/test_cabi.py checks the list against the header)
EXPORTS = [
    "vm_last_error", "vm_device_count", "vm_version", "vm_params_default", "vm_params_parse_xml", "vm_params_write_xml", "vm_tracks_free",
    "vm_pyramid_create", "vm_pyramid_destroy", "vm_level_schedule", "vm_pyramid_alloc", "vm_pyramid_build",
    "vm_pyramid_build_frames", "vm_pyramid_build_finish", "vm_wavefront_plan", "vm_device_enable_peer", "vm_host_pin", "vm_host_unpin",
    "vm_pyramid_num_levels", "vm_pyramid_level_info", "vm_level_get", "vm_level_set", "vm_morph_create", "vm_morph_destroy",
    "vm_morph_set_tracks", "vm_morph_set_constraints", "vm_morph_run", "vm_morph_progress", "vm_morph_executed_pixel_iters",
    "vm_morph_sweep_ms", "vm_morph_attempted_updates", "vm_morph_sweep_busy_ms", "vm_morph_updates_log", "vm_morph_ms_log", "vm_morph_iters_log", "vm_level_cpu_solve", "vm_level_upsample", "vm_level_initialize", "vm_level_init_temp", "vm_level_upsample_frames", "vm_level_initialize_frames",
    "vm_level_optimize_frame", "vm_level_optimize", "vm_level_optimize_chains", "vm_morph_wavefront_prepare", "vm_level_enqueue_jobs", "vm_morph_collect", "vm_level_dev_ptr", "vm_level_mark_v_valid", "vm_dev_copy", "vm_level_energy", "vm_morph_get_vectors", "vm_morph_get_vectors_level", "vm_morph_extract", "vm_morph_render_frames", "vm_stencils_get",
    "vm_render_halfway_dev", "vm_render_halfway", "vm_render_sequence", "vm_qpath_optimize", "vm_qpath_optimize_frames", "vm_dev_alloc", "vm_dev_free", "vm_dev_upload",
    "vm_dev_download", "vm_stream_sync", "vm_kernel_launch_count", "vm_selftest_exact_arith", "vm_debug_sweep_phases",
]

_lib = None


def load():
    """Returns the loaded C-ABI library; raises if it has not been built (python -m videomorphing_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python videomorphing_b200/build.py` "
                           "(nvcc, sm_100a). There is no fallback path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    L.vm_last_error.restype = C.c_char_p
    L.vm_version.restype = C.c_char_p
    L.vm_kernel_launch_count.restype = C.c_uint64
    L.vm_params_default.argtypes = [C.POINTER(VmParams)]
    L.vm_params_parse_xml.argtypes = [C.c_char_p, C.POINTER(VmParams), C.POINTER(VmTracks)]
    L.vm_params_write_xml.argtypes = [C.c_char_p, C.POINTER(VmParams), C.POINTER(VmTracks), i32]
    L.vm_tracks_free.argtypes = [C.POINTER(VmTracks)]
    L.vm_pyramid_create.argtypes = [i32, C.POINTER(vp)]
    L.vm_pyramid_destroy.argtypes = [vp]
    L.vm_level_schedule.argtypes = [i32, i32, i32, i32, i64, i32, C.POINTER(C.c_int32), C.POINTER(f32)]
    L.vm_pyramid_alloc.argtypes = [vp, i32, i32, i32, i32, i64]
    L.vm_pyramid_build.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp]
    L.vm_pyramid_build_frames.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i32, i32, vp]
    L.vm_pyramid_build_finish.argtypes = [vp, vp]
    L.vm_device_enable_peer.argtypes = [i32, i32]
    L.vm_host_pin.argtypes = [vp, C.c_size_t]
    L.vm_host_unpin.argtypes = [vp]
    L.vm_wavefront_plan.argtypes = [i32, C.POINTER(C.c_int32), C.POINTER(f32), i32, C.POINTER(C.c_int32)]
    L.vm_pyramid_num_levels.argtypes = [vp]
    L.vm_pyramid_level_info.argtypes = [vp, i32, C.POINTER(VmLevelInfo)]
    L.vm_level_get.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.vm_level_set.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.vm_morph_create.argtypes = [C.POINTER(VmParams), vp, vp, C.POINTER(vp)]
    L.vm_morph_destroy.argtypes = [vp]
    L.vm_morph_set_tracks.argtypes = [vp, i32, vp, vp, i32, vp, vp, i32, vp, vp]
    L.vm_morph_set_constraints.argtypes = [vp, i32, vp, vp]
    L.vm_morph_run.argtypes = [vp, vp]
    L.vm_morph_progress.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(f32)]
    L.vm_morph_executed_pixel_iters.argtypes = [vp]
    L.vm_morph_executed_pixel_iters.restype = C.c_double
    L.vm_morph_sweep_ms.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.vm_morph_sweep_ms.restype = C.c_double
    L.vm_morph_attempted_updates.argtypes = [vp]
    L.vm_morph_attempted_updates.restype = C.c_double
    L.vm_morph_sweep_busy_ms.argtypes = [vp]
    L.vm_morph_sweep_busy_ms.restype = C.c_double
    L.vm_morph_updates_log.argtypes = [vp, i32, vp]
    L.vm_morph_ms_log.argtypes = [vp, i32, vp]
    L.vm_morph_iters_log.argtypes = [vp, i32, vp]
    L.vm_level_cpu_solve.argtypes = [vp, vp]
    L.vm_level_upsample.argtypes = [vp, i32, vp]
    L.vm_level_initialize.argtypes = [vp, i32, vp]
    L.vm_level_init_temp.argtypes = [vp, i32, i32, i32, vp]
    L.vm_level_optimize_frame.argtypes = [vp, i32, i32, i32, f32, C.POINTER(i32), vp]
    L.vm_level_optimize.argtypes = [vp, i32, f32, vp]
    L.vm_level_upsample_frames.argtypes = [vp, i32, i32, i32, vp]
    L.vm_level_initialize_frames.argtypes = [vp, i32, i32, i32, vp]
    L.vm_level_optimize_chains.argtypes = [vp, i32, f32, i32, vp]
    L.vm_morph_wavefront_prepare.argtypes = [vp, vp]
    L.vm_level_enqueue_jobs.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.vm_morph_collect.argtypes = [vp, vp]
    L.vm_level_dev_ptr.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.vm_level_mark_v_valid.argtypes = [vp, i32]
    L.vm_dev_copy.argtypes = [i32, vp, vp, C.c_size_t, vp]
    L.vm_level_energy.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.vm_morph_get_vectors.argtypes = [vp, vp, vp]
    L.vm_morph_get_vectors_level.argtypes = [vp, i32, vp, vp]
    L.vm_morph_extract.argtypes = [vp, i32, vp]
    L.vm_morph_render_frames.argtypes = [vp, i32, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp]
    L.vm_stencils_get.argtypes = [vp, vp, vp]
    L.vm_render_halfway_dev.argtypes = [vp, i32, i32, i32, i32, f32, f32, i32, vp, vp, vp, vp, vp]
    L.vm_render_halfway.argtypes = [i32, vp, i32, i32, i32, f32, f32, i32, vp, vp, vp, vp, vp]
    L.vm_render_sequence.argtypes = [i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    L.vm_qpath_optimize.argtypes = [i32, vp, vp, i32, i32, i32, f32, C.POINTER(i32), vp]
    L.vm_qpath_optimize_frames.argtypes = [i32, vp, vp, i32, i32, i32, i32, f32, vp, vp]
    L.vm_dev_alloc.argtypes = [i32, C.c_size_t, C.POINTER(vp)]
    L.vm_dev_free.argtypes = [i32, vp]
    L.vm_dev_upload.argtypes = [i32, vp, vp, C.c_size_t, vp]
    L.vm_dev_download.argtypes = [i32, vp, vp, C.c_size_t, vp]
    L.vm_stream_sync.argtypes = [i32, vp]
    L.vm_debug_sweep_phases.argtypes = [i32, vp, i32]
    L.vm_selftest_exact_arith.argtypes = [i32, C.c_uint64, C.POINTER(C.c_uint64)]
    _lib = L
    return L


class VmError(RuntimeError):
    pass


def check(rc):
    if rc < 0:
        raise VmError(f"libvmorph status {rc}: {load().vm_last_error().decode()}")
    return rc
