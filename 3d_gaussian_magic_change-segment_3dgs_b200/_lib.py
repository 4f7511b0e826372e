"""ctypes binding of libgsr.so (include/gsr.h). Fails loudly when the CUDA library is missing: there is no
CPU or PyTorch fallback for any entry point."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgsr.so")

GSR_STAGE_COUNT = 16
c_float_p = ctypes.POINTER(ctypes.c_float)


class GsrView(ctypes.Structure):
    _fields_ = [
        ("image_width", ctypes.c_int32), ("image_height", ctypes.c_int32),
        ("tanfovx", ctypes.c_float), ("tanfovy", ctypes.c_float), ("scale_modifier", ctypes.c_float),
        ("sh_degree", ctypes.c_int32), ("sh_coeffs", ctypes.c_int32), ("num_class", ctypes.c_int32),
        ("prefiltered", ctypes.c_int32), ("debug", ctypes.c_int32),
        ("bg", ctypes.c_void_p), ("viewmatrix", ctypes.c_void_p), ("projmatrix", ctypes.c_void_p), ("campos", ctypes.c_void_p),
    ]


class GsrGaussians(ctypes.Structure):
    pass


GsrGaussians._fields_ = [
        ("P", ctypes.c_int32),
        ("means3D", ctypes.c_void_p), ("shs", ctypes.c_void_p), ("colors_precomp", ctypes.c_void_p), ("segments", ctypes.c_void_p),
        ("opacities", ctypes.c_void_p), ("scales", ctypes.c_void_p), ("rotations", ctypes.c_void_p), ("cov3D_precomp", ctypes.c_void_p),
        ("shs_rest", ctypes.c_void_p), ("raw_params", ctypes.c_int32),
        ("subset", ctypes.c_void_p), ("subset_count", ctypes.c_int32),
        ("parts", ctypes.POINTER(GsrGaussians)), ("num_parts", ctypes.c_int32),
]
GSR_MAX_PARTS = 16


class GsrMicrobench(ctypes.Structure):
    _fields_ = [("ffma_tflops", ctypes.c_float), ("ffma2_tflops", ctypes.c_float), ("ex2_gops", ctypes.c_float), ("shfl_gops", ctypes.c_float),
                ("red_gops", ctypes.c_float), ("sm_clock_mhz_nominal", ctypes.c_float), ("sm_count", ctypes.c_int32)]


class GsrOutputs(ctypes.Structure):
    _fields_ = [("color", ctypes.c_void_p), ("segment", ctypes.c_void_p), ("depth", ctypes.c_void_p), ("alpha", ctypes.c_void_p),
                ("radii", ctypes.c_void_p)]


class GsrState(ctypes.Structure):
    _fields_ = [("geom", ctypes.c_void_p), ("binning", ctypes.c_void_p), ("img", ctypes.c_void_p), ("num_rendered", ctypes.c_int32),
                ("num_visible", ctypes.c_int32)]


class GsrPixelGrads(ctypes.Structure):
    _fields_ = [("dL_dcolor", ctypes.c_void_p), ("dL_dsegment", ctypes.c_void_p), ("dL_ddepth", ctypes.c_void_p), ("dL_dalpha", ctypes.c_void_p)]


class GsrParamGrads(ctypes.Structure):
    _fields_ = [("dL_dmeans3D", ctypes.c_void_p), ("dL_dmeans2D", ctypes.c_void_p), ("dL_dsh", ctypes.c_void_p), ("dL_dcolors", ctypes.c_void_p),
                ("dL_dsegments", ctypes.c_void_p), ("dL_dopacity", ctypes.c_void_p), ("dL_dscales", ctypes.c_void_p),
                ("dL_drotations", ctypes.c_void_p), ("dL_dcov3D", ctypes.c_void_p), ("accumulate", ctypes.c_int32),
                ("dL_dsh_rest", ctypes.c_void_p)]


class GsrStateExport(ctypes.Structure):
    _fields_ = [("depths", ctypes.c_void_p), ("means2D", ctypes.c_void_p), ("conic_opacity", ctypes.c_void_p), ("rgb", ctypes.c_void_p),
                ("clamped", ctypes.c_void_p), ("tiles_touched", ctypes.c_void_p), ("point_keys", ctypes.c_void_p), ("point_list", ctypes.c_void_p),
                ("ranges", ctypes.c_void_p), ("n_contrib", ctypes.c_void_p)]


class GsrAdamGroup(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint64), ("count", ctypes.c_uint64), ("lr", ctypes.c_float), ("step", ctypes.c_int32)]


ALLOC_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t)

# every symbol include/gsr.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = ["gsr_abi_version", "gsr_last_error", "gsr_forward", "gsr_backward_scratch_bytes", "gsr_backward_scratch_bytes_n", "gsr_backward", "gsr_mark_visible",
           "gsr_knn_workspace_bytes", "gsr_knn_dist2", "gsr_export_state", "gsr_launch_count", "gsr_set_profiling", "gsr_get_stage_times",
           "gsr_backward_packets", "gsr_gather_packets", "gsr_gather_packets_v", "gsr_packet_index_words", "gsr_peer_alloc", "gsr_peer_open", "gsr_peer_close",
           "gsr_peer_free", "gsr_peer_copy", "gsr_adam_step", "gsr_select_rows", "gsr_image_loss", "gsr_image_loss_scratch_bytes", "gsr_last_num_visible", "gsr_microbench", "gsr_count_work", "gsr_depth_loss", "gsr_depth_loss_scratch_bytes"]
GSR_ABI_VERSION = 4  # include/gsr.h
GSR_PACKET_WORDS = 16
GSR_PEER_HANDLE_BYTES = 64
GSR_MAX_GATHER_VIEWS = 64

_lib = None


class GsrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libgsr.so is not built (%s). Run `python __graft_entry__.py` or the package's build.py; "
            "this package has no CPU/PyTorch fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.gsr_abi_version.restype = ctypes.c_int
    L.gsr_last_error.restype = ctypes.c_char_p
    L.gsr_forward.restype = ctypes.c_int
    L.gsr_forward.argtypes = [ctypes.POINTER(GsrView), ctypes.POINTER(GsrGaussians), ctypes.POINTER(GsrOutputs), ALLOC_FN, ctypes.c_void_p,
                              ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p]
    L.gsr_backward_scratch_bytes.restype = ctypes.c_size_t
    L.gsr_backward_scratch_bytes.argtypes = [ctypes.c_int32]
    L.gsr_backward_scratch_bytes_n.restype = ctypes.c_size_t
    L.gsr_backward_scratch_bytes_n.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.gsr_backward.restype = ctypes.c_int
    L.gsr_backward.argtypes = [ctypes.POINTER(GsrView), ctypes.POINTER(GsrGaussians), ctypes.c_void_p, ctypes.POINTER(GsrState), ctypes.c_void_p,
                               ctypes.POINTER(GsrPixelGrads), ctypes.POINTER(GsrParamGrads), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_backward_packets.restype = ctypes.c_int
    L.gsr_backward_packets.argtypes = [ctypes.POINTER(GsrView), ctypes.POINTER(GsrGaussians), ctypes.c_void_p, ctypes.POINTER(GsrState),
                                       ctypes.c_void_p, ctypes.POINTER(GsrPixelGrads), ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_gather_packets.restype = ctypes.c_int
    L.gsr_gather_packets.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.POINTER(GsrParamGrads),
                                     ctypes.c_void_p]
    L.gsr_gather_packets_v.restype = ctypes.c_int
    L.gsr_gather_packets_v.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32,
                                       ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32,
                                       ctypes.POINTER(GsrParamGrads), ctypes.c_void_p]
    L.gsr_peer_alloc.restype = ctypes.c_int
    L.gsr_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
    L.gsr_peer_open.restype = ctypes.c_int
    L.gsr_peer_open.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
    L.gsr_peer_close.restype = ctypes.c_int
    L.gsr_peer_close.argtypes = [ctypes.c_void_p]
    L.gsr_peer_free.restype = ctypes.c_int
    L.gsr_peer_free.argtypes = [ctypes.c_void_p]
    L.gsr_peer_copy.restype = ctypes.c_int
    L.gsr_peer_copy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_image_loss_scratch_bytes.restype = ctypes.c_size_t
    L.gsr_image_loss_scratch_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]
    L.gsr_image_loss.restype = ctypes.c_int
    L.gsr_image_loss.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_depth_loss_scratch_bytes.restype = ctypes.c_size_t
    L.gsr_depth_loss_scratch_bytes.argtypes = [ctypes.c_int64]
    L.gsr_depth_loss.restype = ctypes.c_int
    L.gsr_depth_loss.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_select_rows.restype = ctypes.c_int
    L.gsr_select_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64),
                                  ctypes.c_int32, ctypes.c_void_p]
    L.gsr_adam_step.restype = ctypes.c_int
    L.gsr_adam_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(GsrAdamGroup), ctypes.c_int32,
                                ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int32, ctypes.c_void_p]
    L.gsr_packet_index_words.restype = ctypes.c_size_t
    L.gsr_packet_index_words.argtypes = [ctypes.c_int32]
    L.gsr_last_num_visible.restype = ctypes.c_uint32
    L.gsr_mark_visible.restype = ctypes.c_int
    L.gsr_mark_visible.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.gsr_knn_workspace_bytes.restype = ctypes.c_size_t
    L.gsr_knn_workspace_bytes.argtypes = [ctypes.c_int32]
    L.gsr_knn_dist2.restype = ctypes.c_int
    L.gsr_knn_dist2.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.gsr_export_state.restype = ctypes.c_int
    L.gsr_export_state.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(GsrState), ctypes.POINTER(GsrStateExport),
                                   ctypes.c_void_p]
    L.gsr_microbench.restype = ctypes.c_int
    L.gsr_microbench.argtypes = [ctypes.POINTER(GsrMicrobench), ctypes.c_void_p]
    L.gsr_count_work.restype = ctypes.c_int
    L.gsr_count_work.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(GsrState), ctypes.c_void_p, ctypes.c_void_p]
    L.gsr_launch_count.restype = ctypes.c_ulonglong
    L.gsr_set_profiling.restype = None
    L.gsr_set_profiling.argtypes = [ctypes.c_int]
    L.gsr_get_stage_times.restype = ctypes.c_int
    L.gsr_get_stage_times.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_char_p)]
    if L.gsr_abi_version() != GSR_ABI_VERSION:
        raise ImportError("libgsr.so ABI version %d != %d (rebuild: python __graft_entry__.py)" % (L.gsr_abi_version(), GSR_ABI_VERSION))
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().gsr_last_error().decode("utf-8", "replace")
        raise GsrError("%s failed (%d): %s" % (what, rc, msg))


def stage_times():
    ms = (ctypes.c_float * GSR_STAGE_COUNT)()
    names = (ctypes.c_char_p * GSR_STAGE_COUNT)()
    n = lib().gsr_get_stage_times(ms, names)
    return {names[i].decode(): float(ms[i]) for i in range(n) if names[i]}
