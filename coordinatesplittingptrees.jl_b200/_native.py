"""ctypes bindings of the two in-tree native libraries.

libcsp_b200.so  -- the C-ABI CUDA library of include/csp_b200.h (hand-written sm_100a kernels).  This is the product:
                   every batched operation goes through it and there is no CPU fallback -- if it is missing or no CUDA
                   device is present the call raises.
libcsp_host.so  -- the host-side tree (World / Box / addpoint! / flatten), standing in for the Julia object graph.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)
OBJECTIVE = C.CFUNCTYPE(C.c_double, c_f64p, C.c_int, C.c_void_p)


class CspError(RuntimeError):
    """Base class of errors raised by the native libraries."""


class ErrorException(CspError):
    """Julia's ErrorException (e.g. reference src/tree.jl:352, src/cs2.jl:167)."""


class DimensionMismatch(CspError):
    """Julia's DimensionMismatch (reference src/tree.jl:353)."""


class ArgumentError(CspError):
    pass


class BoundsError(CspError):
    pass


class UndefRefError(CspError):
    pass


class NoDeviceError(CspError):
    """No CUDA device: the accelerator has no CPU fallback."""


class TreeDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("p", C.c_int32), ("n_nodes", C.c_int32), ("n_internal", C.c_int32), ("n_leaves", C.c_int32),
                ("lower", c_f64p), ("upper", c_f64p), ("parent", c_i32p), ("childindex", c_u8p), ("internal_id", c_i32p),
                ("leaf_id", c_i32p), ("val", c_f64p), ("inode", c_i32p), ("dims", c_i32p), ("xs", c_f64p), ("child", c_i32p),
                ("pos", c_f64p)]


_HOST_ERR = {-1: ErrorException, -2: DimensionMismatch, -3: AssertionError, -5: ArgumentError, -6: BoundsError}
_DEV_ERR = {-1: ArgumentError, -2: CspError, -3: MemoryError, -4: CspError, -5: NoDeviceError, -6: CspError}

_host = None
_dev = None


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise CspError(f"{name} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` (or make -C "
                       f"{os.path.join(_HERE, 'csrc')}) first; there is no fallback implementation")
    return C.CDLL(path)


def host():
    global _host
    if _host is None:
        L = _load("libcsp_host.so")
        L.csph_last_error.restype = C.c_char_p
        L.csph_tree_new.argtypes = [C.c_int, C.c_int, c_f64p, c_f64p, c_f64p, C.c_double, C.POINTER(C.c_void_p)]
        L.csph_tree_free.argtypes = [C.c_void_p]
        L.csph_tree_free.restype = None
        L.csph_split.argtypes = [C.c_void_p, C.c_int, C.c_int, c_i32p, c_f64p, C.c_int, c_f64p, c_i32p]
        L.csph_addpoint.argtypes = [C.c_void_p, C.c_int, c_f64p, C.c_int, C.c_int, c_i32p, c_i32p, C.c_void_p, C.c_void_p, c_i32p]
        L.csph_addpoints_cycle.argtypes = [C.c_void_p, c_f64p, C.c_long, C.c_int, C.c_int, c_i32p, c_i32p, c_i32p, C.c_void_p,
                                           C.c_void_p]
        L.csph_find_leaf_host.argtypes = [C.c_void_p, C.c_int, c_f64p, C.c_int, c_i32p]
        L.csph_counts.argtypes = [C.c_void_p] + [c_i32p] * 5
        L.csph_box_info.argtypes = [C.c_void_p, C.c_int, c_i32p, c_i32p, c_i32p, c_i32p, c_f64p, c_f64p, c_i32p, c_f64p]
        L.csph_boxbounds.argtypes = [C.c_void_p, C.c_int, C.c_int, c_f64p]
        L.csph_flatten.argtypes = [C.c_void_p, c_f64p, c_f64p, c_i32p, c_u8p, c_i32p, c_i32p, c_f64p, c_i32p, c_i32p, c_f64p, c_i32p,
                                   c_f64p, c_i32p, c_i32p]
        L.csph_split_count.argtypes = [C.c_void_p, c_i32p]
        L.csph_split_log.argtypes = [C.c_void_p, C.c_int, C.c_int, c_i32p, c_i32p, c_f64p, c_f64p]
        L.csph_quadnoise_new.argtypes = [C.c_int, c_f64p, C.c_double, C.c_uint64]
        L.csph_quadnoise_new.restype = C.c_void_p
        L.csph_quadnoise_free.argtypes = [C.c_void_p]
        L.csph_quadnoise_free.restype = None
        L.csph_objective_ptr.argtypes = [C.c_int]
        L.csph_objective_ptr.restype = C.c_void_p
        _host = L
    return _host


def hchk(rc):
    if rc != 0:
        raise _HOST_ERR.get(rc, CspError)(host().csph_last_error().decode())


def dev():
    global _dev
    if _dev is None:
        L = _load("libcsp_b200.so")
        vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
        L.csp_last_error.restype = C.c_char_p
        L.csp_kernel_launches.restype = C.c_int64
        L.csp_modelvalue_native.argtypes = [C.c_int32, C.c_double, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
        L.csp_timing_enable.argtypes = [C.c_int]
        L.csp_timing_read.argtypes = [c_f64p, c_i64p]
        L.csp_device_count.argtypes = [C.POINTER(C.c_int)]
        L.csp_set_device.argtypes = [C.c_int]
        L.csp_tree_upload.argtypes = [C.POINTER(TreeDesc), vpp]
        L.csp_tree_blob_size.argtypes = [C.POINTER(TreeDesc), C.POINTER(C.c_size_t)]
        L.csp_tree_pack.argtypes = [C.POINTER(TreeDesc), vp, C.c_size_t]
        L.csp_tree_upload_blob.argtypes = [vp, C.c_size_t, vpp]
        L.csp_tree_info.argtypes = [vp, c_i64p]
        L.csp_tree_create_live.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, C.c_double, vpp]
        L.csp_tree_append_splits.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
        L.csp_tree_live_maps.argtypes = [vp, vp, vp]
        L.csp_tree_export_derived.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.csp_tree_free.argtypes = [vp]
        L.csp_find_leaf.argtypes = [vp, vp, C.c_int64, vp, vp]
        L.csp_find_leaf_dev.argtypes = [vp, vp, C.c_int64, vp, vp, vp]
        L.csp_fit_boxes.argtypes = [vp, vp, C.c_int64, C.c_int32] + [vp] * 9
        L.csp_coefficients_p.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp, vp]
        L.csp_fit_gdiag.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, vp, vp, vp]
        L.csp_coefficients_p_sparse.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
        L.csp_models_fit.argtypes = [vp, C.c_int32, vpp]
        L.csp_models_refit.argtypes = [vp, C.c_int32, vp, vp]
        L.csp_models_possemidef.argtypes = [vp, vp]
        L.csp_models_upload.argtypes = [C.c_int32, C.c_int64, vp, vp, vp, vp, vpp]
        L.csp_models_download.argtypes = [vp, vp, vp, vp, vp, vp]
        L.csp_models_info.argtypes = [vp, c_i64p]
        L.csp_models_free.argtypes = [vp]
        L.csp_eval.argtypes = [vp, vp, vp, C.c_int64, vp]
        L.csp_eval_dev.argtypes = [vp, vp, vp, C.c_int64, vp, vp]
        L.csp_locate_eval.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp]
        L.csp_locate_eval_dev.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp]
        L.csp_models_alloc.argtypes = [vp, vpp]
        L.csp_models_refit_range.argtypes = [vp, C.c_int32, vp, C.c_int64, C.c_int64, vp]
        L.csp_models_download_rows.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp, vp]
        L.csp_find_leaf_from.argtypes = [vp, C.c_int32, vp, C.c_int64, vp, vp]
        L.csp_find_leaf_from_dev.argtypes = [vp, C.c_int32, vp, C.c_int64, vp, vp, vp]
        L.csp_boxbounds.argtypes = [vp, C.c_int32, vp, vp]
        L.csp_tree_extrema.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
        # multi-GPU layer (comm.cu)
        L.csp_init.argtypes = [C.c_int, c_i32p]
        L.csp_comm_unique_id.argtypes = [vp]
        L.csp_init_rank.argtypes = [C.c_int, C.c_int, C.c_int, vp]
        L.csp_comm_info.argtypes = [c_i64p]
        L.csp_shard_tree_broadcast.argtypes = [vp, C.c_size_t, C.c_int, vpp]
        L.csp_shard_tree_get.argtypes = [vp, C.c_int, vpp]
        L.csp_shard_tree_free.argtypes = [vp]
        L.csp_shard_fit.argtypes = [vp, C.c_int32, vpp, vpp]
        L.csp_shard_models_get.argtypes = [vp, C.c_int, vpp]
        L.csp_shard_models_free.argtypes = [vp]
        L.csp_shard_locate_eval_dev.argtypes = [vp, vp, vpp, c_i64p, vpp, vpp, vpp, C.c_int, vp, vp, C.c_int64, C.c_int64, vpp]
        L.csp_shard_gather_dev.argtypes = [vpp, vpp, c_i64p, C.c_int, vp, vp, C.c_int64, vpp]
        L.csp_shard_gather_alloc.argtypes = [C.c_int64, C.c_int, vpp, vpp]
        L.csp_shard_defer_join.argtypes = [C.c_int]
        L.csp_shard_join.argtypes = [vpp]
        L.csp_shard_locate_eval.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp]
        L.csp_shard_timing.argtypes = [c_f64p]
        _dev = L
    return _dev


def dchk(rc):
    if rc != 0:
        raise _DEV_ERR.get(rc, CspError)(dev().csp_last_error().decode())


EXPORTS = ["csp_version", "csp_last_error", "csp_device_count", "csp_set_device", "csp_tree_upload", "csp_tree_blob_size",
           "csp_tree_pack", "csp_tree_upload_blob", "csp_tree_info", "csp_tree_free", "csp_find_leaf", "csp_find_leaf_dev",
           "csp_fit_boxes", "csp_coefficients_p", "csp_models_fit", "csp_models_refit", "csp_models_upload", "csp_models_download", "csp_models_info",
           "csp_models_free", "csp_eval", "csp_eval_dev", "csp_locate_eval", "csp_locate_eval_dev", "csp_kernel_launches", "csp_timing_enable",
           "csp_timing_read", "csp_modelvalue_native", "csp_coefficients_p_sparse",
           "csp_models_possemidef", "csp_fit_gdiag", "csp_models_alloc", "csp_models_refit_range", "csp_models_download_rows",
           "csp_tree_export_derived", "csp_tree_create_live", "csp_tree_append_splits", "csp_tree_live_maps", "csp_find_leaf_from", "csp_find_leaf_from_dev", "csp_boxbounds", "csp_tree_extrema",
           "csp_init", "csp_comm_unique_id", "csp_init_rank", "csp_shutdown", "csp_comm_info", "csp_shard_tree_broadcast",
           "csp_shard_tree_get", "csp_shard_tree_free", "csp_shard_fit", "csp_shard_models_get", "csp_shard_models_free",
           "csp_shard_locate_eval_dev", "csp_shard_gather_dev", "csp_shard_gather_alloc", "csp_shard_gather_free", "csp_shard_defer_join", "csp_shard_join", "csp_shard_locate_eval", "csp_shard_synchronize", "csp_shard_timing"]
