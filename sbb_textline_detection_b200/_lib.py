"""ctypes binding of include/sbb_textline.h.  There is NO fallback: if the CUDA library is missing
or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SBB_LIB: A/B experiments against an older build of the SAME library (tools/); default = the in-tree build
LIB_PATH = os.environ.get("SBB_LIB") or os.path.join(HERE, "libsbb_textline.so")

SBB_PREC_FP16X3, SBB_PREC_FP16 = 0, 1
SBB_BACKEND_TCGEN05, SBB_BACKEND_SIMT = 0, 1
SBB_MEM_HOST, SBB_MEM_DEVICE = 0, 1

# every symbol include/sbb_textline.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sbb_abi_version", "sbb_last_error", "sbb_model_create", "sbb_model_destroy", "sbb_model_shape",
    "sbb_predict_page_tiled", "sbb_predict_tiles", "sbb_predict_full", "sbb_compute_tile_grid",
    "sbb_model_num_activations", "sbb_model_activation_info", "sbb_model_read_activation",
    "sbb_model_last_launch_count", "sbb_model_set_profiling", "sbb_model_num_layers",
    "sbb_model_layer_time", "sbb_resize_nearest_u8", "sbb_otsu_copy_u8", "sbb_morph5x5_u8", "sbb_rotate_rowsum_u8",
    "sbb_predict_page_tile_range", "sbb_peer_alloc", "sbb_peer_open", "sbb_peer_close", "sbb_peer_free",
    "sbb_plan_decoder_tiles", "sbb_plan_chain_list", "sbb_model_geom_cache_stats",
    "sbb_model_part_times", "sbb_model_set_precision_plan", "sbb_predict_pages_stacked", "sbb_nccl_unique_id", "sbb_nccl_comm_create", "sbb_nccl_comm_destroy", "sbb_model_broadcast",
]


class ModelDesc(C.Structure):
    _fields_ = [("tile_h", C.c_int32), ("tile_w", C.c_int32), ("n_classes", C.c_int32),
                ("precision", C.c_int32), ("backend", C.c_int32), ("device", C.c_int32),
                ("max_batch", C.c_int32), ("reserved", C.c_int32),
                ("weights", C.c_void_p), ("weights_nbytes", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sbb_textline_detection_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the segmentation hot path.")
    l = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    l.sbb_abi_version.restype = C.c_int
    l.sbb_last_error.restype = C.c_char_p
    l.sbb_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    l.sbb_model_destroy.argtypes = [vp]
    l.sbb_model_destroy.restype = None
    l.sbb_model_shape.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.sbb_predict_page_tiled.argtypes = [vp, vp, i32, i32, i64, i32, vp, i64, i32, vp]
    l.sbb_predict_pages_stacked.argtypes = [vp, vp, i32, i32, i32, i64, i32, vp, i64, i32, vp]
    l.sbb_predict_tiles.argtypes = [vp, vp, i32, vp, vp, vp, i32, vp]
    l.sbb_predict_full.argtypes = [vp, vp, vp, i32, vp]
    l.sbb_compute_tile_grid.argtypes = [i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), vp, i32, vp, vp]
    l.sbb_plan_decoder_tiles.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), vp, i32,
                                         C.POINTER(i32)]
    l.sbb_plan_chain_list.argtypes = [i64, i32, i32, i32, vp, i32, C.POINTER(i32)]
    l.sbb_model_num_activations.argtypes = [vp]
    l.sbb_model_activation_info.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    l.sbb_model_read_activation.argtypes = [vp, i32, i32, vp]
    l.sbb_model_last_launch_count.argtypes = [vp]
    l.sbb_model_last_launch_count.restype = i64
    l.sbb_model_geom_cache_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    l.sbb_model_set_profiling.argtypes = [vp, i32]
    l.sbb_model_num_layers.argtypes = [vp]
    l.sbb_model_layer_time.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    l.sbb_resize_nearest_u8.argtypes = [vp, i32, i32, i32, i64, vp, i32, i32, i64, i32, i32, vp]
    l.sbb_otsu_copy_u8.argtypes = [vp, i32, i32, i32, i64, vp, i64, C.POINTER(i32), i32, i32, vp]
    l.sbb_morph5x5_u8.argtypes = [vp, i32, i32, i32, i64, vp, i64, i32, i32, i32, i32, vp]
    l.sbb_rotate_rowsum_u8.argtypes = [vp, i32, i32, i64, i32, i32, i32, vp, i32, vp, i32, i32, vp]
    l.sbb_predict_page_tile_range.argtypes = [vp, vp, i32, i32, i64, i32, vp, i64, i32, i32, i32, vp]
    l.sbb_model_part_times.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i32), i32]
    l.sbb_model_set_precision_plan.argtypes = [vp, C.c_char_p]
    l.sbb_nccl_unique_id.argtypes = [vp]
    l.sbb_nccl_comm_create.argtypes = [vp, i32, i32, i32, C.POINTER(vp)]
    l.sbb_nccl_comm_destroy.argtypes = [vp]
    l.sbb_model_broadcast.argtypes = [vp, C.c_size_t, i32, vp, i32, vp]
    l.sbb_peer_alloc.argtypes = [i32, C.c_size_t, C.POINTER(vp), vp]
    l.sbb_peer_open.argtypes = [i32, vp, C.POINTER(vp)]
    l.sbb_peer_close.argtypes = [vp]
    l.sbb_peer_free.argtypes = [vp]
    _lib = l
    return l


class SbbError(RuntimeError):
    """A call into libsbb_textline.so failed (CUDA error, bad argument, ...).  Callers that sit under one of the
    reference's bare ``except:`` blocks (main.py:1736-1739, 2148) use the type to tell a broken hot path from
    the numerical failures those blocks were written for."""

    def __init__(self, code: int, message: str):
        super().__init__(f"sbb_textline error {code}: {message}")
        self.code = code


def check(rc: int):
    if rc != 0:
        raise SbbError(rc, lib().sbb_last_error().decode(errors='replace'))
