"""ctypes binding of libtruely_b200.so (include/truely_b200.h).  No fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRL_LIB_PATH") or os.path.join(_HERE, "libtruely_b200.so")      # override: experiment builds only

TRL_OK, TRL_E_INVALID, TRL_E_CUDA, TRL_E_CAPACITY, TRL_E_NOMEM, TRL_E_STATE = 0, -1, -2, -3, -4, -5
EMB_DIM = 512
MAX_SCALES = 24
NUM_STAGES = 13


class TrlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtruely_b200 error {code}: {msg}")
        self.code = code


class Weights(C.Structure):
    _fields_ = [("h_pnet", C.POINTER(C.c_float)), ("pnet_len", C.c_size_t),
                ("h_rnet", C.POINTER(C.c_float)), ("rnet_len", C.c_size_t),
                ("h_onet", C.POINTER(C.c_float)), ("onet_len", C.c_size_t),
                ("h_facenet", C.POINTER(C.c_float)), ("facenet_len", C.c_size_t)]


class Config(C.Structure):
    _fields_ = [("min_face_size", C.c_int), ("thresholds", C.c_float * 3), ("factor", C.c_double),
                ("crop_size", C.c_int), ("cand_cap_scale", C.c_int), ("cand_cap_frame", C.c_int),
                ("box_cap_frame", C.c_int), ("facenet_impl", C.c_int), ("pnet_precision", C.c_int),
                ("mode", C.c_int), ("margin", C.c_int)]


class Stamp(C.Structure):
    _fields_ = [("ox", C.c_int), ("oy", C.c_int), ("w", C.c_int), ("h", C.c_int), ("idx_off", C.c_longlong)]


_P = C.c_void_p
# name -> (restype, argtypes); exactly the symbols include/truely_b200.h declares
SIGNATURES = {
    "trl_default_config": (None, [C.POINTER(Config)]),
    "trl_facenet_blob_len": (C.c_size_t, []),
    "trl_pnet_precision": (C.c_int, [_P]),
    "trl_create": (C.c_int, [C.c_int, C.POINTER(Weights), C.POINTER(Config), C.POINTER(_P)]),
    "trl_destroy": (None, [_P]),
    "trl_last_error": (C.c_char_p, [_P]),
    "trl_pyramid_geometry": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "trl_pyramid": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "trl_pyramid_pairs_size": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "trl_pyramid_pairs": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "trl_pnet_screen_maps": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "trl_pnet": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "trl_nms": (C.c_int, [_P, _P, _P, C.c_int, C.c_float, C.c_int, _P, _P, _P]),
    "trl_crop_resample": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P]),
    "trl_rnet": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "trl_onet": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "trl_detect": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "trl_crop_align": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P]),
    "trl_facenet": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "trl_facenet_valid": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "trl_facenet_norm": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "trl_extract_face": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "trl_extract_faces_all": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int,
                                        _P, _P, _P, _P, _P, _P]),
    "trl_consistency": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "trl_consistency_clips": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "trl_shard_record_bytes": (C.c_size_t, [C.c_int]),
    "trl_shard_pack": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "trl_shard_resolve": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "trl_overlay_set_stamps": (C.c_int, [_P, _P, C.c_int, C.POINTER(Stamp), C.c_int, _P, C.c_longlong, C.c_int]),
    "trl_overlay": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "trl_process": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "trl_detect_align": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "trl_detect_align_async": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "trl_pipeline_join": (C.c_int, [_P, _P]),
    "trl_set_capacity": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "trl_check_capacity": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "trl_launch_count": (C.c_longlong, [_P]),
    "trl_set_profiling": (C.c_int, [_P, C.c_int]),
    "trl_read_stage_times": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "trl_stage_name": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "trl_host_alloc": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(_P)]),
    "trl_host_free": (C.c_int, [_P]),
}
# validation-only exports (not part of the public header)
DEBUG_SIGNATURES = {
    "trl_debug_facenet_num_layers": (C.c_int, [_P]),
    "trl_debug_facenet_layer": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "trl_debug_facenet_output": (C.c_int, [_P, C.c_int, C.c_int, _P]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol.  Raises if it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for table in (SIGNATURES, DEBUG_SIGNATURES):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)      # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib
