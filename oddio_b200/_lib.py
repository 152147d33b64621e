"""ctypes binding of liboddio_b200.so (include/oddio_b200.h). No torch, no CPU fallback.

The library is built in-tree by ``python -m oddio_b200.build``; importing this module never
compiles anything and fails loudly if the shared object is missing.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ODB_SO") or os.path.join(HERE, "liboddio_b200.so")

ODB_OK = 0
ODB_E_INVALID = -1
ODB_E_CUDA = -2
ODB_E_UNSUPPORTED = -3
ODB_E_NOMEM = -4

CHAIN_SPEED = 0x1
CHAIN_FIXED_GAIN = 0x2
CHAIN_GAIN = 0x4
CHAIN_CYCLE = 0x8

EPILOGUE_NONE = 0
EPILOGUE_TANH = 1
EPILOGUE_REINHARD = 2


class OddioError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oddio_b200 error {code}: {msg}")
        self.code = code


class Chain(C.Structure):
    """struct odb_chain"""

    _fields_ = [
        ("frames", C.c_uint64),
        ("start_seconds", C.c_double),
        ("flags", C.c_uint32),
        ("speed", C.c_float),
        ("fixed_gain_db", C.c_float),
        ("gain_ratio", C.c_float),
    ]


vp, f32, f64, u32, u64, i32 = C.c_void_p, C.c_float, C.c_double, C.c_uint32, C.c_uint64, C.c_int
fp = C.POINTER(C.c_float)
pvp = C.POINTER(C.c_void_p)
pu64 = C.POINTER(C.c_uint64)
pi32 = C.POINTER(C.c_int)

# name -> argtypes; every one returns int status except the two noted below. This table is the
# Python statement of include/oddio_b200.h and tests/test_abi.py checks the two against each other.
SIGNATURES = {
    "odb_ctx_create": [i32, pvp],
    "odb_ctx_create_on_stream": [i32, vp, pvp],
    "odb_ctx_destroy": [vp],
    "odb_ctx_synchronize": [vp],
    "odb_ctx_stream": [vp, pvp],
    "odb_pin_buffer": [vp, vp, u64],
    "odb_unpin_buffer": [vp, vp],
    "odb_frames_from_slice": [vp, u32, i32, fp, u64, pu64],
    "odb_frames_from_device": [vp, u32, i32, vp, u64, pu64],
    "odb_frames_release": [vp, u64],
    "odb_scene_create": [vp, pvp],
    "odb_scene_destroy": [vp],
    "odb_scene_set_epilogue": [vp, i32],
    "odb_scene_play": [vp, C.POINTER(Chain), fp, fp, f32, pu64],
    "odb_scene_play_buffered": [vp, C.POINTER(Chain), fp, fp, f32, f32, u32, f32, pu64],
    "odb_scene_set_listener_rotation": [vp, fp],
    "odb_spatial_set_motion": [vp, u64, fp, fp, i32],
    "odb_spatial_set_motion_many": [vp, u32, pu64, fp, fp, C.POINTER(C.c_uint8)],
    "odb_spatial_is_finished": [vp, u64, pi32],
    "odb_scene_sample": [vp, f32, fp, u32],
    "odb_scene_run": [vp, u32, fp, u32],
    "odb_scene_sample_device": [vp, f32, vp, u32],
    "odb_scene_len": [vp, i32, pu64],
    "odb_mixer_create": [vp, i32, pvp],
    "odb_mixer_destroy": [vp],
    "odb_mixer_set_epilogue": [vp, i32],
    "odb_mixer_play": [vp, C.POINTER(Chain), pu64],
    "odb_mixed_stop": [vp, u64],
    "odb_mixed_is_stopped": [vp, u64, pi32],
    "odb_mixer_sample": [vp, f32, fp, u32],
    "odb_mixer_run": [vp, u32, fp, u32],
    "odb_mixer_sample_device": [vp, f32, vp, u32],
    "odb_mixer_len": [vp, pu64],
    "odb_source_set_speed": [vp, u64, f32],
    "odb_source_set_amplitude_ratio": [vp, u64, f32],
    "odb_source_set_gain_db": [vp, u64, f32],
    "odb_source_playback_position": [vp, u64, C.POINTER(C.c_double)],
    "odb_source_frames_is_finished": [vp, u64, pi32],
    "odb_source_cursor": [vp, u64, C.POINTER(C.c_double), fp],
    "odb_last_launch_count": [vp, C.POINTER(C.c_uint32)],
    "odb_last_job_counters": [vp, C.POINTER(C.c_uint32)],
    "odb_set_profiling": [vp, i32],
    "odb_last_mix_kernel_ms": [vp, fp],
    "odb_set_kernel_variant": [vp, i32],
    "odb_frames_from_i16": [vp, u32, i32, C.POINTER(C.c_int16), u64, i32, pu64],
    "odb_scene_sample_i16": [vp, f32, C.POINTER(C.c_int16), u32],
    "odb_mixer_sample_i16": [vp, f32, C.POINTER(C.c_int16), u32],
    "odb_exchange_create": [vp, i32, i32, u32, i32, pvp],
    "odb_exchange_destroy": [vp],
    "odb_exchange_export": [vp, vp],
    "odb_exchange_connect": [vp, vp],
    "odb_exchange_allreduce": [vp, vp, u32, i32, vp],
    "odb_exchange_push": [vp, vp, u32, vp],
    "odb_exchange_pull": [vp, vp, u32, i32, vp],
    "odb_scene_sample_exchange": [vp, vp, f32, vp, u32, i32, i32, pi32],
}
NON_STATUS = {"odb_last_error": (C.c_char_p, []), "odb_abi_version": (C.c_uint32, []),
              "odb_exchange_handle_size": (C.c_int, [])}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every symbol of the header. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise OddioError(ODB_E_CUDA, f"{SO_PATH} not built (run `python -m oddio_b200.build`); there is no CPU fallback")
    L = C.CDLL(SO_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    for name, (res, args) in NON_STATUS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status: int) -> None:
    if status != ODB_OK:
        raise OddioError(status, load().odb_last_error().decode("utf-8", "replace"))
