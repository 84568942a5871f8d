"""ctypes binding of libv2v_b200.so (the C ABI declared in include/v2v_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails the
caller gets a ``V2VError``.  Structures mirror the header field for field.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("V2V_B200_LIB") or os.path.join(_HERE, "lib", "libv2v_b200.so")   # env: A/B tuning only

# enums (include/v2v_b200.h)
NOISE_NONE, NOISE_EXPLICIT, NOISE_PHILOX = 0, 1, 2
THRES_PER_CLIP, THRES_PER_PIXEL = 0, 1
U8, I8, U16, I16, I32, I64, F32, F64 = range(8)
SCATTER_H5_DISCRETE, SCATTER_H5_INTERP, SCATTER_TORCH_DISCRETE, SCATTER_TORCH_BILINEAR = range(4)
POL_SIGNED, POL_POS_ONLY, POL_NEG_ONLY, POL_SPLIT = range(4)

ERR_NAMES = {0: "OK", -1: "INVALID_ARG", -2: "SHAPE", -3: "ALIGNMENT", -4: "CUDA", -5: "UNSUPPORTED", -6: "NO_DEVICE"}

_p = C.c_void_p


class V2VError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"v2v_b200: {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class EsimDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("num_bins", C.c_int32), ("frames_per_bin", C.c_int32),
        ("noise_mode", C.c_int32), ("put_noise_external", C.c_int32),
        ("threshold_mode", C.c_int32), ("frame_out_mode", C.c_int32),
        ("frames", _p), ("lut", _p), ("pos_thres", _p), ("neg_thres", _p), ("base_noise_std", _p),
        ("u0", _p), ("hot_noise", _p), ("base_gauss", _p),
        ("hot_pixel_fraction", _p), ("hot_pixel_std", _p),
        ("seed", C.c_uint64), ("clip_index_base", C.c_uint64),
        ("potential_in", _p), ("potential_out", _p),
        ("voxel", _p), ("voxel_row_stride", C.c_int64), ("voxel_plane_stride", C.c_int64),
        ("frame_out", _p), ("stats", _p),
        ("frame_index", _p), ("raw_frames_per_clip", C.c_int32), ("kernel_flags", C.c_int32), ("value_map", _p),
    ]


ESIM_FLAG_GENERIC, ESIM_FLAG_SMALL_FAST, ESIM_FLAG_SMALL_P1, ESIM_FLAG_STAGED = 1, 2, 4, 16
V2E_FLAG_GENERIC, V2E_FLAG_FAST, V2E_FLAG_DIVERGENT_DIV, V2E_FLAG_U8_INTENSITY = 1, 2, 4, 8


def esim_flag_geom(g: int) -> int:
    return (g & 0xF) << 8


class V2eDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("num_bins", C.c_int32), ("frames_per_bin", C.c_int32),
        ("noise_mode", C.c_int32), ("state_f32", C.c_int32),
        ("fps", C.c_double),
        ("cutoff_hz", C.c_double), ("leak_rate_hz", C.c_double), ("shot_noise_rate_hz", C.c_double),
        ("leak_jitter_fraction", C.c_double),
        ("frames", _p), ("lut", _p), ("pos_thres", _p), ("neg_thres", _p), ("noise_rate", _p),
        ("leak_randn", _p), ("pos_shot", _p), ("neg_shot", _p),
        ("shot_pos_scale", _p), ("shot_neg_scale", _p),
        ("pos_thres_nominal", C.c_double), ("neg_thres_nominal", C.c_double),
        ("seed", C.c_uint64), ("clip_index_base", C.c_uint64),
        ("voxel", _p), ("stats", _p),
        ("frame_index", _p), ("raw_frames_per_clip", C.c_int32), ("kernel_flags", C.c_int32), ("value_map", _p),
        ("thres_per_interval", C.c_int32), ("reserved1", C.c_int32),
    ]


class ScatterDesc(C.Structure):
    _fields_ = [
        ("num_events", C.c_int64),
        ("num_windows", C.c_int32), ("num_bins", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("mode", C.c_int32), ("polarity_mode", C.c_int32),
        ("xs_dtype", C.c_int32), ("ys_dtype", C.c_int32), ("ts_dtype", C.c_int32), ("ps_dtype", C.c_int32),
        ("out_dtype", C.c_int32),
        ("xs", _p), ("ys", _p), ("ts", _p), ("ps", _p),
        ("window_offsets", _p), ("voxel", _p), ("dropped", _p),
        ("workspace", _p), ("workspace_bytes", C.c_int64),
        ("unsorted", _p), ("kernel_flags", C.c_int32), ("tuning_splits", C.c_int32), ("tuning_smem_kb", C.c_int32),
        ("reserved0", C.c_int32),
    ]


SCATTER_FLAG_RANGES, SCATTER_FLAG_GENERIC_SCAN, SCATTER_FLAG_NO_PACKED16 = 1, 2, 4


class ImageDesc(C.Structure):
    _fields_ = [
        ("num_events", C.c_int64),
        ("H", C.c_int32), ("W", C.c_int32),
        ("bilinear", C.c_int32), ("padding", C.c_int32), ("clip_out_of_range", C.c_int32),
        ("xs_dtype", C.c_int32), ("ys_dtype", C.c_int32), ("ps_dtype", C.c_int32), ("out_dtype", C.c_int32),
        ("xs", _p), ("ys", _p), ("ps", _p), ("image", _p), ("dropped", _p),
    ]


# every symbol include/v2v_b200.h declares: (name, restype, argtypes)
SYMBOLS = {
    "v2v_abi_version": (C.c_int, []),
    "v2v_last_error": (C.c_char_p, []),
    "v2v_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "v2v_launch_count": (C.c_longlong, []),
    "v2v_esim_frames_to_voxel": (C.c_int, [C.POINTER(EsimDesc), _p]),
    "v2v_esim_philox_fields": (C.c_int, [C.POINTER(EsimDesc), _p, _p, _p, _p]),
    "v2v_noise_direction_table": (C.c_int, [C.POINTER(C.c_uint32)]),
    "v2v_rng_words": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, _p, _p]),
    "v2v_v2e_frames_to_voxel": (C.c_int, [C.POINTER(V2eDesc), _p]),
    "v2v_v2e_shot_scales": (C.c_int, [C.POINTER(V2eDesc), _p, _p, _p]),
    "v2v_v2e_philox_fields": (C.c_int, [C.POINTER(V2eDesc), _p, _p, _p, _p]),
    "v2v_events_to_voxel": (C.c_int, [C.POINTER(ScatterDesc), _p]),
    "v2v_scatter_workspace_bytes": (C.c_int64, [C.POINTER(ScatterDesc)]),
    "v2v_events_to_image": (C.c_int, [C.POINTER(ImageDesc), _p]),
    "v2v_voxel_add_noise": (C.c_int, [_p, C.c_int64, _p, _p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _p]),
    "v2v_voxel_add_map": (C.c_int, [_p, C.c_int64, C.c_int64, _p, _p]),
    "v2v_voxel_bin_abs_sums": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, _p, _p]),
    "v2v_put_accumulate_bins": (C.c_int, [_p, C.c_int64, _p, _p, C.c_int64, C.c_int32, C.c_int64, _p, _p]),
    "v2v_take_bins": (C.c_int, [_p, C.c_int64, _p, _p, C.c_int64, C.c_int32, C.c_int64, _p]),
    "v2v_voxel_value_hist": (C.c_int, [_p, C.c_int32, C.c_int64, _p, _p]),
    "v2v_voxel_normalize": (C.c_int, [_p, C.c_int32, C.c_int64, _p, _p, _p]),
    "v2v_bgr_to_gray": (C.c_int, [_p, C.c_int32, _p, C.c_int64, _p]),
    "v2v_searchsorted_f64": (C.c_int, [_p, C.c_int64, _p, C.c_int64, _p, _p]),
    "v2v_pack_events_n5": (C.c_int, [_p, C.c_int, _p, C.c_int, _p, C.c_int, _p, C.c_int, C.c_int64, _p, _p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise V2VError(-6, f"{LIB_PATH} is missing: run `python -m v2v_b200.build` "
                           "(there is no CPU fallback for the video-to-voxel path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.v2v_abi_version() != 2:
        raise V2VError(-5, f"ABI version {lib.v2v_abi_version()} != 2")
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().v2v_last_error()
        raise V2VError(code, msg.decode() if msg else "")


def launch_count() -> int:
    return int(load().v2v_launch_count())


def device_info(device: int = 0):
    maj, mnr, sms = C.c_int(), C.c_int(), C.c_int()
    check(load().v2v_device_info(device, C.byref(maj), C.byref(mnr), C.byref(sms)))
    return maj.value, mnr.value, sms.value
