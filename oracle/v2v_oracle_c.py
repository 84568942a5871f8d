"""ctypes access to the C restatement of the oracle (oracle/v2v_oracle_c.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "v2v_oracle_c.c")):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
        _lib.orc_esim_video_to_voxel.restype = None
        _lib.orc_make_voxel.restype = None
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def esim_video_to_voxel(video, pos, neg, base_noise_std, u0, hot, g, external=False, lut=None, return_state=False):
    """Same contract as v2v_oracle.esim_video_to_voxel (float64 [N-1,H,W])."""
    import v2v_oracle as orc
    video = np.ascontiguousarray(video, dtype=np.uint8)
    n, h, w = video.shape
    lut = np.ascontiguousarray(orc.esim_log_lut() if lut is None else lut, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    hot = None if hot is None else np.ascontiguousarray(hot, dtype=np.float64)
    g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
    out = np.empty((n - 1, h, w), dtype=np.float64)
    pot = np.empty((h, w), dtype=np.float64)
    load().orc_esim_video_to_voxel(_p(video), C.c_int(n), C.c_int64(h * w), _p(lut), C.c_double(pos), C.c_double(neg),
                                   C.c_double(base_noise_std), _p(u0), _p(hot), _p(g), C.c_int(int(external)), _p(out), _p(pot))
    return (out, pot) if return_state else out


def make_voxel(ts, xs, ys, ps, num_bins, height, width, interpolate_bins=False):
    """Same contract as v2v_oracle.make_voxel (float64 [bins,H,W])."""
    ts = np.ascontiguousarray(ts)
    f32 = ts.dtype == np.float32
    if not f32:
        ts = ts.astype(np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.int64)
    ys = np.ascontiguousarray(ys, dtype=np.int64)
    ps = np.ascontiguousarray(ps, dtype=np.uint8)
    vox = np.zeros((num_bins, height, width), dtype=np.float64)
    load().orc_make_voxel(_p(ts), C.c_int(int(f32)), _p(xs), _p(ys), _p(ps), C.c_int64(ts.shape[0]), C.c_int(num_bins),
                          C.c_int(height), C.c_int(width), C.c_int(int(interpolate_bins)), _p(vox))
    return vox
