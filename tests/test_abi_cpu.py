"""CPU: the C-ABI library loads, exports every symbol include/v2v_b200.h declares, mirrors the structs, and
rejects bad arguments before touching CUDA (no compute without a GPU)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "v2v_b200.h")


@pytest.fixture(scope="module")
def lib():
    from v2v_b200 import _lib, build
    build.build()
    return _lib


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(v2v_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 10
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(v2v_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(lib.SYMBOLS), (sorted(set(names) ^ set(lib.SYMBOLS)))
    h = lib.load()
    assert h.v2v_abi_version() == 2
    assert h.v2v_launch_count() >= 0


def test_struct_layout_matches_header(lib, tmp_path):
    """sizeof / offsetof of every descriptor as the C compiler sees them == the ctypes mirrors."""
    structs = {"v2v_esim_desc": lib.EsimDesc, "v2v_v2e_desc": lib.V2eDesc, "v2v_scatter_desc": lib.ScatterDesc,
               "v2v_image_desc": lib.ImageDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_argument_validation_without_gpu(lib):
    h = lib.load()
    assert h.v2v_esim_frames_to_voxel(None, None) == -1
    assert b"NULL" in h.v2v_last_error()
    d = lib.EsimDesc()
    d.B, d.N, d.H, d.W, d.num_bins, d.frames_per_bin = 1, 8, 4, 4, 5, 1          # (N-1) % 5 != 0
    assert h.v2v_esim_frames_to_voxel(C.byref(d), None) == -2                   # data/v2v_datasets.py:365
    assert b"multiple" in h.v2v_last_error()
    d.N = 11
    assert h.v2v_esim_frames_to_voxel(C.byref(d), None) == -1                   # NULL data pointers
    d.B = 0
    assert h.v2v_esim_frames_to_voxel(C.byref(d), None) == 0                    # empty batch: nothing to do
    s = lib.ScatterDesc()
    s.num_bins, s.mode = 0, 0
    assert h.v2v_events_to_voxel(C.byref(s), None) == -1
    s.num_bins, s.mode = 5, 9
    assert h.v2v_events_to_voxel(C.byref(s), None) == -1
    s.mode, s.out_dtype = 0, lib.F32
    assert h.v2v_events_to_voxel(C.byref(s), None) == 0                         # zero windows
    v = lib.V2eDesc()
    v.B, v.N, v.H, v.W, v.num_bins, v.frames_per_bin, v.fps = 1, 7, 2, 2, 5, 1, 24.0
    assert h.v2v_v2e_frames_to_voxel(C.byref(v), None) == -2
    maj, mnr, sms = C.c_int(), C.c_int(), C.c_int()
    rc = h.v2v_device_info(0, C.byref(maj), C.byref(mnr), C.byref(sms))
    assert rc in (0, -6)


def test_no_cpu_fallback_in_product():
    """The product package never imports the oracle and refuses CPU tensors."""
    import torch
    import v2v_b200 as v2v
    pkg = os.path.join(ROOT, "v2v_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("CPU oracle", "").replace("the oracle", ""), fn
    with pytest.raises(v2v.V2VError):
        v2v.frames_to_voxel(torch.zeros((1, 6, 4, 4), dtype=torch.uint8), 0.2, 0.2, num_bins=5)


def test_host_sampling_law_matches_oracle():
    """sample_v2e_params (product host code) == oracle restatement == reference (pinned by the golden file)."""
    import v2v_oracle as orc
    from conftest import golden
    from v2v_b200 import V2VVoxelizer, sample_v2e_params
    for name in golden("esim").names("esimds_"):
        c = golden("esim").case(name)
        cfg = {str(k): eval(str(v)) for k, v in zip(c["cfg_keys"], c["cfg_vals"])}
        vz = V2VVoxelizer(cfg)
        fixed = (None, None) if float(c["fixed_pos"]) < 0 else (float(c["fixed_pos"]), float(c["fixed_neg"]))
        np.random.seed(int(c["seed"]))
        p = sample_v2e_params(vz, fixed[0], fixed[1])
        for k in p:
            assert p[k] == float(c[f"p_{k}"]), (name, k)
        np.random.seed(int(c["seed"]))
        q = orc.sample_esim_params(np.random, vz.threshold_range, vz.max_thres_pos_neg_gap, vz.base_noise_std_range,
                                   vz.hot_pixel_fraction_range, vz.hot_pixel_std_range, vz.scale_noise_strength,
                                   vz.put_noise_external, fixed[0], fixed[1])
        assert p == q


def test_luts_match_oracle():
    import v2v_oracle as orc
    from v2v_b200 import esim_log_lut
    from v2v_b200.v2e import v2e_log_lut
    assert np.array_equal(esim_log_lut(), orc.esim_log_lut())
    assert np.array_equal(v2e_log_lut(), orc.v2e_log_lut())


def test_reference_draw_order_helper():
    import v2v_oracle as orc
    from v2v_b200 import draw_reference_randomness
    np.random.seed(5)
    a = draw_reference_randomness(6, 5, 7, 0.3, 2.0)
    np.random.seed(5)
    b = orc.esim_draw_randomness(6, 5, 7, 0.3, 2.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
