"""Timing experiments only: build variants of libv2v_b200.so with parts of the throughput kernel knocked out
(V2V_ABL_* macros in csrc/esim_fast.cu; results are wrong by construction) to see what a launch's time is made of.

    python tools/build_ablations.py            # -> v2v_b200/lib/abl_<name>.so
    V2V_B200_LIB=v2v_b200/lib/abl_noise.so python tools/profile_esim.py --noise philox --clips 32 --iters 6 --time --stats
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from v2v_b200 import build as b  # noqa: E402

VARIANTS = {"mem": ["-DV2V_ABL_MEM"], "nopipe": ["-DV2V_EXP_NOPIPE"], "selcross": ["-DV2V_SEL_CROSS"], "kpf2": ["-DV2V_KPF=2"],
            "nopipe_selcross": ["-DV2V_EXP_NOPIPE", "-DV2V_SEL_CROSS"]}
if os.environ.get("ABL_ALL"):
    VARIANTS.update({"noise": ["-DV2V_ABL_NOISE"], "trigger": ["-DV2V_ABL_TRIGGER"], "cross": ["-DV2V_ABL_CROSS", "-DV2V_ABL_TRIGGER"],
                     "stats": ["-DV2V_ABL_STATS"],
                     "all": ["-DV2V_ABL_NOISE", "-DV2V_ABL_TRIGGER", "-DV2V_ABL_CROSS", "-DV2V_ABL_STATS"]})


def one(item):
    name, defs = item
    obj = os.path.join(b.LIBDIR, f"abl_{name}.o")
    lib = os.path.join(b.LIBDIR, f"abl_{name}.so")
    subprocess.run([b._nvcc()] + b.NVCC_FLAGS + defs + ["-c", os.path.join(b.CSRC, "esim_fast.cu"), "-o", obj], check=True)
    others = [os.path.join(b.LIBDIR, s[:-3] + ".o") for s in b.SOURCES if s != "esim_fast.cu"]
    subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, obj] + others, check=True)
    return lib


if __name__ == "__main__":
    b.build()
    with ThreadPoolExecutor(max_workers=6) as ex:
        for lib in ex.map(one, VARIANTS.items()):
            print(lib)
